// TEST INFRASTRUCTURE -- stand-in (boost is absent in this image); see serialization.hpp
#pragma once
#include "serialization.hpp"
#ifndef BOOST_SERIALIZATION_ASSUME_ABSTRACT
#define BOOST_SERIALIZATION_ASSUME_ABSTRACT(T)
#endif
