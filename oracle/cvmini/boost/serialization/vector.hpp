// TEST INFRASTRUCTURE -- stand-in (boost is absent in this image); see serialization.hpp
#pragma once
#include "serialization.hpp"
