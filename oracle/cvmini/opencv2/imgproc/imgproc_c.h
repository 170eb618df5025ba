// forwards to the stand-in (see cvmini.hpp)
#pragma once
#include "cvmini.hpp"
