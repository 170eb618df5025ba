// TEST INFRASTRUCTURE -- CPU oracle for the PL-VI-ORBSLAM3 front-end hot path.
//
// This directory restates, on the CPU and in scalar C++, the arithmetic of the
// reference path (ORBextractor / Lineextractor+LSD+LBD / ORBmatcher / LineMatcher)
// including the OpenCV primitives it calls (OpenCV is an un-vendored dependency of
// the reference: "4.2" per /root/reference/README.md:4; semantics restated here are
// those of OpenCV 4.x as probed with cv2 4.13, IPP off).
//
// PARITY PINNING: the reference ships no tests, golden vectors or fixtures with
// expected outputs for this path, and it cannot be compiled in this image (no
// OpenCV C++ headers, no Eigen).  The OpenCV-resident primitives restated here are
// pinned against cv2 4.13 (tests/test_oracle_vs_cv2.py, run in the build container,
// and committed fixtures under tests/golden/).  The reference-owned logic (grid
// FAST loop, octree, IC_Angle, rBRIEF, LSD region growing, LBD, matchers) has no
// runnable reference => "parity unpinned" for those parts; see DESIGN.md.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (libplvi_cuda.so) never does.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace plvio {

typedef unsigned char u8;

// cvRound(float/double): round half to even (SSE cvtss2si semantics).
static inline int cv_round(double v) { return (int)std::nearbyint(v); }
static inline int cv_roundf(float v) { return (int)std::nearbyintf(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }

// cv::borderInterpolate(p, len, BORDER_REFLECT_101)
static inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

// cv::fastAtan2 (degrees), scalar model of OpenCV's atan_f32
// (modules/core/src/mathfuncs_core.simd.hpp); call sites in the reference:
// src/ORBextractor.cc:101, src/LSD/lsd.cpp:579,681,774-775.
float fast_atan2(float y, float x);

}  // namespace plvio
