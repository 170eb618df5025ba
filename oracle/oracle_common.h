// TEST INFRASTRUCTURE -- CPU oracle for the PL-VI-ORBSLAM3 front-end hot path.
//
// This directory restates, on the CPU and in scalar C++, the arithmetic of the
// reference path (ORBextractor / Lineextractor+LSD+LBD / ORBmatcher / LineMatcher)
// including the OpenCV primitives it calls (OpenCV is an un-vendored dependency of
// the reference: "4.2" per /root/reference/README.md:4; semantics restated here are
// those of OpenCV 4.x as probed with cv2 4.13, IPP off).
//
// PARITY PINNING (two layers, both bit-exact):
//  1. OpenCV-resident primitives (OpenCV is absent in this image: no headers, no libraries): restated here and
//     pinned against python cv2 4.13 (tests/test_oracle_vs_cv2.py in the build container, committed vectors
//     under tests/golden/).
//  2. Reference-owned logic (grid FAST loop, DistributeOctTree, IC_Angle, rBRIEF, LSD, KeyLine assembly, LBD,
//     Lineextractor, DBoW2 vocabulary loading + transform): pinned against the reference's OWN SOURCES, compiled unmodified from /root/reference by
//     oracle/Makefile.ref into oracle/_ref/libplvi_ref.so over the OpenCV/Eigen stand-in of oracle/cvmini/
//     (containers re-implemented; primitives = layer 1).  tests/test_oracle_vs_ref.py (live + the committed
//     outputs tests/golden/ref_outputs.npz): keypoints, descriptors, KeyLines, LBD bytes and line equations are
//     byte-identical.  Host-libm float functions the reference calls (cosf, sinf, atan2f) are restated from
//     glibc (namespace glibcm below) and equal the image's libm exhaustively.
//  3. The searches: src/ORBmatcher.cc and src/LineMatcher.cpp include the whole SLAM object graph (Frame, KeyFrame,
//     MapPoint, g2o, Sophus, boost).  They are compiled unmodified with a force-included header that defines those
//     headers' include guards and declares plain-data stand-ins (oracle/cvmini/slam_mock_orb.h, slam_mock.h), into
//     oracle/_ref/libplvi_ref_orbmatcher.so and libplvi_ref.so; tests/test_oracle_vs_ref_matchers.py and
//     test_oracle_vs_ref.py pin every search restated in oracle_match.cpp against them bit-exactly (identity poses:
//     the oracle boundary starts at the projected point).  Frame.cc / KeyFrame.cc / MapPoint.cc cannot be compiled
//     in the same library (their class definitions are the thing being replaced there): GetFeaturesInArea,
//     GetLinesInArea, lineDescriptorMAD are supplied to the compiled matchers by this restatement.  MapPoint.cc + MapPoint.h compile
//     unmodified over stand-in KeyFrame / Frame / Map (libplvi_ref_mappoint.so): ComputeDistinctiveDescriptors is
//     pinned.  Frame.cc + Frame.h compile unmodified over stand-in collaborators (slam_mock_frame.h,
//     libplvi_ref_frame.so): AssignFeaturesToGrid, GetFeaturesInArea, lineDescriptorMAD, UndistortKeyPoints/KeyLines
//     and ComputeStereoMatches are pinned (tests/test_oracle_vs_ref_frame.py); KeyFrame.cc + KeyFrame.h live in the
//     same library: KeyFrame::GetFeaturesInArea / GetLinesInArea / lineDescriptorMAD are pinned too.  Pinhole.cpp +
//     Pinhole.h compile unmodified (libplvi_ref_pinhole.so): the epipolar-line test is pinned for F12 = [t12]x.
//     See DESIGN.md section 2.
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

// cosf / sinf as the reference computes them.  Its sources call cos()/sin() on float arguments with <cmath>'s
// overloads in scope (src/ORBextractor.cc:110-111, src/LSD/lsd.cpp:678-679,
// Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:1131-1132), i.e. the host libm's cosf / sinf.
// glibc 2.28 .. 2.4x (the reference's Ubuntu 20.04 has 2.31, this image 2.39) implements them in double
// precision: quadrant reduction with a 2^24-scaled 2/pi, then a degree-7 / degree-8 polynomial, and rounds once
// to float (sysdeps/ieee754/flt-32/s_sincosf.h, from ARM's optimized routines; results are NOT always the
// correctly rounded ones: 0.04 % / 0.1 % of the floats in [0, 2 pi] differ from (float)cos((double)x)).  The
// restatement below uses the fused multiply-adds of the FMA build that x86-64 glibc selects at run time on every
// CPU with FMA3; it equals this image's libm bit for bit on all 2.2e9 floats with |x| < 120 (exhaustive scan,
// tools/scan_sincosf.c).  Valid for |x| < 120 (angles here lie in [-2 pi, 2 pi]).
namespace glibcm {
static inline uint32_t abstop12(float x) { uint32_t u; memcpy(&u, &x, 4); return (u >> 20) & 0x7ff; }
static inline float poly(double x, double x2, int tab, int n) {
  static const double C[2][5] = {{0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16},
                                 {-0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16}};
  static const double S[3] = {-0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13};
  if ((n & 1) == 0) {
    const double x3 = x * x2, s1 = std::fma(x2, S[2], S[1]), x7 = x3 * x2, s = std::fma(x3, S[0], x);
    return (float)std::fma(x7, s1, s);
  }
  const double* c = C[tab];
  const double x4 = x2 * x2, c2 = std::fma(x2, c[4], c[3]), c1 = std::fma(x2, c[1], c[0]), x6 = x4 * x2;
  return (float)std::fma(x6, c2, std::fma(x4, c[2], c1));
}
static inline double reduce_fast(double x, int* np) {
  const double r = x * 0x1.45F306DC9C883p+23;
  const int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return std::fma(-(double)n, 0x1.921FB54442D18p0, x);
}
static inline float sinf(float y) {
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) return abstop12(y) < abstop12(0x1p-12f) ? y : poly(x, x * x, 0, 0);
  int n;
  x = reduce_fast(x, &n);
  const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  return poly(x * s, x * x, (n & 2) ? 1 : 0, n);
}
static inline float cosf(float y) {
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) return abstop12(y) < abstop12(0x1p-12f) ? 1.0f : poly(x, x * x, 0, 1);
  int n;
  x = reduce_fast(x, &n);
  const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  return poly(x * s, x * x, (n & 2) ? 1 : 0, n ^ 1);
}

// atan2f as the reference computes KeyLine::angle (LSDDetector_custom.cpp:336 calls atan2() on two floats with
// <cmath>'s overloads in scope => the host libm's atan2f).  glibc < 2.41 (reference platform 2.31, this image
// 2.39): fdlibm's float algorithm (sysdeps/ieee754/flt-32/e_atan2f.c + s_atanf.c), plain float arithmetic without
// fused operations; about 12 % of its results are 1 ulp away from the correctly rounded value.  Checked against
// this image's libm: atanf on all 2^32 floats, atan2f on 5.6e8 pairs, 0 mismatches (tools/scan_sincosf.c).
static inline float atanf(float x) {
  static const float hi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  static const float lo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  static const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                               9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                               4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
  int32_t hx; memcpy(&hx, &x, 4);
  const int32_t ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) {   // |x| >= 2^25
    if (ix > 0x7f800000) return x + x;
    return hx > 0 ? hi[3] + lo[3] : -hi[3] - lo[3];
  }
  if (ix < 0x3ee00000) {    // |x| < 0.4375
    if (ix < 0x31000000) return x;
    id = -1;
  } else {
    x = std::fabs(x);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
      else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
      else { id = 3; x = -1.0f / x; }
    }
  }
  const float z = x * x, w = z * z;
  const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  const float r = hi[id] - ((x * (s1 + s2) - lo[id]) - x);
  return hx < 0 ? -r : r;
}
static inline float atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  int32_t hx, hy; memcpy(&hx, &x, 4); memcpy(&hy, &y, 4);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return atanf(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) return m < 2 ? y : (m == 2 ? pi + tiny : -pi - tiny);
  if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000 || iy == 0x7f800000) return std::atan2(y, x);   // infinities: never produced by the path
  const int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = atanf(std::fabs(y / x));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}
}  // namespace glibcm

}  // namespace plvio
