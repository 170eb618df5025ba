// TEST INFRASTRUCTURE -- lets src/CameraModels/Pinhole.cpp of the reference compile unmodified with its own Pinhole.h and
// GeometricCamera.h: TwoViewReconstruction.h is guarded out (stand-in below), the Eigen overloads are named only.
// Called: Pinhole::epipolarConstrain (with unit intrinsics: K^-T and K^-1 are the identity, so F12 = [t12]x R12 is exact
// whatever the inverse's rounding), project, toK (oracle/ref_glue_pinhole.cpp).
#pragma once
#define TwoViewReconstruction_H
#include <vector>
#include "cvmini.hpp"
#include "eigenmini.hpp"
namespace Eigen {
template <typename T, int R, int C> struct Matrix {
  T v[R * C];
  T& operator()(int r, int c) { return v[r * C + c]; }
};
template <> struct Matrix<double, 2, 1> : Vector2d {};
template <> struct Matrix<double, 3, 1> : Vector3d {};
}  // namespace Eigen
namespace ORB_SLAM3 {
class Frame;
class TwoViewReconstruction {
 public:
  TwoViewReconstruction(const cv::Mat&) {}
  bool Reconstruct(const std::vector<cv::KeyPoint>&, const std::vector<cv::KeyPoint>&, const std::vector<int>&, cv::Mat&, cv::Mat&,
                   std::vector<cv::Point3f>&, std::vector<bool>&) { cv::cvmini_unreachable("TwoViewReconstruction"); }
  template <typename... A> bool ReconstructwithLine(A&&...) { cv::cvmini_unreachable("TwoViewReconstruction"); }
};
}  // namespace ORB_SLAM3
