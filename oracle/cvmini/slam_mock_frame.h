// TEST INFRASTRUCTURE -- lets src/Frame.cc of the reference compile unmodified WITH ITS OWN include/Frame.h: the headers
// Frame.h / Frame.cc pull in besides OpenCV (ImuTypes.h, ORBVocabulary.h, G2oTypes.h, Converter.h, MapPoint.h, MapLine.h,
// KeyFrame.h, the camera models) are guarded out and replaced by the stand-ins below.  Only member functions that touch
// plain data are called (oracle/ref_glue_frame.cpp); the constructors and the pose / IMU / stereo-fisheye code compile
// and are never reached.
#pragma once
#define IMUTYPES_H
#define ORBVOCABULARY_H
#define G2OTYPES_H
#define CONVERTER_H
#define MAPLINE_H
#define CAMERAMODELS_GEOMETRICCAMERA_H
#define CAMERAMODELS_PINHOLE_H
#define CAMERAMODELS_KANNALABRANDT8_H
#include <cmath>
#include <climits>
#include <map>
#include <set>
#include <list>
#include <tuple>
#include <vector>
#include <mutex>
#include <thread>
#include <iostream>
#include <unordered_set>
#include "cvmini.hpp"
#include "eigenmini.hpp"
#include <boost/serialization/array.hpp>
#define SLAM_MOCK_REAL_FRAME
#include "slam_mock_orb.h"   // stand-in MapPoint / KeyFrame / GeometricCamera (the set ORBmatcher.cc compiles against)

namespace Eigen {   // names only: the stereo-line code that uses them is never reached
template <typename T, int R, int C> struct Matrix;
template <> struct Matrix<double, 6, 1> : Vector6d {};
struct Matrix3d {};
static inline Vector3d operator+(const Vector3d& a, const Vector3d& b) { return Vector3d(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]); }
static inline Vector3d operator*(const Matrix3d&, const Vector3d&) { cv::cvmini_unreachable("Matrix3d * Vector3d"); }
}  // namespace Eigen

namespace ORB_SLAM3 {
class MapLine {
 public:
  bool mbTrackInView = false;
  float mTrackProjsX = 0, mTrackProjsY = 0, mTrackProjeX = 0, mTrackProjeY = 0, mnTrackangle = 0;
  // plain data filled by the glue (the drop-in build compiles the product's LineMatcher.cpp in this class set)
  cv::Mat mDesc, mNormal;
  Eigen::Matrix<double, 6, 1> mWorldPos;
  bool mBad = false;
  int mObs = 1, mnPredLevel = 0, mFusedIdx = -1;
  bool isBad() { return mBad; }
  int Observations() { return mObs; }
  Eigen::Matrix<double, 6, 1> GetWorldPos() { return mWorldPos; }
  cv::Mat GetNormal() { return mNormal; }
  cv::Mat GetDescriptor() { return mDesc; }
  float GetMaxDistanceInvariance() { return 1e30f; }
  float GetMinDistanceInvariance() { return 0.0f; }
  int PredictScale(const float&, const float&) { return mnPredLevel; }
  void Replace(MapLine* p) { p->mFusedIdx = mFusedIdx; }
  void AddObservation(KeyFrame*, size_t idx) { mFusedIdx = (int)idx; }
  // KeyFrame.cc
  int GetIndexInKeyFrame(KeyFrame*) { return -1; }
  std::map<KeyFrame*, size_t> GetObservations() { return std::map<KeyFrame*, size_t>(); }
  void EraseObservation(KeyFrame*) {}
};
class ConstraintPoseImu {};
namespace IMU {
class Bias { public: float bax = 0, bay = 0, baz = 0, bwx = 0, bwy = 0, bwz = 0; };
class Calib { public: cv::Mat Tcb, Tbc; };
class Preintegrated {
 public:
  void SetNewBias(const Bias&) { cvmini_unreachable("IMU::Preintegrated"); }
  void CopyFrom(Preintegrated*) { cvmini_unreachable("IMU::Preintegrated"); }
};
}  // namespace IMU
// ORBVocabulary / LineVocabulary = DBoW2::TemplatedVocabulary: only transform() is named (Frame::ComputeBoW; pinned
// separately through the reference's own DBoW2, libplvi_ref.so)
struct MockVocabulary {
  void transform(const std::vector<cv::Mat>&, DBoW2::BowVector&, DBoW2::FeatureVector&, int) { cvmini_unreachable("vocabulary"); }
};
typedef MockVocabulary ORBVocabulary;
typedef MockVocabulary LineVocabulary;
class Pinhole : public GeometricCamera {};
class KannalaBrandt8 : public GeometricCamera {
 public:
  std::vector<int> mvLappingArea;
  float TriangulateMatches(GeometricCamera*, const cv::KeyPoint&, const cv::KeyPoint&, const cv::Mat&, const cv::Mat&, const float,
                           const float, cv::Mat&) { cvmini_unreachable("KannalaBrandt8"); }
};
class Converter {
 public:
  static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& d) {   // src/Converter.cc:34-42: one row header per descriptor
    std::vector<cv::Mat> v;
    v.reserve(d.rows);
    for (int j = 0; j < d.rows; j++) v.push_back(d.row(j));
    return v;
  }
  template <typename T> static cv::Mat toCvMat(const T&) { cvmini_unreachable("Converter::toCvMat"); }
  static Eigen::Matrix3d toMatrix3d(const cv::Mat&) { cvmini_unreachable("Converter::toMatrix3d"); }
  static Eigen::Vector3d toVector3d(const cv::Mat&) { cvmini_unreachable("Converter::toVector3d"); }
};
}  // namespace ORB_SLAM3
