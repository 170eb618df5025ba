// TEST INFRASTRUCTURE -- lets src/ORBmatcher.cc of the reference compile unmodified without the SLAM object graph.
// Force-included (-include) ahead of the translation unit: it defines the include guards of Frame.h, KeyFrame.h and
// MapPoint.h (the real headers pull in DBoW2's vocabulary, g2o, Sophus, boost serialization, the IMU types, the Atlas)
// and declares stand-ins carrying exactly the members ORBmatcher.cc touches.  State the matchers read (keys,
// descriptors, map points, feature vectors, scale tables, image bounds, projections cached on the map points) is
// plain data filled by oracle/ref_glue_orbmatcher.cpp; Frame::GetFeaturesInArea
// (src/Frame.cc:677-763; not compiled -- it belongs to Frame.cc) forwards to the oracle's restatement.  Pose / camera arithmetic (cv::Mat products) compiles against the stand-ins and aborts if reached:
// only the functions listed in ref_glue_orbmatcher.cpp are called.
#pragma once
#ifndef SLAM_MOCK_REAL_FRAME   // libplvi_ref_frame.so: the reference's own Frame.h / Frame.cc (see slam_mock_frame.h)
#define FRAME_H
#endif
#ifndef SLAM_MOCK_REAL_KEYFRAME   // libplvi_ref_keyframe.so: the reference's own KeyFrame.h / KeyFrame.cc (see slam_mock_keyframe.h)
#define KEYFRAME_H
#endif
#ifdef SLAM_MOCK_REAL_MAPLINE   // MapLine.cc + the reference's own MapLine.h (into libplvi_ref_mappoint.so): also needs Map, no Converter.h / g2o
#define CONVERTER_H
#include "eigenmini.hpp"
#endif
#ifdef SLAM_MOCK_REAL_MAPPOINT   // libplvi_ref_mappoint.so: the reference's own MapPoint.h / MapPoint.cc over stand-in KeyFrame / Frame / Map
#define MAP_H
#include <mutex>
#else
#define MAPPOINT_H
#endif
#include <cmath>
#include <map>
#include <set>
#include <list>
#include <vector>
#include "cvmini.hpp"
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

using namespace std;

namespace ORB_SLAM3 {
using cv::cvmini_unreachable;

class KeyFrame;
class Frame;
class MapPoint;
class MapLine;
class Map;

extern "C" int plvio_epipolar_constrain(float x1, float y1, float x2, float y2, const float* F12, float unc);

// Stand-in pinhole camera.  project = Pinhole::project (src/CameraModels/Pinhole.cpp:27-39: fx * x / z + cx in float).
// epipolarConstrain: Pinhole.cpp:135-157 builds F12 = K1^-T [t12]x R12 K2^-1 with cv::Mat products and a matrix inverse
// (not modelled by the stand-in) and then applies the point-to-epipolar-line test; here F12 is GIVEN (mF12, what the
// oracle and the CUDA path take as input) and the test itself (Pinhole.cpp:142-156) is the oracle's restatement, which is
// pinned against the reference's own Pinhole.cpp for F12 = [t12]x (libplvi_ref_pinhole.so, tests/test_oracle_vs_ref_frame.py).
class GeometricCamera {
 public:
  float fx = 1, fy = 1, cx = 0, cy = 0;
  float mF12[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  virtual ~GeometricCamera() {}
  virtual cv::Point2f project(const cv::Point3f& p) { return cv::Point2f(fx * p.x / p.z + cx, fy * p.y / p.z + cy); }
  virtual cv::Point2f project(const cv::Mat& m) {
    const float* p = m.ptr<float>();
    return project(cv::Point3f(m.at<float>(0), m.at<float>(1), m.at<float>(2)));
  }
  virtual cv::Mat toK() {   // Pinhole::toK (src/CameraModels/Pinhole.cpp:124-128)
    cv::Mat K = cv::Mat::zeros(3, 3, CV_32F);
    K.at<float>(0, 0) = fx; K.at<float>(0, 2) = cx; K.at<float>(1, 1) = fy; K.at<float>(1, 2) = cy; K.at<float>(2, 2) = 1.0f;
    return K;
  }
  unsigned int GetId() { return 0; }
  virtual float getParameter(const int i) { return i == 0 ? fx : (i == 1 ? fy : (i == 2 ? cx : cy)); }
  virtual float uncertainty2(const cv::Mat&) { cvmini_unreachable("GeometricCamera::uncertainty2"); }
  virtual bool epipolarConstrain(GeometricCamera*, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat&, const cv::Mat&,
                                 const float, const float unc) {
    return plvio_epipolar_constrain(kp1.pt.x, kp1.pt.y, kp2.pt.x, kp2.pt.y, mF12, unc) != 0;
  }
  virtual bool matchAndtriangulate(const cv::KeyPoint&, const cv::KeyPoint&, GeometricCamera*, cv::Mat&, cv::Mat&, const float,
                                   const float, cv::Mat&) { cvmini_unreachable("GeometricCamera::matchAndtriangulate"); }
};

// Frame::GetFeaturesInArea forwards to the oracle's restatement (oracle_match.cpp: plvio_grid_*)
extern "C" int plvio_grid_features_in_area(const void* g, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap);

#ifndef SLAM_MOCK_REAL_MAPPOINT
class MapPoint {
 public:
  // SearchByProjection(F, vpMapPoints) reads the projection cached by Frame::isInFrustum (src/Frame.cc:765-...)
  bool mbTrackInView = false, mbTrackInViewR = false;
  float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackProjYR = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackViewCos = 0,
        mTrackViewCosR = 0;
  int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
  long unsigned int mnLastFrameSeen = 0, mnFuseCandidateForKF = 0, mnId = 0;
  cv::Mat mDesc, mWorldPos;
  bool mBad = false;
  int mObs = 1;
  cv::Mat GetDescriptor() { return mDesc; }
  bool isBad() { return mBad; }
  int Observations() { return mObs; }
  cv::Mat GetWorldPos() { return mWorldPos; }
  // Fuse / SearchByProjection(KF, Scw): the checks before the search read these; the glue sets them so that every
  // point passes (the oracle / CUDA boundary starts after those checks); the bookkeeping after a hit is recorded, not applied
  cv::Mat mNormal;
  float mMaxDist = 1e30f, mMinDist = 0.0f;
  int mnPredLevel = 0, mFusedIdx = -1;
  cv::Mat GetNormal() { return mNormal; }
  float GetMaxDistanceInvariance() { return mMaxDist; }
  float GetMinDistanceInvariance() { return mMinDist; }
  int PredictScale(const float&, KeyFrame*) { return mnPredLevel; }
  int PredictScale(const float&, Frame*) { return mnPredLevel; }
  bool IsInKeyFrame(KeyFrame*) { return false; }
  // pMPinKF->Replace(pMP) (equal Observations(): always this direction): the query point inherits the feature it hit
  void Replace(MapPoint* p) { p->mFusedIdx = mFusedIdx; }
  void AddObservation(KeyFrame*, int idx) { mFusedIdx = idx; }
  std::tuple<int, int> GetIndexInKeyFrame(KeyFrame*) { return std::tuple<int, int>(-1, -1); }
  std::map<KeyFrame*, std::tuple<int, int>> GetObservations() { return std::map<KeyFrame*, std::tuple<int, int>>(); }   // KeyFrame.cc
  void EraseObservation(KeyFrame*) {}
};

#else
class MapLine;
class Map {
 public:
  long unsigned int mnId = 0;
  std::mutex mMutexPointCreation, mMutexLineCreation;
  void EraseMapPoint(MapPoint*) {}
  void EraseMapLine(MapLine*) {}
  long unsigned int GetId() { return mnId; }
};
#endif
#ifdef SLAM_MOCK_REAL_MAPLINE
class Converter {
 public:
  // Converter::toCvMat(Eigen::Matrix<double,3,1>) (src/Converter.cc): 3x1 CV_32F
  static cv::Mat toCvMat(const Eigen::Vector3d& m) {
    cv::Mat r(3, 1, CV_32F);
    for (int i = 0; i < 3; i++) r.at<float>(i) = (float)m(i);
    return r.clone();
  }
};
#endif

#ifndef SLAM_MOCK_REAL_FRAME
class Frame {
 public:
  int N = 0, Nleft = -1, Nright = -1;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
  std::vector<float> mvuRight, mvDepth;
  cv::Mat mDescriptors, mDescriptorsRight, mTcw, mTlr, mTrl;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec;
  GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
  float mbf = 0, mb = 0, fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0;
  float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0, mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  long unsigned int mnId = 0;
  const void* grid = nullptr;   // plvio_grid_create over mvKeysUn
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1,
                                        const bool bRight = false) const {
    if (bRight) cvmini_unreachable("Frame::GetFeaturesInArea(bRight)");
    std::vector<int> tmp(N > 0 ? N : 1);
    const int k = plvio_grid_features_in_area(grid, x, y, r, minLevel, maxLevel, tmp.data(), (int)tmp.size());
    return std::vector<size_t>(tmp.begin(), tmp.begin() + k);
  }
  cv::Mat GetRelativePoseTrl() { cvmini_unreachable("Frame"); }
  cv::Mat GetRelativePoseTlr() { cvmini_unreachable("Frame"); }
  // MapPoint.cc / MapLine.cc (only compiled into libplvi_ref_mappoint.so; never reached)
  struct KeyLineStub { int octave = 0; };
  std::vector<KeyLineStub> mvKeysUn_Line;
  std::vector<float> mvScaleFactors_l;
  int mnScaleLevels_l = 0;
  cv::Mat mDescriptors_Line;
  cv::Mat mRwc, mOw;
  int mnScaleLevels = 0;
  float mfLogScaleFactor = 0;
  cv::Mat GetCameraCenter() { cvmini_unreachable("Frame::GetCameraCenter"); }
};

#endif

#ifndef SLAM_MOCK_REAL_KEYFRAME
class KeyFrame {
 public:
  int N = 0, NLeft = -1, NRight = -1;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
  std::vector<float> mvuRight, mvDepth;
  cv::Mat mDescriptors;
  std::vector<MapPoint*> mvpMapPoints;
  DBoW2::BowVector mBowVec;
  DBoW2::FeatureVector mFeatVec;
  GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
  float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0, mThDepth = 0;
  float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0, mfLogScaleFactor = 0, mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
  int mnScaleLevels = 0;
  std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
  long unsigned int mnId = 0;
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  MapPoint* GetMapPoint(const size_t& i) { return mvpMapPoints[i]; }
  std::set<MapPoint*> GetMapPoints() { return std::set<MapPoint*>(); }
  const void* grid = nullptr;   // plvio_grid_create over mvKeysUn
  // KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:1200-1244) = the cell walk and radius test of the Frame version without
  // a level test: the oracle's restatement with minLevel = maxLevel = -1
  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const {
    if (bRight) cvmini_unreachable("KeyFrame::GetFeaturesInArea(bRight)");
    std::vector<int> tmp(N > 0 ? N : 1);
    const int k = plvio_grid_features_in_area(grid, x, y, r, -1, -1, tmp.data(), (int)tmp.size());
    return std::vector<size_t>(tmp.begin(), tmp.begin() + k);
  }
  bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }
  cv::Mat mRcw, mtcw, mOw;   // 3x3, 3x1, 3x1 CV_32F
  cv::Mat GetRotation() { return mRcw.clone(); }
  cv::Mat GetTranslation() { return mtcw.clone(); }
  cv::Mat GetCameraCenter() { return mOw.clone(); }
  cv::Mat GetPose() { cvmini_unreachable("KeyFrame"); }
  cv::Mat GetPoseInverse() { cvmini_unreachable("KeyFrame"); }
  cv::Mat GetRightPose() { cvmini_unreachable("KeyFrame"); }
  cv::Mat GetRightPoseInverse() { cvmini_unreachable("KeyFrame"); }
  cv::Mat GetRightCameraCenter() { cvmini_unreachable("KeyFrame"); }
  cv::Mat GetRightRotation() { cvmini_unreachable("KeyFrame"); }
  cv::Mat GetRightTranslation() { cvmini_unreachable("KeyFrame"); }
  cv::Mat GetRelativePoseTrl() { cvmini_unreachable("KeyFrame"); }
  cv::Mat GetRelativePoseTlr() { cvmini_unreachable("KeyFrame"); }
  void AddMapPoint(MapPoint*, const size_t&) {}   // recorded on the map point (AddObservation), not applied
  // MapPoint.cc / MapLine.cc (only compiled into libplvi_ref_mappoint.so)
  cv::Mat mDescriptors_l;
  std::vector<std::pair<float, float>> mvDepth_l;
  std::vector<float> mvScaleFactors_l;
  int mnScaleLevels_l = 0;
  void EraseMapLineMatch(const size_t&) {}
  void EraseMapLineMatch(MapLine*) {}
  void ReplaceMapLineMatch(const size_t&, MapLine*) {}
  bool mBad = false;
  long unsigned int mnFrameId = 0;
  bool isBad() { return mBad; }
  void EraseMapPointMatch(const int&) {}
  void EraseMapPointMatch(MapPoint*) {}
  void ReplaceMapPointMatch(const int&, MapPoint*) {}
  Map* GetMap() { return nullptr; }
};

#endif

}  // namespace ORB_SLAM3
