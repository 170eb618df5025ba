// TEST INFRASTRUCTURE -- lets src/LineMatcher.cpp of the reference compile unmodified without the SLAM object graph.
// It is force-included (-include) ahead of the translation unit: it defines the include guards of Frame.h, KeyFrame.h,
// MapPoint.h, MapLine.h and Converter.h (so the real headers, which pull in DBoW2 / g2o / Sophus / boost, are skipped)
// and declares stand-ins carrying exactly the members LineMatcher.cpp touches.  Only the descriptor matchers
// (matchNNR, match, SerachForInitialize, SearchForTriangulation, distance, DescriptorDistance) and Fuse are called
// (Fuse with an identity pose and a unit pinhole: its own projection arithmetic is then exact); SearchByProjection and
// matchGrid compile against the stand-ins and are never reached.  lineDescriptorMAD and GetLinesInArea (Frame.cc /
// KeyFrame.cc, not compiled) forward to the oracle's restatement.
#pragma once
#define FRAME_H
#define KEYFRAME_H
#define MAPPOINT_H
#define MAPLINE_H
#define CONVERTER_H
#include <cmath>
#include <unordered_set>
#include "cvmini.hpp"
#include "eigenmini.hpp"
#include <line_descriptor_custom.hpp>
#include <line_descriptor/descriptor_custom.hpp>

using namespace std;
using namespace cv;
using namespace cv::line_descriptor;

extern "C" int plvio_lines_in_area(const unsigned char* keylines, int n, float x1, float y1, float x2, float y2, float r, int* out);
extern "C" void plvio_line_descriptor_mad(const int* d0, const int* d1, int n, double* nn_mad, double* nn12_mad);

namespace ORB_SLAM3 {

typedef Eigen::Vector6d Vector6d;
class KeyFrame;
class Frame;

class GeometricCamera {
 public:
  virtual ~GeometricCamera() {}
  virtual cv::Point2f project(const cv::Mat&) { cvmini_unreachable("GeometricCamera::project"); }
};

class MapPoint {};

class MapLine {
 public:
  cv::Mat mDesc, mNormal;
  Vector6d mWorldPos;
  bool mBad = false;
  int mnPredLevel = 0, mFusedIdx = -1;
  cv::Mat GetDescriptor() { return mDesc; }
  Vector6d GetWorldPos() { return mWorldPos; }
  bool isBad() { return mBad; }
  int Observations() { return 0; }
  void Replace(MapLine*) { cvmini_unreachable("MapLine::Replace"); }
  void AddObservation(KeyFrame*, size_t idx) { mFusedIdx = (int)idx; }   // bookkeeping recorded, not applied
  float GetMaxDistanceInvariance() { return 1e30f; }
  float GetMinDistanceInvariance() { return 0.0f; }
  cv::Mat GetNormal() { return mNormal; }
  int PredictScale(const float&, const float&) { return mnPredLevel; }
  int PredictScale(const float&, KeyFrame*) { return mnPredLevel; }
};

// Frame::lineDescriptorMAD / KeyFrame::lineDescriptorMAD (src/Frame.cc:1089-1113, src/KeyFrame.cc:411-435)
static inline void slam_mock_mad(const std::vector<std::vector<cv::DMatch>>& m, double& nn_mad, double& nn12_mad) {
  std::vector<int> d0(m.size()), d1(m.size());
  for (size_t i = 0; i < m.size(); i++) { d0[i] = (int)m[i][0].distance; d1[i] = (int)m[i][1].distance; }
  plvio_line_descriptor_mad(d0.data(), d1.data(), (int)m.size(), &nn_mad, &nn12_mad);
}

class Frame {
 public:
  cv::Mat mDescriptors_Line, mTcw;
  int N_l = 0;
  std::vector<MapLine*> mvpMapLines;
  std::vector<bool> mvbOutlier_Line;
  GeometricCamera* mpCamera = nullptr;
  float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0;
  std::vector<KeyLine> mvKeys_Line, mvKeysUn_Line;
  std::vector<float> mvScaleFactors_l;
  double inv_width = 0, inv_height = 0;
  void lineDescriptorMAD(std::vector<std::vector<cv::DMatch>> matches, double& nn_mad, double& nn12_mad) const {
    slam_mock_mad(matches, nn_mad, nn12_mad);
  }
};

class KeyFrame {
 public:
  cv::Mat mDescriptors_l;
  std::vector<MapLine*> mvpMapLines;
  float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0, mfLogScaleFactor = 0;
  std::vector<float> mvScaleFactors;
  std::vector<KeyLine> mvKeys_Line;
  void lineDescriptorMAD(std::vector<std::vector<cv::DMatch>> matches, double& nn_mad, double& nn12_mad) const {
    slam_mock_mad(matches, nn_mad, nn12_mad);
  }
  MapLine* GetMapLine(const size_t& i) { return i < mvpMapLines.size() ? mvpMapLines[i] : nullptr; }
  cv::Mat mRcw, mtcw, mOw;   // 3x3, 3x1, 3x1 CV_32F
  cv::Mat GetRotation() { return mRcw.clone(); }
  cv::Mat GetTranslation() { return mtcw.clone(); }
  cv::Mat GetCameraCenter() { return mOw.clone(); }
  std::vector<size_t> GetLinesInArea(const float& x1, const float& y1, const float& x2, const float& y2, const float& r,
                                     const int minLevel = -1, const int maxLevel = -1) const {
    if (minLevel > 0 || maxLevel > 0) cvmini_unreachable("KeyFrame::GetLinesInArea with levels");
    static_assert(sizeof(KeyLine) == 68, "KeyLine POD layout");
    std::vector<int> tmp(mvKeys_Line.size() + 1);
    const int k = plvio_lines_in_area((const unsigned char*)mvKeys_Line.data(), (int)mvKeys_Line.size(), x1, y1, x2, y2, r, tmp.data());
    return std::vector<size_t>(tmp.begin(), tmp.begin() + k);
  }
  void AddMapLine(MapLine*, const size_t&) {}   // recorded on the map line (AddObservation), not applied
};

class Converter {
 public:
  template <typename T> static cv::Mat toCvMat(const T&) { cvmini_unreachable("Converter::toCvMat"); }
};

}  // namespace ORB_SLAM3
