// TEST INFRASTRUCTURE -- lets src/LineMatcher.cpp of the reference compile unmodified without the SLAM object graph.
// It is force-included (-include) ahead of the translation unit: it defines the include guards of Frame.h, KeyFrame.h,
// MapPoint.h, MapLine.h and Converter.h (so the real headers, which pull in DBoW2 / g2o / Sophus / boost, are skipped)
// and declares stand-ins carrying exactly the members LineMatcher.cpp touches.  Only the descriptor matchers
// (matchNNR, match, SerachForInitialize, SearchForTriangulation, distance, DescriptorDistance) are called; the
// projection-based functions compile against the stand-ins and are never reached.  lineDescriptorMAD (Frame.cc /
// KeyFrame.cc, not compiled) forwards to the oracle's restatement.
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
  cv::Mat mDesc;
  cv::Mat GetDescriptor() { return mDesc; }
  Vector6d GetWorldPos() { cvmini_unreachable("MapLine::GetWorldPos"); }
  bool isBad() { return false; }
  int Observations() { return 0; }
  void Replace(MapLine*) { cvmini_unreachable("MapLine::Replace"); }
  void AddObservation(KeyFrame*, size_t) { cvmini_unreachable("MapLine::AddObservation"); }
  float GetMaxDistanceInvariance() { cvmini_unreachable("MapLine"); }
  float GetMinDistanceInvariance() { cvmini_unreachable("MapLine"); }
  cv::Mat GetNormal() { cvmini_unreachable("MapLine"); }
  int PredictScale(const float&, const float&) { cvmini_unreachable("MapLine"); }
  int PredictScale(const float&, KeyFrame*) { cvmini_unreachable("MapLine"); }
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
  cv::Mat GetRotation() { cvmini_unreachable("KeyFrame"); }
  cv::Mat GetTranslation() { cvmini_unreachable("KeyFrame"); }
  cv::Mat GetCameraCenter() { cvmini_unreachable("KeyFrame"); }
  std::vector<size_t> GetLinesInArea(const float&, const float&, const float&, const float&, const float&, const int = -1,
                                     const int = -1) const { cvmini_unreachable("KeyFrame::GetLinesInArea"); }
  void AddMapLine(MapLine*, const size_t&) { cvmini_unreachable("KeyFrame"); }
};

class Converter {
 public:
  template <typename T> static cv::Mat toCvMat(const T&) { cvmini_unreachable("Converter::toCvMat"); }
};

}  // namespace ORB_SLAM3
