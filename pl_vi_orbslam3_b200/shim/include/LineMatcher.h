// Drop-in for the reference's include/LineMatcher.h:24-109.  The kNN / grid searches run in libplvi_cuda.so; the two
// popcount helpers stay host-inline (pure functions on two 32-byte rows).
#pragma once

#include <cmath>
#include <list>
#include <unordered_set>
#include <utility>
#include <vector>

#include "plvi_cv_compat.h"

#ifdef PLVI_HAVE_OPENCV
#include <opencv2/core.hpp>
#include <opencv2/features2d.hpp>

#include "gridStructure.h"
#include "Frame.h"
#include "MapPoint.h"
#include "MapLine.h"
#endif

namespace ORB_SLAM3 {

class Frame;
class KeyFrame;
class MapPoint;
class MapLine;

// grid-cell coordinates of a line's two end points (include/LineMatcher.h:41-42)
typedef std::pair<int, int> point_2d;
typedef std::pair<point_2d, point_2d> line_2d;

// 2-D helpers Frame::ComputeStereoMatches_Lines uses on direction pairs (include/LineMatcher.h:44-53)
inline double dot(const std::pair<double, double>& a, const std::pair<double, double>& b) { return a.first * b.first + a.second * b.second; }
inline void normalize(std::pair<double, double>& v) {
  const double magnitude = std::sqrt(dot(v, v));
  v.first /= magnitude;
  v.second /= magnitude;
}

#ifdef PLVI_HAVE_OPENCV
// orderings of kNN-2 result rows declared by the reference header (include/LineMatcher.h:56-76)
struct compare_descriptor_by_NN_dist {
  inline bool operator()(const std::vector<cv::DMatch>& a, const std::vector<cv::DMatch>& b) { return a[0].distance < b[0].distance; }
};
struct conpare_descriptor_by_NN12_dist {
  inline bool operator()(const std::vector<cv::DMatch>& a, const std::vector<cv::DMatch>& b) {
    return (a[1].distance - a[0].distance) > (b[1].distance - b[0].distance);
  }
};
struct sort_descriptor_by_queryIdx {
  inline bool operator()(const std::vector<cv::DMatch>& a, const std::vector<cv::DMatch>& b) { return a[0].queryIdx < b[0].queryIdx; }
};
#endif

class LineMatcher {
 public:
  static const int TH_HIGH = 100, TH_LOW = 50;   // src/LineMatcher.cpp:37-38

  // include/LineMatcher.h:91-95 (src/LineMatcher.cpp:41-61, 92-111).  Like the reference, matches_12 is resize()d to
  // desc1.rows; every entry is then written (the reference leaves stale entries of a non-empty vector in place, none of
  // its call sites passes one).  With fewer than two rows in the train set the reference reads knnMatch's result out
  // of bounds; here such a call yields no match.
  static int matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12) { return knn(desc1, desc2, nnr, matches_12, 0); }
  static int match(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12) { return knn(desc1, desc2, nnr, matches_12, 1); }

  // src/LineMatcher.cpp:173-189 and the ">> 25" variant LineMatcher::Fuse uses (:487-499), kept bit-compatible
  static int distance(const cv::Mat& a, const cv::Mat& b) { return plvi_inline_hamming256(a.ptr(0), b.ptr(0), 0); }
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return plvi_inline_hamming256(a.ptr(0), b.ptr(0), 1); }

#ifdef PLVI_HAVE_OPENCV
  // include/LineMatcher.h:93-105 -- same signatures (shim/src/LineMatcher.cpp)
  static int match(const std::vector<MapLine*>& mvpLocalMapLines, Frame& CurrentFrame, float nnr, std::vector<int>& matches_12);
  int SerachForInitialize(Frame& InitialFrame, Frame& CurrentFrame, std::vector<std::pair<int, int> >& LineMatches);
  int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<std::pair<size_t, size_t> >& vMatchedPairs);
  static int matchGrid(const std::vector<line_2d>& lines1, const cv::Mat& desc1, const GridStructure& grid, const cv::Mat& desc2,
                       const std::vector<std::pair<double, double> >& directions2, const GridWindow& w, std::vector<int>& matches_12);
  // dead code in the reference (its call sites, Frame::grid_Line and the loop that fills it are commented out:
  // src/Tracking.cc:3972-3986); declared for source compatibility, throws when called
  static int SearchByProjection(Frame& CurrentFrame, Frame& LastFrame, const GridStructure& grid, const float& th, const float& angth);
  int Fuse(KeyFrame* pKF, const std::vector<MapLine*>& vpMapLines, const float th = 3.0);
#endif

  // The stereo line search of Frame::ComputeStereoMatches_Lines (src/Frame.cc:1421-1451) straight from the two KeyLine
  // sets: grid fill + matchGrid in one call (the unmodified Frame.cc keeps working through matchGrid above).
  static int matchStereoLines(const std::vector<cv::line_descriptor::KeyLine>& linesLeft, const cv::Mat& desc1,
                              const std::vector<cv::line_descriptor::KeyLine>& linesRight, const cv::Mat& desc2, double inv_width,
                              double inv_height, std::vector<int>& matches_12, int gridRows = 48, int gridCols = 64) {
    const int n1 = (int)linesLeft.size(), n2 = (int)linesRight.size();
    if (n1 != desc1.rows) throw std::runtime_error("[matchGrid] Each line needs a corresponding descriptor!");
    matches_12.assign(n1, -1);
    if (n1 == 0) return 0;
    std::vector<float> s1((size_t)n1 * 4), s2((size_t)(n2 > 0 ? n2 : 1) * 4);
    for (int i = 0; i < n1; i++) {
      const cv::line_descriptor::KeyLine& k = linesLeft[i];
      s1[4 * i] = k.startPointX; s1[4 * i + 1] = k.startPointY; s1[4 * i + 2] = k.endPointX; s1[4 * i + 3] = k.endPointY;
    }
    for (int i = 0; i < n2; i++) {
      const cv::line_descriptor::KeyLine& k = linesRight[i];
      s2[4 * i] = k.startPointX; s2[4 * i + 1] = k.startPointY; s2[4 * i + 2] = k.endPointX; s2[4 * i + 3] = k.endPointY;
    }
    std::vector<uint8_t> ta, tb;
    int nm = 0;
    plvi_shim::check(plvi_line_match_grid_host(plvi_shim::MatcherHandle::get(), s1.data(), plvi_shim::packed_rows(desc1, n1, ta), n1, s2.data(),
                                               plvi_shim::packed_rows(desc2, n2, tb), n2, inv_width, inv_height, gridRows, gridCols, 7, 0, 2,
                                               2, matches_12.data(), &nm),
                     "LineMatcher::matchStereoLines");
    return nm;
  }

 private:
  static int knn(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12, int mutual) {
    const int n1 = desc1.rows, n2 = desc2.rows;
    matches_12.resize(n1, -1);
    if (n1 == 0) return 0;
    std::vector<int> fresh(n1, -1);
    int nm = 0;
    if (n2 >= 2) {
      std::vector<uint8_t> ta, tb;
      plvi_shim::check(plvi_line_match(plvi_shim::MatcherHandle::get(), 1, plvi_shim::packed_rows(desc1, n1, ta), &n1, n1,
                                       plvi_shim::packed_rows(desc2, n2, tb), &n2, n2, nnr, mutual, fresh.data(), &nm, 0),
                       "LineMatcher::match");
    }
    for (int i = 0; i < n1; i++) matches_12[i] = fresh[i];
    return nm;
  }
};

}  // namespace ORB_SLAM3
