// TEST INFRASTRUCTURE -- C entry points over the reference's LineMatcher (src/LineMatcher.cpp compiled unmodified
// with cvmini/slam_mock.h force-included in place of Frame.h / KeyFrame.h / MapLine.h / MapPoint.h / Converter.h;
// oracle/Makefile.ref).  Only the descriptor matchers are reachable; they touch the descriptor matrices, the
// per-line map-line pointers and lineDescriptorMAD of the stand-in Frame / KeyFrame and nothing else.
// This file is compiled with the same -include so that it sees the same stand-in classes as LineMatcher.cpp.
#include <cstring>
#include <vector>
#include "LineMatcher.h"   // /root/reference/include (its own includes of Frame.h etc. are guarded out by slam_mock.h)

using ORB_SLAM3::LineMatcher;

static cv::Mat desc_mat(const unsigned char* d, int n) {
  cv::Mat m(n, 32, CV_8UC1);
  for (int r = 0; r < n; r++) memcpy(m.ptr(r), d + 32 * (size_t)r, 32);
  return m;
}

// LineMatcher::matchNNR (src/LineMatcher.cpp:40-60)
extern "C" int plviref_line_match_nnr(const unsigned char* d1, int n1, const unsigned char* d2, int n2, float nnr, int* m12) {
  std::vector<int> m;
  const int k = LineMatcher::matchNNR(desc_mat(d1, n1), desc_mat(d2, n2), nnr, m);
  for (int i = 0; i < n1; i++) m12[i] = m[i];
  return k;
}

// LineMatcher::match(desc1, desc2, nnr, matches_12) (src/LineMatcher.cpp:91-111)
extern "C" int plviref_line_match(const unsigned char* d1, int n1, const unsigned char* d2, int n2, float nnr, int* m12) {
  std::vector<int> m;
  const int k = LineMatcher::match(desc_mat(d1, n1), desc_mat(d2, n2), nnr, m);
  for (int i = 0; i < n1; i++) m12[i] = m[i];
  return k;
}

// LineMatcher::match(vpLocalMapLines, CurrentFrame, nnr, matches_12) (src/LineMatcher.cpp:62-89): desc1 gathered from
// the map lines' GetDescriptor(), desc2 = the frame's line descriptors.
extern "C" int plviref_line_match_maplines(const unsigned char* d1, int n1, const unsigned char* d2, int n2, float nnr, int* m12) {
  std::vector<ORB_SLAM3::MapLine> lines(n1);
  std::vector<ORB_SLAM3::MapLine*> ptrs(n1);
  for (int i = 0; i < n1; i++) { lines[i].mDesc = desc_mat(d1 + 32 * (size_t)i, 1); ptrs[i] = &lines[i]; }
  ORB_SLAM3::Frame F;
  F.mDescriptors_Line = desc_mat(d2, n2);
  std::vector<int> m;
  const int k = LineMatcher::match(ptrs, F, nnr, m);
  for (int i = 0; i < n1; i++) m12[i] = m[i];
  return k;
}

// LineMatcher::SerachForInitialize (src/LineMatcher.cpp:113-141): pairs (queryIdx, trainIdx) in query order.
extern "C" int plviref_line_search_for_initialize(const unsigned char* d1, int n1, const unsigned char* d2, int n2, int* pairs) {
  ORB_SLAM3::Frame F1, F2;
  F1.mDescriptors_Line = desc_mat(d1, n1);
  F2.mDescriptors_Line = desc_mat(d2, n2);
  std::vector<std::pair<int, int>> out;
  LineMatcher lm;
  const int k = lm.SerachForInitialize(F1, F2, out);
  for (size_t i = 0; i < out.size(); i++) { pairs[2 * i] = out[i].first; pairs[2 * i + 1] = out[i].second; }
  return k;
}

// LineMatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs) (src/LineMatcher.cpp:143-171): has1 / has2 flag the
// lines that already own a map line (GetMapLine != NULL), which the function skips.
extern "C" int plviref_line_search_for_triangulation(const unsigned char* d1, int n1, const unsigned char* d2, int n2,
                                                     const unsigned char* has1, const unsigned char* has2, int* pairs) {
  ORB_SLAM3::KeyFrame K1, K2;
  ORB_SLAM3::MapLine some;
  K1.mDescriptors_l = desc_mat(d1, n1);
  K2.mDescriptors_l = desc_mat(d2, n2);
  K1.mvpMapLines.assign(n1, nullptr);
  K2.mvpMapLines.assign(n2, nullptr);
  for (int i = 0; i < n1; i++) if (has1 && has1[i]) K1.mvpMapLines[i] = &some;
  for (int i = 0; i < n2; i++) if (has2 && has2[i]) K2.mvpMapLines[i] = &some;
  std::vector<std::pair<size_t, size_t>> out;
  LineMatcher lm;
  const int k = lm.SearchForTriangulation(&K1, &K2, out);
  for (size_t i = 0; i < out.size(); i++) { pairs[2 * i] = (int)out[i].first; pairs[2 * i + 1] = (int)out[i].second; }
  return k;
}

// LineMatcher::distance / DescriptorDistance (src/LineMatcher.cpp:173-200, 466-...)
extern "C" int plviref_line_distance(const unsigned char* a, const unsigned char* b, int which) {
  return which ? LineMatcher::DescriptorDistance(desc_mat(a, 1), desc_mat(b, 1)) : LineMatcher::distance(desc_mat(a, 1), desc_mat(b, 1));
}

// LineMatcher::Fuse(pKF, vpMapLines, th) (src/LineMatcher.cpp:373-485): identity pose, fx = fy = 1, cx = cy = 0, map line i
// with endpoints (u1, v1, 1) / (u2, v2, 1), its normal along OM, unbounded distance invariance, PredictScale = level:
// every line passes the checks before the search and the reference's own projection yields the given endpoints
// exactly.  q: 6 floats per line = u1, v1, u2, v2, (unused), level; flags[i] != 0: isBad().  bounds = {mnMinX, mnMaxX,
// mnMinY, mnMaxY}.  best_idx[i] = keyline the reference fused map line i with (AddObservation), or -1.
extern "C" int plviref_line_fuse(const unsigned char* keylines, const unsigned char* desc, int n, const float* bounds,
                                 const float* scale_factors, int nlevels, const float* q, const unsigned char* flags,
                                 const unsigned char* qdesc, int nq, float th, int* best_idx) {
  ORB_SLAM3::KeyFrame K;
  K.mvKeys_Line.resize(n);
  memcpy((void*)K.mvKeys_Line.data(), keylines, (size_t)n * sizeof(KeyLine));
  K.mDescriptors_l = desc_mat(desc, n);
  K.mvpMapLines.assign(n, nullptr);
  K.fx = K.fy = 1; K.cx = K.cy = 0; K.mbf = 0;
  K.mnMinX = bounds[0]; K.mnMaxX = bounds[1]; K.mnMinY = bounds[2]; K.mnMaxY = bounds[3];
  K.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
  K.mRcw = cv::Mat::zeros(3, 3, CV_32F);
  for (int i = 0; i < 3; i++) K.mRcw.at<float>(i, i) = 1.0f;
  K.mtcw = cv::Mat::zeros(3, 1, CV_32F);
  K.mOw = cv::Mat::zeros(3, 1, CV_32F);
  std::vector<ORB_SLAM3::MapLine> lines(nq);
  std::vector<ORB_SLAM3::MapLine*> ptrs(nq);
  for (int i = 0; i < nq; i++) {
    ORB_SLAM3::MapLine& m = lines[i];
    m.mBad = flags && flags[i];
    const float* p = q + 6 * (size_t)i;
    m.mWorldPos(0) = p[0]; m.mWorldPos(1) = p[1]; m.mWorldPos(2) = 1.0;
    m.mWorldPos(3) = p[2]; m.mWorldPos(4) = p[3]; m.mWorldPos(5) = 1.0;
    m.mNormal = cv::Mat(3, 1, CV_32F);   // along OM = midpoint - Ow:  OM . pn = |OM|^2 >= 0.5 |OM|  (|OM| >= 1)
    m.mNormal.at<float>(0) = 0.5f * (p[0] + p[2]); m.mNormal.at<float>(1) = 0.5f * (p[1] + p[3]); m.mNormal.at<float>(2) = 1.0f;
    m.mnPredLevel = (int)p[5];
    m.mDesc = desc_mat(qdesc + 32 * (size_t)i, 1);
    ptrs[i] = &m;
  }
  LineMatcher lm;
  const int k = lm.Fuse(&K, ptrs, th);
  for (int i = 0; i < nq; i++) best_idx[i] = lines[i].mFusedIdx;
  return k;
}

// The line search of Frame::ComputeStereoMatches_Lines (src/Frame.cc:1421-1448): the grid and the direction table are
// filled exactly as that function does -- with the reference's own GridStructure / getLineCoords / LineIterator
// (src/gridStructure.cpp, src/LineIterator.cpp, compiled unmodified into this library) -- and handed to the reference's
// LineMatcher::matchGrid (src/LineMatcher.cpp:191-272).  seg = (startPointX, startPointY, endPointX, endPointY) per line.
#include "gridStructure.h"
extern "C" int plviref_line_match_grid(const float* seg1, const unsigned char* d1, int n1, const float* seg2,
                                       const unsigned char* d2, int n2, double inv_width, double inv_height, int grid_rows,
                                       int grid_cols, int win_left, int win_right, int win_up, int win_down, int* m12) {
  using namespace ORB_SLAM3;
  std::vector<line_2d> coords;
  for (int i = 0; i < n1; i++) {
    const float* s = seg1 + 4 * (size_t)i;
    coords.push_back(std::make_pair(std::make_pair(s[0] * inv_width, s[1] * inv_height),
                                    std::make_pair(s[2] * inv_width, s[3] * inv_height)));
  }
  std::list<std::pair<int, int>> line_coords;
  GridStructure grid(grid_rows, grid_cols);
  std::vector<std::pair<double, double>> directions(n2);
  for (int idx = 0; idx < n2; idx++) {
    const float* s = seg2 + 4 * (size_t)idx;
    std::pair<double, double>& v = directions[idx];
    v = std::make_pair((s[2] - s[0]) * inv_width, (s[3] - s[1]) * inv_height);
    normalize(v);
    getLineCoords(s[0] * inv_width, s[1] * inv_height, s[2] * inv_width, s[3] * inv_height, line_coords);
    for (const std::pair<int, int>& p : line_coords) grid.at(p.first, p.second).push_back(idx);
  }
  GridWindow w;
  w.width = std::make_pair(win_left, win_right);
  w.height = std::make_pair(win_up, win_down);
  std::vector<int> m;
  const int k = LineMatcher::matchGrid(coords, desc_mat(d1, n1), grid, desc_mat(d2, n2), directions, w, m);
  for (int i = 0; i < n1; i++) m12[i] = m[i];
  return k;
}
