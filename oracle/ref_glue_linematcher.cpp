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
