// Drop-ins for the Hamming-search members of the reference's ORBmatcher
// (include/ORBmatcher.h:40-89) and LineMatcher (include/LineMatcher.h:87-107).  The searches
// that walk Frame / MapPoint graphs keep their host-side geometry in the caller (see
// INTEGRATION.md); what is forwarded here is the descriptor work.
#pragma once
#include <stdexcept>
#include "plvi_cv_compat.h"

namespace ORB_SLAM3 {

class PlviMatcherHandle {
 public:
  static plvi_matcher* get() {
    static PlviMatcherHandle inst;
    return inst.h_;
  }
 private:
  PlviMatcherHandle() { plvi_shim::check(plvi_matcher_create(&h_, 1, 16384, 16384, 0, nullptr), "matcher"); }
  ~PlviMatcherHandle() { plvi_matcher_destroy(h_); }
  plvi_matcher* h_ = nullptr;
};

class ORBmatcher {
 public:
  static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;
  ORBmatcher(float nnratio = 0.6f, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

  // static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b)  (src/ORBmatcher.cc:2350-2366)
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    int d = 0;
    plvi_shim::check(plvi_hamming256(PlviMatcherHandle::get(), a.data, b.data, 1, 0, &d, 0), "DescriptorDistance");
    return d;
  }

  // Descriptor part of SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono):
  // the caller projects LastFrame's map points (src/ORBmatcher.cc:1992-2023) into plvi_query records.
  // blocked (optional, one byte per current key): the key already holds a map point with Observations() > 0
  // (src/ORBmatcher.cc:87-89, 2037-2039), or, for SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th, ratioHamming)
  // (:473-704, which claims features while it iterates: mode FRAME, mbCheckOrientation = false), vpMatched[i] != NULL;
  // thDist < 0: TH_HIGH (TH_LOW for INIT), else e.g. ORBdist of the relocalisation overload or TH_LOW * ratioHamming.
  int SearchByProjection(const std::vector<cv::KeyPoint>& curKeysUn, const cv::Mat& curDesc, const plvi_grid& grid,
                         std::vector<plvi_query>& queries, const cv::Mat& queryDesc, std::vector<int>& matchOfCurKey,
                         int mode = PLVI_SEARCH_FRAME, const std::vector<uint8_t>* blocked = nullptr, int thDist = -1) {
    const int n = (int)curKeysUn.size(), nq = (int)queries.size();
    matchOfCurKey.assign(n, -1);
    std::vector<int> mq(nq > 0 ? nq : 1);
    int nm = 0;
    if (n == 0 || nq == 0) return 0;
    std::vector<uint8_t> d((size_t)n * 32), qd((size_t)nq * 32);
    for (int i = 0; i < n; i++) std::memcpy(&d[(size_t)i * 32], curDesc.ptr(i), 32);
    for (int i = 0; i < nq; i++) std::memcpy(&qd[(size_t)i * 32], queryDesc.ptr(i), 32);
    plvi_shim::check(plvi_search_by_projection(PlviMatcherHandle::get(), mode, 1,
                                               reinterpret_cast<const plvi_keypoint*>(curKeysUn.data()), d.data(),
                                               (blocked && (int)blocked->size() >= n) ? blocked->data() : nullptr, &n, n,
                                               &grid, queries.data(), qd.data(), &nq, nq,
                                               thDist >= 0 ? thDist : (mode == PLVI_SEARCH_INIT ? TH_LOW : TH_HIGH),
                                               mfNNratio, mbCheckOrientation ? 1 : 0, matchOfCurKey.data(), mq.data(), &nm, 0),
                     "SearchByProjection");
    return nm;
  }

 protected:
  float mfNNratio;
  bool mbCheckOrientation;
};

class LineMatcher {
 public:
  static const int TH_HIGH = 100, TH_LOW = 50;

  // static int match(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12)
  // (src/LineMatcher.cpp:92-111).  Like the reference, matches_12 is resize()d, not reset: entries that
  // exist on entry and get no new match keep their value before the mutual check.
  static int match(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12) {
    return run(desc1, desc2, nnr, matches_12, 1);
  }
  static int matchNNR(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12) {
    return run(desc1, desc2, nnr, matches_12, 0);
  }
  static int distance(const cv::Mat& a, const cv::Mat& b) {
    int d = 0;
    plvi_shim::check(plvi_hamming256(PlviMatcherHandle::get(), a.data, b.data, 1, 0, &d, 0), "LineMatcher::distance");
    return d;
  }
  // the >>25 variant of the reference (src/LineMatcher.cpp:487-499), kept bit-compatible
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    int d = 0;
    plvi_shim::check(plvi_hamming256(PlviMatcherHandle::get(), a.data, b.data, 1, 1, &d, 0), "LineMatcher::DescriptorDistance");
    return d;
  }

  // The line search of Frame::ComputeStereoMatches_Lines (src/Frame.cc:1421-1448): takes the place of its grid fill
  // (GridStructure(FRAME_GRID_ROWS, FRAME_GRID_COLS), getLineCoords per right keyline, the directions table) and of
  // the call LineMatcher::matchGrid(coords, mDescriptors_Line, grid, mDescriptorsRight_Line, directions, w, matches_12)
  // (include/LineMatcher.h:101, src/LineMatcher.cpp:191-272) with the window of that call site (7, 0) x (2, 2).
  // The depth / disparity filter that follows stays in Frame.
  static int matchGrid(const std::vector<cv::line_descriptor::KeyLine>& linesLeft, const cv::Mat& desc1,
                       const std::vector<cv::line_descriptor::KeyLine>& linesRight, const cv::Mat& desc2, double inv_width,
                       double inv_height, std::vector<int>& matches_12, int gridRows = 48, int gridCols = 64) {
    const int n1 = (int)linesLeft.size(), n2 = (int)linesRight.size();
    if (n1 != desc1.rows) throw std::runtime_error("[matchGrid] Each line needs a corresponding descriptor!");
    matches_12.resize(n1, -1);
    if (n1 == 0) return 0;
    std::vector<float> s1((size_t)n1 * 4), s2((size_t)(n2 > 0 ? n2 : 1) * 4);
    std::vector<uint8_t> a((size_t)n1 * 32), b((size_t)(n2 > 0 ? n2 : 1) * 32);
    for (int i = 0; i < n1; i++) {
      const cv::line_descriptor::KeyLine& k = linesLeft[i];
      s1[4 * i] = k.startPointX; s1[4 * i + 1] = k.startPointY; s1[4 * i + 2] = k.endPointX; s1[4 * i + 3] = k.endPointY;
      std::memcpy(&a[(size_t)i * 32], desc1.ptr(i), 32);
    }
    for (int i = 0; i < n2; i++) {
      const cv::line_descriptor::KeyLine& k = linesRight[i];
      s2[4 * i] = k.startPointX; s2[4 * i + 1] = k.startPointY; s2[4 * i + 2] = k.endPointX; s2[4 * i + 3] = k.endPointY;
      std::memcpy(&b[(size_t)i * 32], desc2.ptr(i), 32);
    }
    std::vector<int> fresh(n1, -1);
    int nm = 0;
    plvi_shim::check(plvi_line_match_grid_host(PlviMatcherHandle::get(), s1.data(), a.data(), n1, s2.data(), b.data(), n2,
                                               inv_width, inv_height, gridRows, gridCols, 7, 0, 2, 2, fresh.data(), &nm),
                     "LineMatcher::matchGrid");
    // like the reference, matches_12 is resize()d, not reset: the mutual check rewrites every entry it visits
    for (int i = 0; i < n1; i++) matches_12[i] = fresh[i];
    return nm;
  }

 private:
  static int run(const cv::Mat& desc1, const cv::Mat& desc2, float nnr, std::vector<int>& matches_12, int mutual) {
    const int n1 = desc1.rows, n2 = desc2.rows;
    matches_12.resize(n1, -1);
    if (n1 == 0) return 0;
    std::vector<uint8_t> a((size_t)n1 * 32), b((size_t)(n2 > 0 ? n2 : 1) * 32);
    for (int i = 0; i < n1; i++) std::memcpy(&a[(size_t)i * 32], desc1.ptr(i), 32);
    for (int i = 0; i < n2; i++) std::memcpy(&b[(size_t)i * 32], desc2.ptr(i), 32);
    std::vector<int> fresh(n1, -1);
    int nm = 0;
    plvi_shim::check(plvi_line_match(PlviMatcherHandle::get(), 1, a.data(), &n1, n1, b.data(), &n2, n2 > 0 ? n2 : 1, nnr, mutual,
                                     fresh.data(), &nm, 0), "LineMatcher::match");
    for (int i = 0; i < n1; i++)
      if (fresh[i] >= 0 || mutual) matches_12[i] = fresh[i];
    return nm;
  }
};

}  // namespace ORB_SLAM3
