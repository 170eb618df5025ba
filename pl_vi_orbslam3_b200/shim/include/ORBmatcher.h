// Drop-in for the reference's include/ORBmatcher.h:33-101.  Every search keeps its reference signature; the host
// side (shim/src/ORBmatcher.cc) walks Frame / KeyFrame / MapPoint exactly as the reference does up to the point where
// a candidate has been projected, hands the Hamming search to libplvi_cuda.so (include/plvi.h) and writes the result
// back into mvpMapPoints / the output vectors.  DescriptorDistance stays a host-inline popcount: the reference calls it
// 10^4..10^5 times per frame from MapPoint.cc:378, MapLine.cc:305 and Frame.cc:1303.
#ifndef ORBMATCHER_H
#define ORBMATCHER_H
#include <set>
#include <utility>
#include <vector>

#include "plvi_cv_compat.h"

#ifdef PLVI_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>

#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"
#endif

namespace ORB_SLAM3 {

class ORBmatcher {
 public:
  ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

  // static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b)   (include/ORBmatcher.h:43, src/ORBmatcher.cc:2350-2366)
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return plvi_inline_hamming256(a.ptr(0), b.ptr(0), 0); }
  // the batched device form of the same function: n descriptor pairs (row i of a against row i of b) in one launch
  static void DescriptorDistanceBatch(const cv::Mat& a, const cv::Mat& b, std::vector<int>& out) {
    out.resize(a.rows);
    std::vector<uint8_t> ta, tb;
    if (a.rows > 0)
      plvi_shim::check(plvi_hamming256(plvi_shim::MatcherHandle::get(), plvi_shim::packed_rows(a, a.rows, ta),
                                       plvi_shim::packed_rows(b, a.rows, tb), a.rows, 0, out.data(), 0), "DescriptorDistanceBatch");
  }

#ifdef PLVI_HAVE_OPENCV
  // include/ORBmatcher.h:47-81 -- same signatures, same results (src/ORBmatcher.cc)
  int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                         const float thFarPoints = 50.0f);
  int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
  int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);
  int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th,
                         float ratioHamming = 1.0);
  int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, const std::vector<KeyFrame*>& vpPointsKFs,
                         std::vector<MapPoint*>& vpMatched, std::vector<KeyFrame*>& vpMatchedKF, int th, float ratioHamming = 1.0);
  int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
  int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
  int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                              int windowSize = 10);
  int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                             const bool bOnlyStereo, const bool bCoarse = false);
  int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                             const bool bOnlyStereo, std::vector<cv::Mat>& vMatchedPoints);
  int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                   const cv::Mat& t12, const float th);
  int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0, const bool bRight = false);
  int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint);
#endif

  // Descriptor-level form of the guided searches for callers that hold plain arrays (no Frame objects): the train side
  // is mvKeysUn + mDescriptors + the grid parameters, one plvi_query per projected candidate in the reference's
  // iteration order; matchOfCurKey[i] = query assigned to train keypoint i or -1.  See include/plvi.h.
  int SearchByProjectionRaw(const std::vector<cv::KeyPoint>& curKeysUn, const cv::Mat& curDesc, const plvi_grid& grid,
                            std::vector<plvi_query>& queries, const cv::Mat& queryDesc, std::vector<int>& matchOfCurKey,
                            int mode = PLVI_SEARCH_FRAME, const std::vector<uint8_t>* blocked = nullptr, int thDist = -1) {
    const int n = (int)curKeysUn.size(), nq = (int)queries.size();
    matchOfCurKey.assign(n, -1);
    if (n == 0 || nq == 0) return 0;
    std::vector<int> mq(nq);
    std::vector<uint8_t> td, tq;
    int nm = 0;
    plvi_shim::check(plvi_search_by_projection(plvi_shim::MatcherHandle::get(), mode, 1,
                                               reinterpret_cast<const plvi_keypoint*>(curKeysUn.data()),
                                               plvi_shim::packed_rows(curDesc, n, td),
                                               (blocked && (int)blocked->size() >= n) ? blocked->data() : nullptr, &n, n, &grid,
                                               queries.data(), plvi_shim::packed_rows(queryDesc, nq, tq), &nq, nq,
                                               thDist >= 0 ? thDist : (mode == PLVI_SEARCH_INIT ? 50 : 100), mfNNratio,
                                               mbCheckOrientation ? 1 : 0, matchOfCurKey.data(), mq.data(), &nm, 0),
                     "SearchByProjection");
    return nm;
  }

 public:
  static const int TH_LOW;         // 50   (src/ORBmatcher.cc:36-38; defined in shim/src/ORBmatcher.cc)
  static const int TH_HIGH;        // 100
  static const int HISTO_LENGTH;   // 30

 protected:
  float RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }   // src/ORBmatcher.cc:216-222

  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace ORB_SLAM3

#endif  // ORBMATCHER_H
