// Reference-signature ORBmatcher on top of the C ABI (include/plvi.h).  Takes the place of the reference's
// src/ORBmatcher.cc in a SLAM build: every function walks Frame / KeyFrame / MapPoint as the
// reference does up to the projected candidate (same cv::Mat expressions, so the float results are those of the
// reference), batches the candidates into plvi_query records in the reference's iteration order, runs the Hamming
// search on the GPU and applies the bookkeeping (mvpMapPoints, Replace / AddObservation, output vectors) in that order.
//
// Supported rigs: monocular and rectified stereo / RGB-D (Frame::Nleft == -1).  The two-camera KannalaBrandt8 rig
// (Nleft != -1, mpCamera2, bRight) is rejected with std::runtime_error: its right-camera searches are not built.
#include "ORBmatcher.h"

#include <climits>
#include <cmath>
#include <stdexcept>

using namespace std;

namespace ORB_SLAM3 {

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

namespace {

using plvi_shim::check;
using plvi_shim::MatcherHandle;
using plvi_shim::packed_rows;

void single_camera_only(bool twoCameras, const char* where) {
  if (twoCameras) throw std::runtime_error(std::string(where) + ": the two-camera KannalaBrandt8 rig (Nleft != -1) is not supported by libplvi_cuda");
}

template <class FrameLike> plvi_grid grid_of(const FrameLike& f) {
  plvi_grid g;
  g.min_x = (float)f.mnMinX; g.min_y = (float)f.mnMinY;
  g.inv_w = f.mfGridElementWidthInv; g.inv_h = f.mfGridElementHeightInv;
  return g;
}

const plvi_keypoint* as_plvi(const std::vector<cv::KeyPoint>& v) { return reinterpret_cast<const plvi_keypoint*>(v.data()); }

// The projected candidates of one search, in the reference's iteration order.
struct Candidates {
  std::vector<plvi_query> q;
  std::vector<uint8_t> desc;   // 32 bytes per query
  std::vector<int> src;        // index of the candidate in the caller's list
  std::vector<float> ur;       // right-image coordinate of the projection (rectified stereo)
  int size() const { return (int)q.size(); }
  void add(int srcIdx, float u, float v, float radius, int minLevel, int maxLevel, float angle, int flags, const cv::Mat& d, float uRight = 0.f) {
    plvi_query e;
    e.u = u; e.v = v; e.radius = radius; e.min_level = minLevel; e.max_level = maxLevel; e.angle = angle; e.flags = flags;
    q.push_back(e);
    const size_t o = desc.size();
    desc.resize(o + 32);
    std::memcpy(&desc[o], d.ptr(0), 32);
    src.push_back(srcIdx);
    ur.push_back(uRight);
  }
};

// does the frame hold rectified-stereo observations?  The searches test mvuRight > 0 (src/ORBmatcher.cc:91, 2041), the
// reprojection gate of Fuse mvuRight >= 0 (:1530)
bool any_stereo(const std::vector<float>& uRight, int n, bool zeroCounts) {
  for (int i = 0; i < n && i < (int)uRight.size(); i++)
    if (uRight[i] > 0.f || (zeroCounts && uRight[i] >= 0.f)) return true;
  return false;
}

// plvi_search_by_projection on one frame / keyframe with host arrays
int guided_search(int mode, const std::vector<cv::KeyPoint>& keys, const cv::Mat& descriptors, int n, const std::vector<uint8_t>& blocked,
                  const plvi_grid& grid, Candidates& c, int thDist, float nnratio, int checkOri, const std::vector<float>* uRight,
                  std::vector<int>& matchTrain, std::vector<int>& matchQuery) {
  matchTrain.assign(n > 0 ? n : 1, -1);
  const int nq = c.size();
  matchQuery.assign(nq > 0 ? nq : 1, -1);
  if (n <= 0 || nq == 0) return 0;
  std::vector<uint8_t> tmp;
  plvi_matcher* m = MatcherHandle::get();
  if (uRight) check(plvi_matcher_set_stereo(m, uRight->data(), c.ur.data(), 1, n, nq, 0), "plvi_matcher_set_stereo");
  int nm = 0;
  check(plvi_search_by_projection(m, mode, 1, as_plvi(keys), packed_rows(descriptors, n, tmp), blocked.empty() ? nullptr : blocked.data(), &n,
                                  n, &grid, c.q.data(), c.desc.data(), &nq, nq, thDist, nnratio, checkOri, matchTrain.data(),
                                  matchQuery.data(), &nm, 0),
        "plvi_search_by_projection");
  return nm;
}

// plvi_search_in_radius on one keyframe with host arrays: best_idx per candidate (-1: none within thDist)
void radius_search(KeyFrame* pKF, Candidates& c, double chi2, int thDist, bool stereoGate, std::vector<int>& bestIdx) {
  const int n = pKF->N, nq = c.size();
  bestIdx.assign(nq > 0 ? nq : 1, -1);
  if (n <= 0 || nq == 0) return;
  std::vector<int> bestDist(nq);
  std::vector<uint8_t> tmp;
  float invSigma2[16] = {0};
  for (size_t i = 0; i < pKF->mvInvLevelSigma2.size() && i < 16; i++) invSigma2[i] = pKF->mvInvLevelSigma2[i];
  const plvi_grid grid = grid_of(*pKF);
  int nfound = 0;
  const bool stereo = stereoGate && any_stereo(pKF->mvuRight, n, true);
  check(plvi_search_in_radius_host(MatcherHandle::get(), as_plvi(pKF->mvKeysUn), packed_rows(pKF->mDescriptors, n, tmp), n, &grid, c.q.data(),
                                   c.desc.data(), nq, invSigma2, chi2, thDist, stereo ? pKF->mvuRight.data() : nullptr,
                                   stereo ? c.ur.data() : nullptr, bestIdx.data(), bestDist.data(), &nfound),
        "plvi_search_in_radius");
}

// Scw = [s R | s t]  ->  R, t, camera centre   (src/ORBmatcher.cc:482-487, 598-602, 1620-1625)
struct SimilarityPose {
  cv::Mat Rcw, tcw, Ow;
  explicit SimilarityPose(const cv::Mat& Scw) {
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    Rcw = sRcw / scw;
    tcw = Scw.rowRange(0, 3).col(3) / scw;
    Ow = -Rcw.t() * tcw;
  }
};

// The checks every keyframe projection search runs on a map point before it looks for features (positive depth, inside
// the image, distance within the scale-invariance range, viewing angle below 60 degrees, predicted level, radius):
// src/ORBmatcher.cc:507-545 (= 625-667), 1455-1508, 1643-1683.  viaCamera selects mpCamera->project(Point3f) against
// the explicit fx * (x / z) + cx of the vpPointsKFs overload.
struct KeyFrameProjection { float u, v, ur; int level; float radius; };
bool project_into_keyframe(KeyFrame* pKF, GeometricCamera* pCamera, MapPoint* pMP, const cv::Mat& Rcw, const cv::Mat& tcw, const cv::Mat& Ow,
                           bool viaCamera, float th, KeyFrameProjection& out) {
  cv::Mat p3Dw = pMP->GetWorldPos();
  cv::Mat p3Dc = Rcw * p3Dw + tcw;
  if (p3Dc.at<float>(2) < 0.0f) return false;
  const float invz = 1 / p3Dc.at<float>(2);
  if (viaCamera) {
    const cv::Point2f uv = pCamera->project(cv::Point3f(p3Dc.at<float>(0), p3Dc.at<float>(1), p3Dc.at<float>(2)));
    out.u = uv.x; out.v = uv.y;
  } else {
    const float x = p3Dc.at<float>(0) * invz;
    const float y = p3Dc.at<float>(1) * invz;
    out.u = pKF->fx * x + pKF->cx;
    out.v = pKF->fy * y + pKF->cy;
  }
  if (!pKF->IsInImage(out.u, out.v)) return false;
  out.ur = out.u - pKF->mbf * invz;
  const float maxDistance = pMP->GetMaxDistanceInvariance();
  const float minDistance = pMP->GetMinDistanceInvariance();
  cv::Mat PO = p3Dw - Ow;
  const float dist3D = cv::norm(PO);
  if (dist3D < minDistance || dist3D > maxDistance) return false;
  cv::Mat Pn = pMP->GetNormal();
  if (PO.dot(Pn) < 0.5 * dist3D) return false;
  out.level = pMP->PredictScale(dist3D, pKF);
  out.radius = th * pKF->mvScaleFactors[out.level];
  return true;
}

// Walk of two DBoW2 feature vectors over their common nodes (src/ORBmatcher.cc:286-292,435-448): calls
// visit(indices1, indices2) for every shared node, in ascending node order.
template <class Visit> void common_nodes(const DBoW2::FeatureVector& fv1, const DBoW2::FeatureVector& fv2, Visit visit) {
  DBoW2::FeatureVector::const_iterator it1 = fv1.begin(), it2 = fv2.begin();
  while (it1 != fv1.end() && it2 != fv2.end()) {
    if (it1->first == it2->first) {
      visit(it1->second, it2->second);
      ++it1; ++it2;
    } else if (it1->first < it2->first) {
      it1 = fv1.lower_bound(it2->first);
    } else {
      it2 = fv2.lower_bound(it1->first);
    }
  }
}

}  // namespace

// src/ORBmatcher.cc:44-214.  Mode MAPPOINTS of the guided search: best / second best with the level-aware ratio test.
int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th, const bool bFarPoints, const float thFarPoints) {
  single_camera_only(F.Nleft != -1, "ORBmatcher::SearchByProjection(Frame, MapPoints)");
  const bool bFactor = th != 1.0;
  Candidates c;
  for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
    MapPoint* pMP = vpMapPoints[iMP];
    if (!pMP->mbTrackInView && !pMP->mbTrackInViewR) continue;
    if (bFarPoints && pMP->mTrackDepth > thFarPoints) continue;
    if (pMP->isBad()) continue;
    if (!pMP->mbTrackInView) continue;
    const int level = pMP->mnTrackScaleLevel;
    float r = RadiusByViewingCos(pMP->mTrackViewCos);   // window size depends on the viewing direction
    if (bFactor) r *= th;
    // a feature that receives this point blocks later candidates only if the point has observations (:87-89)
    c.add((int)iMP, pMP->mTrackProjX, pMP->mTrackProjY, r * F.mvScaleFactors[level], level - 1, level, 0.f,
          pMP->Observations() > 0 ? 0 : 2, pMP->GetDescriptor(), pMP->mTrackProjXR);
  }
  const int n = F.N;
  std::vector<uint8_t> blocked(n > 0 ? n : 1, 0);
  for (int i = 0; i < n; i++) blocked[i] = (F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0) ? 1 : 0;
  const bool stereo = any_stereo(F.mvuRight, n, false);
  std::vector<int> mt, mq;
  const int nmatches = guided_search(PLVI_SEARCH_MAPPOINTS, F.mvKeysUn, F.mDescriptors, n, blocked, grid_of(F), c, TH_HIGH, mfNNratio, 0,
                                     stereo ? &F.mvuRight : nullptr, mt, mq);
  for (int i = 0; i < n; i++)
    if (mt[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[c.src[mt[i]]];
  return nmatches;
}

// src/ORBmatcher.cc:1962-2178.  Mode FRAME: features are claimed in the order of LastFrame's points; rotation histogram.
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
  single_camera_only(CurrentFrame.Nleft != -1 || LastFrame.Nleft != -1, "ORBmatcher::SearchByProjection(Frame, Frame)");
  const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
  const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
  const cv::Mat twc = -Rcw.t() * tcw;
  const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
  const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
  const cv::Mat tlc = Rlw * twc + tlw;
  // stereo / RGB-D only: the camera moved along its axis by more than the baseline, so the scale changes one way
  const bool bForward = tlc.at<float>(2) > CurrentFrame.mb && !bMono;
  const bool bBackward = -tlc.at<float>(2) > CurrentFrame.mb && !bMono;

  Candidates c;
  for (int i = 0; i < LastFrame.N; i++) {
    MapPoint* pMP = LastFrame.mvpMapPoints[i];
    if (!pMP || LastFrame.mvbOutlier[i]) continue;
    cv::Mat x3Dw = pMP->GetWorldPos();
    cv::Mat x3Dc = Rcw * x3Dw + tcw;
    const float invzc = 1.0 / x3Dc.at<float>(2);
    if (invzc < 0) continue;
    cv::Point2f uv = CurrentFrame.mpCamera->project(x3Dc);
    if (uv.x < CurrentFrame.mnMinX || uv.x > CurrentFrame.mnMaxX) continue;
    if (uv.y < CurrentFrame.mnMinY || uv.y > CurrentFrame.mnMaxY) continue;
    const int nLastOctave = LastFrame.mvKeys[i].octave;
    const float radius = th * CurrentFrame.mvScaleFactors[nLastOctave];
    int minLevel = nLastOctave - 1, maxLevel = nLastOctave + 1;
    if (bForward) { minLevel = nLastOctave; maxLevel = -1; }
    else if (bBackward) { minLevel = 0; maxLevel = nLastOctave; }
    const float ur = uv.x - CurrentFrame.mbf * invzc;
    c.add(i, uv.x, uv.y, radius, minLevel, maxLevel, LastFrame.mvKeysUn[i].angle, pMP->Observations() > 0 ? 0 : 2, pMP->GetDescriptor(), ur);
  }
  const int n = CurrentFrame.N;
  std::vector<uint8_t> blocked(n > 0 ? n : 1, 0);
  for (int i = 0; i < n; i++)
    blocked[i] = (CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0) ? 1 : 0;
  const bool stereo = any_stereo(CurrentFrame.mvuRight, n, false);
  std::vector<int> mt, mq;
  const int nmatches = guided_search(PLVI_SEARCH_FRAME, CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors, n, blocked, grid_of(CurrentFrame), c,
                                     TH_HIGH, mfNNratio, mbCheckOrientation ? 2 : 0, stereo ? &CurrentFrame.mvuRight : nullptr, mt, mq);
  for (int i = 0; i < n; i++) {
    if (mt[i] >= 0) CurrentFrame.mvpMapPoints[i] = LastFrame.mvpMapPoints[c.src[mt[i]]];
    else if (mt[i] == -2) CurrentFrame.mvpMapPoints[i] = static_cast<MapPoint*>(NULL);   // assigned, then removed by the rotation histogram
  }
  return nmatches;
}

// src/ORBmatcher.cc:2180-2302 (relocalisation): any feature that already holds a point is skipped.
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) {
  single_camera_only(CurrentFrame.Nleft != -1, "ORBmatcher::SearchByProjection(Frame, KeyFrame)");
  const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
  const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
  const cv::Mat Ow = -Rcw.t() * tcw;
  const vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();

  Candidates c;
  for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {
    MapPoint* pMP = vpMPs[i];
    if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
    cv::Mat x3Dw = pMP->GetWorldPos();
    cv::Mat x3Dc = Rcw * x3Dw + tcw;
    const cv::Point2f uv = CurrentFrame.mpCamera->project(x3Dc);
    if (uv.x < CurrentFrame.mnMinX || uv.x > CurrentFrame.mnMaxX) continue;
    if (uv.y < CurrentFrame.mnMinY || uv.y > CurrentFrame.mnMaxY) continue;
    cv::Mat PO = x3Dw - Ow;
    float dist3D = cv::norm(PO);
    const float maxDistance = pMP->GetMaxDistanceInvariance();
    const float minDistance = pMP->GetMinDistanceInvariance();
    if (dist3D < minDistance || dist3D > maxDistance) continue;
    const int level = pMP->PredictScale(dist3D, &CurrentFrame);
    c.add((int)i, uv.x, uv.y, th * CurrentFrame.mvScaleFactors[level], level - 1, level + 1, pKF->mvKeysUn[i].angle, 0, pMP->GetDescriptor());
  }
  const int n = CurrentFrame.N;
  std::vector<uint8_t> blocked(n > 0 ? n : 1, 0);
  for (int i = 0; i < n; i++) blocked[i] = CurrentFrame.mvpMapPoints[i] ? 1 : 0;
  std::vector<int> mt, mq;
  const int nmatches = guided_search(PLVI_SEARCH_FRAME, CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors, n, blocked, grid_of(CurrentFrame), c,
                                     ORBdist, mfNNratio, mbCheckOrientation ? 2 : 0, nullptr, mt, mq);
  for (int i = 0; i < n; i++) {
    if (mt[i] >= 0) CurrentFrame.mvpMapPoints[i] = vpMPs[c.src[mt[i]]];
    else if (mt[i] == -2) CurrentFrame.mvpMapPoints[i] = NULL;
  }
  return nmatches;
}

// src/ORBmatcher.cc:473-596 (loop detection): features are claimed while the points are visited, no rotation check.
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th,
                                   float ratioHamming) {
  const SimilarityPose pose(Scw);
  set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
  spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
  Candidates c;
  for (int iMP = 0, iendMP = vpPoints.size(); iMP < iendMP; iMP++) {
    MapPoint* pMP = vpPoints[iMP];
    if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
    KeyFrameProjection p;
    if (!project_into_keyframe(pKF, pKF->mpCamera, pMP, pose.Rcw, pose.tcw, pose.Ow, true, (float)th, p)) continue;
    c.add(iMP, p.u, p.v, p.radius, p.level - 1, p.level, 0.f, 0, pMP->GetDescriptor());
  }
  const int n = pKF->N;
  std::vector<uint8_t> blocked(n > 0 ? n : 1, 0);
  for (int i = 0; i < n; i++) blocked[i] = vpMatched[i] ? 1 : 0;
  std::vector<int> mt, mq;
  const int nmatches = guided_search(PLVI_SEARCH_FRAME, pKF->mvKeysUn, pKF->mDescriptors, n, blocked, grid_of(*pKF), c,
                                     (int)std::floor(TH_LOW * ratioHamming), mfNNratio, 0, nullptr, mt, mq);
  for (int i = 0; i < n; i++)
    if (mt[i] >= 0) vpMatched[i] = vpPoints[c.src[mt[i]]];
  return nmatches;
}

// src/ORBmatcher.cc:588-704 (place recognition): as above, also reports the keyframe every matched point came from.
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, const vector<KeyFrame*>& vpPointsKFs,
                                   vector<MapPoint*>& vpMatched, vector<KeyFrame*>& vpMatchedKF, int th, float ratioHamming) {
  const SimilarityPose pose(Scw);
  set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
  spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
  Candidates c;
  for (int iMP = 0, iendMP = vpPoints.size(); iMP < iendMP; iMP++) {
    MapPoint* pMP = vpPoints[iMP];
    if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
    KeyFrameProjection p;
    if (!project_into_keyframe(pKF, pKF->mpCamera, pMP, pose.Rcw, pose.tcw, pose.Ow, false, (float)th, p)) continue;
    c.add(iMP, p.u, p.v, p.radius, p.level - 1, p.level, 0.f, 0, pMP->GetDescriptor());
  }
  const int n = pKF->N;
  std::vector<uint8_t> blocked(n > 0 ? n : 1, 0);
  for (int i = 0; i < n; i++) blocked[i] = vpMatched[i] ? 1 : 0;
  std::vector<int> mt, mq;
  const int nmatches = guided_search(PLVI_SEARCH_FRAME, pKF->mvKeysUn, pKF->mDescriptors, n, blocked, grid_of(*pKF), c,
                                     (int)std::floor(TH_LOW * ratioHamming), mfNNratio, 0, nullptr, mt, mq);
  for (int i = 0; i < n; i++)
    if (mt[i] >= 0) {
      vpMatched[i] = vpPoints[c.src[mt[i]]];
      vpMatchedKF[i] = vpPointsKFs[c.src[mt[i]]];
    }
  return nmatches;
}

// src/ORBmatcher.cc:269-471: brute force inside the common vocabulary nodes, TH_LOW, ratio test, rotation histogram.
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) {
  single_camera_only(F.Nleft != -1 || pKF->mpCamera2, "ORBmatcher::SearchByBoW(KeyFrame, Frame)");
  const vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
  vpMapPointMatches = vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
  Candidates c;
  std::vector<int> items;   // the frame's features grouped by node, each group in vIndicesF order
  common_nodes(pKF->mFeatVec, F.mFeatVec, [&](const std::vector<unsigned int>& vIndicesKF, const std::vector<unsigned int>& vIndicesF) {
    const int start = (int)items.size();
    items.insert(items.end(), vIndicesF.begin(), vIndicesF.end());
    const int end = (int)items.size();
    for (size_t iKF = 0; iKF < vIndicesKF.size(); iKF++) {
      const unsigned int realIdxKF = vIndicesKF[iKF];
      MapPoint* pMP = vpMapPointsKF[realIdxKF];
      if (!pMP || pMP->isBad()) continue;
      c.add((int)realIdxKF, 0.f, 0.f, 0.f, start, end, pKF->mvKeysUn[realIdxKF].angle, 0, pKF->mDescriptors.row(realIdxKF));
    }
  });
  const int n = F.N, nq = c.size(), ni = (int)items.size();
  if (n <= 0 || nq == 0 || ni == 0) return 0;
  std::vector<int> mt(n, -1), mq(nq, -1);
  std::vector<uint8_t> tmp;
  int nmatches = 0;
  check(plvi_search_by_bow(MatcherHandle::get(), 1, as_plvi(F.mvKeys), packed_rows(F.mDescriptors, n, tmp), &n, n, items.data(), ni, c.q.data(),
                           c.desc.data(), &nq, nq, TH_LOW, mfNNratio, mbCheckOrientation ? 1 : 0, mt.data(), mq.data(), &nmatches, 0),
        "plvi_search_by_bow");
  for (int i = 0; i < n; i++)
    if (mt[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[c.src[mt[i]]];
  return nmatches;
}

// src/ORBmatcher.cc:823-963: keyframe against keyframe, only features that hold good map points on both sides.
int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) {
  single_camera_only(pKF1->NLeft != -1 || pKF2->NLeft != -1, "ORBmatcher::SearchByBoW(KeyFrame, KeyFrame)");
  const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
  const vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
  vpMatches12 = vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
  Candidates c;
  std::vector<int> items;
  common_nodes(pKF1->mFeatVec, pKF2->mFeatVec, [&](const std::vector<unsigned int>& ind1, const std::vector<unsigned int>& ind2) {
    const int start = (int)items.size();
    items.insert(items.end(), ind2.begin(), ind2.end());
    const int end = (int)items.size();
    for (size_t i1 = 0; i1 < ind1.size(); i1++) {
      const size_t idx1 = ind1[i1];
      MapPoint* pMP1 = vpMapPoints1[idx1];
      if (!pMP1 || pMP1->isBad()) continue;
      c.add((int)idx1, 0.f, 0.f, 0.f, start, end, pKF1->mvKeysUn[idx1].angle, 0, pKF1->mDescriptors.row(idx1));
    }
  });
  const int n = (int)vpMapPoints2.size(), nq = c.size(), ni = (int)items.size();
  if (n <= 0 || nq == 0 || ni == 0) return 0;
  std::vector<uint8_t> blocked(n, 0), tmp;
  for (int i = 0; i < n; i++) blocked[i] = (!vpMapPoints2[i] || vpMapPoints2[i]->isBad()) ? 1 : 0;
  std::vector<int> mt(n, -1), mq(nq, -1);
  int nmatches = 0;
  check(plvi_search_by_bow_kf(MatcherHandle::get(), 1, as_plvi(pKF2->mvKeysUn), packed_rows(pKF2->mDescriptors, n, tmp), blocked.data(), &n, n,
                              items.data(), ni, c.q.data(), c.desc.data(), &nq, nq, TH_LOW, mfNNratio, mbCheckOrientation ? 1 : 0, mt.data(),
                              mq.data(), &nmatches, 0),
        "plvi_search_by_bow_kf");
  for (int q = 0; q < nq; q++)
    if (mq[q] >= 0) vpMatches12[c.src[q]] = vpMapPoints2[mq[q]];
  return nmatches;
}

// src/ORBmatcher.cc:706-820: level-0 features of F1 against a window around vbPrevMatched in F2; a better match steals
// the feature of a worse one.
int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12, int windowSize) {
  const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
  vnMatches12 = vector<int>(n1, -1);
  Candidates c;
  for (int i1 = 0; i1 < n1; i1++) {
    const cv::KeyPoint& kp1 = F1.mvKeysUn[i1];
    c.add(i1, vbPrevMatched[i1].x, vbPrevMatched[i1].y, (float)windowSize, kp1.octave, kp1.octave, kp1.angle, kp1.octave > 0 ? 1 : 0,
          F1.mDescriptors.row(i1));
  }
  std::vector<int> mt, mq;
  const int nmatches = guided_search(PLVI_SEARCH_INIT, F2.mvKeysUn, F2.mDescriptors, n2, std::vector<uint8_t>(), grid_of(F2), c, TH_LOW,
                                     mfNNratio, mbCheckOrientation ? 1 : 0, nullptr, mt, mq);
  for (int i1 = 0; i1 < n1; i1++) {
    vnMatches12[i1] = (n2 > 0) ? mq[i1] : -1;
    if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2.mvKeysUn[vnMatches12[i1]].pt;   // "Update prev matched"
  }
  return nmatches;
}

// src/ORBmatcher.cc:965-1206.  Pinhole cameras: the epipolar test is Pinhole::epipolarConstrain
// (src/CameraModels/Pinhole.cpp:135-157), which rebuilds K1^-T [t12]x R12 K2^-1 from R12 = R1w R2w^T and
// t12 = -R1w R2w^T t2w + t1w -- the very expression, on the very operands, that the callers evaluate for the F12 they
// pass in (LocalMapping::ComputeF12, src/LocalMapping.cc:1469-1486; Tracking::ComputeF12, src/Tracking.cc:6195); the
// reference ignores that argument, this implementation uses it.
int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vector<pair<size_t, size_t> >& vMatchedPairs,
                                       const bool bOnlyStereo, const bool bCoarse) {
  single_camera_only(pKF1->mpCamera2 || pKF2->mpCamera2 || pKF1->NLeft != -1 || pKF2->NLeft != -1, "ORBmatcher::SearchForTriangulation");
  // epipole of camera 1 in image 2
  cv::Mat Cw = pKF1->GetCameraCenter();
  cv::Mat R2w = pKF2->GetRotation();
  cv::Mat t2w = pKF2->GetTranslation();
  cv::Mat C2 = R2w * Cw + t2w;
  const cv::Point2f ep = pKF2->mpCamera->project(C2);

  plvi_epipolar geom;
  std::memset(&geom, 0, sizeof(geom));
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 3; k++) geom.F12[3 * r + k] = F12.at<float>(r, k);
  geom.ep_x = ep.x; geom.ep_y = ep.y;
  for (size_t i = 0; i < pKF2->mvScaleFactors.size() && i < 16; i++) geom.scale_factors[i] = pKF2->mvScaleFactors[i];
  for (size_t i = 0; i < pKF2->mvLevelSigma2.size() && i < 16; i++) geom.level_sigma2[i] = pKF2->mvLevelSigma2[i];
  geom.coarse = bCoarse ? 1 : 0;
  geom.check_epipole = 1;

  Candidates c;
  std::vector<int> items;
  common_nodes(pKF1->mFeatVec, pKF2->mFeatVec, [&](const std::vector<unsigned int>& ind1, const std::vector<unsigned int>& ind2) {
    const int start = (int)items.size();
    items.insert(items.end(), ind2.begin(), ind2.end());
    const int end = (int)items.size();
    for (size_t i1 = 0; i1 < ind1.size(); i1++) {
      const size_t idx1 = ind1[i1];
      if (pKF1->GetMapPoint(idx1)) continue;   // only features that are not tracked yet
      const bool bStereo1 = pKF1->mvuRight[idx1] >= 0;
      if (bOnlyStereo && !bStereo1) continue;
      const cv::KeyPoint& kp1 = pKF1->mvKeysUn[idx1];
      c.add((int)idx1, kp1.pt.x, kp1.pt.y, 0.f, start, end, kp1.angle, bStereo1 ? 4 : 0, pKF1->mDescriptors.row(idx1));
    }
  });
  vMatchedPairs.clear();
  const int n = pKF2->N, nq = c.size(), ni = (int)items.size();
  if (n <= 0 || nq == 0 || ni == 0) return 0;
  std::vector<uint8_t> blocked(n, 0), tmp;
  for (int i = 0; i < n; i++) {
    const bool bStereo2 = pKF2->mvuRight[i] >= 0;
    blocked[i] = ((pKF2->GetMapPoint(i) || (bOnlyStereo && !bStereo2)) ? 1 : 0) | (bStereo2 ? 2 : 0);
  }
  std::vector<int> mq(nq, -1);
  int nmatches = 0;
  check(plvi_search_for_triangulation_host(MatcherHandle::get(), as_plvi(pKF2->mvKeysUn), packed_rows(pKF2->mDescriptors, n, tmp), blocked.data(),
                                           n, items.data(), ni, c.q.data(), c.desc.data(), nq, &geom, TH_LOW, mbCheckOrientation ? 1 : 0,
                                           mq.data(), &nmatches),
        "plvi_search_for_triangulation");
  vector<int> vMatches12(pKF1->N, -1);
  for (int q = 0; q < nq; q++) vMatches12[c.src[q]] = mq[q];
  vMatchedPairs.reserve(nmatches);
  for (size_t i = 0, iend = vMatches12.size(); i < iend; i++)
    if (vMatches12[i] >= 0) vMatchedPairs.push_back(make_pair(i, (size_t)vMatches12[i]));
  return nmatches;
}

// src/ORBmatcher.cc:1208-1397: accepts a candidate only when GeometricCamera::matchAndtriangulate succeeds, which the
// pinhole model never does (include/CameraModels/Pinhole.h:95-98 returns false); no call site in the reference.  With
// pinhole cameras the reference therefore returns no pair, and so does this.
int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat /*F12*/, vector<pair<size_t, size_t> >& vMatchedPairs,
                                       const bool /*bOnlyStereo*/, vector<cv::Mat>& /*vMatchedPoints*/) {
  single_camera_only(pKF1->mpCamera2 || pKF2->mpCamera2, "ORBmatcher::SearchForTriangulation(..., vMatchedPoints)");
  vMatchedPairs.clear();
  return 0;
}

// src/ORBmatcher.cc:1736-1960: both directions through the similarity, then the mutual-agreement test.
int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th) {
  const float& fx = pKF1->fx; const float& fy = pKF1->fy; const float& cx = pKF1->cx; const float& cy = pKF1->cy;
  cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation();
  cv::Mat R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
  cv::Mat sR12 = s12 * R12;
  cv::Mat sR21 = (1.0 / s12) * R12.t();
  cv::Mat t21 = -sR21 * t12;
  const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
  const vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
  const int N1 = vpMapPoints1.size(), N2 = vpMapPoints2.size();
  vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
  for (int i = 0; i < N1; i++) {
    MapPoint* pMP = vpMatches12[i];
    if (!pMP) continue;
    vbAlreadyMatched1[i] = true;
    const int idx2 = get<0>(pMP->GetIndexInKeyFrame(pKF2));
    if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
  }
  // one direction: points of `from` through  x_to = sR * (Rw x + tw) + t  into `to`
  auto direction = [&](const vector<MapPoint*>& pts, const vector<bool>& done, const cv::Mat& Rw, const cv::Mat& tw, const cv::Mat& sR,
                       const cv::Mat& t, KeyFrame* to, vector<int>& match) {
    Candidates c;
    for (int i = 0, iend = pts.size(); i < iend; i++) {
      MapPoint* pMP = pts[i];
      if (!pMP || done[i] || pMP->isBad()) continue;
      cv::Mat p3Dw = pMP->GetWorldPos();
      cv::Mat p3Dfrom = Rw * p3Dw + tw;
      cv::Mat p3Dto = sR * p3Dfrom + t;
      if (p3Dto.at<float>(2) < 0.0) continue;
      const float invz = 1.0 / p3Dto.at<float>(2);
      const float x = p3Dto.at<float>(0) * invz;
      const float y = p3Dto.at<float>(1) * invz;
      const float u = fx * x + cx;
      const float v = fy * y + cy;
      if (!to->IsInImage(u, v)) continue;
      const float maxDistance = pMP->GetMaxDistanceInvariance();
      const float minDistance = pMP->GetMinDistanceInvariance();
      const float dist3D = cv::norm(p3Dto);
      if (dist3D < minDistance || dist3D > maxDistance) continue;
      const int level = pMP->PredictScale(dist3D, to);
      c.add(i, u, v, th * to->mvScaleFactors[level], level - 1, level, 0.f, 0, pMP->GetDescriptor());
    }
    std::vector<int> best;
    radius_search(to, c, 0.0, TH_HIGH, false, best);
    match.assign(pts.size(), -1);
    for (int q = 0; q < c.size(); q++) match[c.src[q]] = best[q];
  };
  vector<int> vnMatch1, vnMatch2;
  direction(vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, vnMatch1);
  direction(vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12, pKF1, vnMatch2);
  int nFound = 0;
  for (int i1 = 0; i1 < N1; i1++) {
    const int idx2 = vnMatch1[i1];
    if (idx2 >= 0 && vnMatch2[idx2] == i1) {
      vpMatches12[i1] = vpMapPoints2[idx2];
      nFound++;
    }
  }
  return nFound;
}

// src/ORBmatcher.cc:1399-1610.  The searches of all candidates are independent (nothing is skipped because of an
// earlier hit), so they run as one batch; what follows a hit -- Replace / AddObservation / AddMapPoint -- is applied in
// the reference's order, with the two checks that earlier bookkeeping can change (isBad, IsInKeyFrame) repeated there.
int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th, const bool bRight) {
  single_camera_only(bRight || pKF->NLeft != -1, "ORBmatcher::Fuse");
  cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Ow = pKF->GetCameraCenter();
  Candidates c;
  const int nMPs = vpMapPoints.size();
  for (int i = 0; i < nMPs; i++) {
    MapPoint* pMP = vpMapPoints[i];
    if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
    KeyFrameProjection p;
    if (!project_into_keyframe(pKF, pKF->mpCamera, pMP, Rcw, tcw, Ow, true, th, p)) continue;
    c.add(i, p.u, p.v, p.radius, p.level - 1, p.level, 0.f, 0, pMP->GetDescriptor(), p.ur);
  }
  std::vector<int> best;
  radius_search(pKF, c, 5.99, TH_LOW, true, best);
  int nFused = 0;
  for (int q = 0; q < c.size(); q++) {
    if (best[q] < 0) continue;
    MapPoint* pMP = vpMapPoints[c.src[q]];
    if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
    MapPoint* pMPinKF = pKF->GetMapPoint(best[q]);
    if (pMPinKF) {
      if (!pMPinKF->isBad()) {
        if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
        else pMPinKF->Replace(pMP);
      }
    } else {
      pMP->AddObservation(pKF, best[q]);
      pKF->AddMapPoint(pMP, best[q]);
    }
    nFused++;
  }
  return nFused;
}

// src/ORBmatcher.cc:1612-1734 (loop correction): no reprojection gate; a feature that already holds a good point is
// reported in vpReplacePoint instead of being replaced here.
int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint) {
  const SimilarityPose pose(Scw);
  const set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
  Candidates c;
  const int nPoints = vpPoints.size();
  for (int iMP = 0; iMP < nPoints; iMP++) {
    MapPoint* pMP = vpPoints[iMP];
    if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
    KeyFrameProjection p;
    if (!project_into_keyframe(pKF, pKF->mpCamera, pMP, pose.Rcw, pose.tcw, pose.Ow, true, th, p)) continue;
    c.add(iMP, p.u, p.v, p.radius, p.level - 1, p.level, 0.f, 0, pMP->GetDescriptor());
  }
  std::vector<int> best;
  radius_search(pKF, c, 0.0, TH_LOW, false, best);
  int nFused = 0;
  for (int q = 0; q < c.size(); q++) {
    if (best[q] < 0) continue;
    const int iMP = c.src[q];
    MapPoint* pMPinKF = pKF->GetMapPoint(best[q]);
    if (pMPinKF) {
      if (!pMPinKF->isBad()) vpReplacePoint[iMP] = pMPinKF;
    } else {
      vpPoints[iMP]->AddObservation(pKF, best[q]);
      pKF->AddMapPoint(vpPoints[iMP], best[q]);
    }
    nFused++;
  }
  return nFused;
}

}  // namespace ORB_SLAM3
