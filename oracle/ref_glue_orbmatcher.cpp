// TEST INFRASTRUCTURE -- C entry points over the reference's ORBmatcher (src/ORBmatcher.cc compiled unmodified with
// cvmini/slam_mock_orb.h force-included in place of Frame.h / KeyFrame.h / MapPoint.h; oracle/Makefile.ref).
// Called: SearchByProjection(F, vpMapPoints, th), SearchByProjection(CurrentFrame, LastFrame, th, bMono),
// SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming), SearchForInitialization, SearchByBoW(KF, F),
// SearchByBoW(KF, KF), SearchForTriangulation, Fuse (both overloads), SearchBySim3 -- and through them DescriptorDistance,
// RadiusByViewingCos, ComputeThreeMaxima.  Pose arithmetic runs on the stand-in's small CV_32F algebra with identity
// poses (every product exact; the oracle / CUDA boundary starts at the projected point).  Also the relocalisation
// overload SearchByProjection(F, pKF, sAlreadyFound, th, ORBdist).  Not called: the stereo / two-camera branches.
// Compiled with the same -include so that it sees the same stand-in classes as ORBmatcher.cc.
#include <cstring>
#include <vector>
#include "ORBmatcher.h"   // /root/reference/include (its includes of Frame.h etc. are guarded out by slam_mock_orb.h)

using namespace ORB_SLAM3;

extern "C" void* plvio_grid_create(const void* keys, int n, float minX, float minY, float invW, float invH);
extern "C" void plvio_grid_destroy(void* g);

static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint POD layout");

namespace {
cv::Mat desc_mat(const unsigned char* d, int n) {
  cv::Mat m(n > 0 ? n : 1, 32, CV_8UC1);
  for (int r = 0; r < n; r++) memcpy(m.ptr(r), d + 32 * (size_t)r, 32);
  return m;
}
std::vector<cv::KeyPoint> key_vec(const cv::KeyPoint* k, int n) { return std::vector<cv::KeyPoint>(k, k + n); }
void fill_fv(DBoW2::FeatureVector& fv, const int* nodes, const int* start, const int* feats, int nn) {
  for (int i = 0; i < nn; i++)
    for (int j = start[i]; j < start[i + 1]; j++) fv.addFeature((DBoW2::NodeId)nodes[i], (unsigned int)feats[j]);
}
struct GridOwner {
  void* g;
  float par[4];
  GridOwner(const cv::KeyPoint* k, int n, const float* grid) : g(plvio_grid_create(k, n, grid[0], grid[1], grid[2], grid[3])) {
    memcpy(par, grid, sizeof(par));
  }
  ~GridOwner() { plvio_grid_destroy(g); }
  // the reference's ORBmatcher.cc reaches the grid through GetFeaturesInArea (-> the oracle's grid object); the
  // drop-in build reads the four parameters Frame / KeyFrame hold (mnMinX, mnMinY, mfGridElementWidthInv / HeightInv)
  template <class FrameLike> void attach(FrameLike& f) const {
    f.grid = g;
    f.mnMinX = par[0]; f.mnMinY = par[1];
    f.mfGridElementWidthInv = par[2]; f.mfGridElementHeightInv = par[3];
  }
};
}  // namespace

// ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints=false) (src/ORBmatcher.cc:44-214), monocular frame
// (Nleft == -1, mvuRight < 0, no right-camera projections).  grid = {mnMinX, mnMinY, mfGridElementWidthInv,
// mfGridElementHeightInv}; blocked[i] = F.mvpMapPoints[i] already holds a point with Observations() > 0; per map point:
// proj (x, y), viewCos, predicted level, flags bit0 = !mbTrackInView, bit1 = Observations() == 0, bit2 = isBad().
// match_train[i] = index of the map point assigned to feature i by this call, or -1.
extern "C" int plviref_orb_search_by_projection_mappoints(const cv::KeyPoint* keys, const unsigned char* desc, int n,
                                                           const unsigned char* blocked, const float* grid, const float* scale_factors,
                                                           int nlevels, const float* proj, const float* viewcos, const int* level,
                                                           const int* flags, const unsigned char* qdesc, int nq, float th,
                                                           float nnratio, int* match_train) {
  Frame F;
  F.N = n;
  F.mvKeysUn = F.mvKeys = key_vec(keys, n);
  F.mDescriptors = desc_mat(desc, n);
  F.mvuRight.assign(n, -1.0f);
  F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
  GridOwner go(keys, n, grid);
  go.attach(F);
  MapPoint old;   // stands for every map point the frame held before the call
  old.mObs = 1;
  F.mvpMapPoints.assign(n, nullptr);
  for (int i = 0; i < n; i++) if (blocked && blocked[i]) F.mvpMapPoints[i] = &old;
  std::vector<MapPoint> mps(nq);
  std::vector<MapPoint*> ptrs(nq);
  for (int i = 0; i < nq; i++) {
    MapPoint& m = mps[i];
    m.mnId = i;
    m.mbTrackInView = !(flags[i] & 1);
    m.mObs = (flags[i] & 2) ? 0 : 1;
    m.mBad = (flags[i] & 4) != 0;
    m.mTrackProjX = proj[2 * i];
    m.mTrackProjY = proj[2 * i + 1];
    m.mTrackViewCos = viewcos[i];
    m.mnTrackScaleLevel = level[i];
    m.mDesc = desc_mat(qdesc + 32 * (size_t)i, 1);
    ptrs[i] = &m;
  }
  ORBmatcher matcher(nnratio, true);
  const int k = matcher.SearchByProjection(F, ptrs, th, false, 50.0f);
  for (int i = 0; i < n; i++) match_train[i] = (F.mvpMapPoints[i] && F.mvpMapPoints[i] != &old) ? (int)F.mvpMapPoints[i]->mnId : -1;
  return k;
}

// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:706-821): prev_matched (x, y per F1 key) is updated in place.
extern "C" int plviref_orb_search_for_initialization(const cv::KeyPoint* keys1, const unsigned char* desc1, int n1,
                                                      const cv::KeyPoint* keys2, const unsigned char* desc2, int n2, const float* grid2,
                                                      float* prev_matched, int window, float nnratio, int check_ori, int* matches12) {
  Frame F1, F2;
  F1.N = n1;
  F1.mvKeysUn = F1.mvKeys = key_vec(keys1, n1);
  F1.mDescriptors = desc_mat(desc1, n1);
  F2.N = n2;
  F2.mvKeysUn = F2.mvKeys = key_vec(keys2, n2);
  F2.mDescriptors = desc_mat(desc2, n2);
  GridOwner go(keys2, n2, grid2);
  go.attach(F2);
  std::vector<cv::Point2f> prev(n1);
  for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
  std::vector<int> m12;
  ORBmatcher matcher(nnratio, check_ori != 0);
  const int k = matcher.SearchForInitialization(F1, F2, prev, m12, window);
  for (int i = 0; i < n1; i++) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
  return k;
}

// ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (src/ORBmatcher.cc:269-471), monocular (Nleft == -1, no second
// camera).  mp1[i]: 0 = no map point, 1 = good map point, 2 = bad map point.  Feature vectors as CSR (nodes ascending).
// match_train[i2] = keyframe feature whose map point was assigned to frame feature i2, or -1.
extern "C" int plviref_orb_search_by_bow_kf_f(const cv::KeyPoint* keys1, const unsigned char* desc1, const unsigned char* mp1, int n1,
                                               const int* fv1_nodes, const int* fv1_start, const int* fv1_feats, int nn1,
                                               const cv::KeyPoint* keys2, const unsigned char* desc2, int n2, const int* fv2_nodes,
                                               const int* fv2_start, const int* fv2_feats, int nn2, float nnratio, int check_ori,
                                               int* match_train) {
  KeyFrame KF;
  KF.N = n1;
  KF.mvKeysUn = KF.mvKeys = key_vec(keys1, n1);
  KF.mDescriptors = desc_mat(desc1, n1);
  std::vector<MapPoint> mps(n1);
  KF.mvpMapPoints.assign(n1, nullptr);
  for (int i = 0; i < n1; i++) {
    mps[i].mnId = i;
    mps[i].mBad = mp1[i] == 2;
    if (mp1[i]) KF.mvpMapPoints[i] = &mps[i];
  }
  fill_fv(KF.mFeatVec, fv1_nodes, fv1_start, fv1_feats, nn1);
  Frame F;
  F.N = n2;
  F.mvKeysUn = F.mvKeys = key_vec(keys2, n2);
  F.mDescriptors = desc_mat(desc2, n2);
  fill_fv(F.mFeatVec, fv2_nodes, fv2_start, fv2_feats, nn2);
  std::vector<MapPoint*> out;
  ORBmatcher matcher(nnratio, check_ori != 0);
  const int k = matcher.SearchByBoW(&KF, F, out);
  for (int i = 0; i < n2; i++) match_train[i] = out[i] ? (int)out[i]->mnId : -1;
  return k;
}

// ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (src/ORBmatcher.cc:823-963).  matches12[i1] = KF2 feature whose map
// point was matched to KF1 feature i1, or -1.
extern "C" int plviref_orb_search_by_bow_kf_kf(const cv::KeyPoint* keys1, const unsigned char* desc1, const unsigned char* mp1, int n1,
                                                const int* fv1_nodes, const int* fv1_start, const int* fv1_feats, int nn1,
                                                const cv::KeyPoint* keys2, const unsigned char* desc2, const unsigned char* mp2, int n2,
                                                const int* fv2_nodes, const int* fv2_start, const int* fv2_feats, int nn2,
                                                float nnratio, int check_ori, int* matches12) {
  KeyFrame K1, K2;
  std::vector<MapPoint> m1(n1), m2(n2);
  K1.N = n1;
  K1.mvKeysUn = K1.mvKeys = key_vec(keys1, n1);
  K1.mDescriptors = desc_mat(desc1, n1);
  K1.mvpMapPoints.assign(n1, nullptr);
  for (int i = 0; i < n1; i++) { m1[i].mnId = i; m1[i].mBad = mp1[i] == 2; if (mp1[i]) K1.mvpMapPoints[i] = &m1[i]; }
  fill_fv(K1.mFeatVec, fv1_nodes, fv1_start, fv1_feats, nn1);
  K2.N = n2;
  K2.mvKeysUn = K2.mvKeys = key_vec(keys2, n2);
  K2.mDescriptors = desc_mat(desc2, n2);
  K2.mvpMapPoints.assign(n2, nullptr);
  for (int i = 0; i < n2; i++) { m2[i].mnId = i; m2[i].mBad = mp2[i] == 2; if (mp2[i]) K2.mvpMapPoints[i] = &m2[i]; }
  fill_fv(K2.mFeatVec, fv2_nodes, fv2_start, fv2_feats, nn2);
  std::vector<MapPoint*> out;
  ORBmatcher matcher(nnratio, check_ori != 0);
  const int k = matcher.SearchByBoW(&K1, &K2, out);
  for (int i = 0; i < n1; i++) matches12[i] = out[i] ? (int)out[i]->mnId : -1;
  return k;
}

namespace {
cv::Mat eye_f32(int n) {
  cv::Mat m = cv::Mat::zeros(n, n, CV_32F);
  for (int i = 0; i < n; i++) m.at<float>(i, i) = 1.0f;
  return m;
}
cv::Mat vec3_f32(float x, float y, float z) {
  cv::Mat m(3, 1, CV_32F);
  m.at<float>(0) = x; m.at<float>(1) = y; m.at<float>(2) = z;
  return m;
}
}  // namespace

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono = true) (src/ORBmatcher.cc:1962-2178), monocular
// frames.  The oracle / CUDA boundary starts at the PROJECTED point, so the poses are the identity and map point i sits
// at (uv[i], 1) in front of a unit pinhole camera: Rcw * x + tcw and fx * x / z + cx are then exact and the
// reference's own projection code yields uv[i] bit for bit.  bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}.
// flags bit0: LastFrame holds no map point there (or it is an outlier), bit1: Observations() == 0.
extern "C" int plviref_orb_search_by_projection_frame(const cv::KeyPoint* keys2, const unsigned char* desc2, int n2,
                                                       const unsigned char* blocked, const float* grid, const float* bounds,
                                                       const float* scale_factors, int nlevels, const cv::KeyPoint* keys1, int n1,
                                                       const float* uv, const int* flags, const unsigned char* qdesc, float th,
                                                       int check_ori, int* match_train) {
  GeometricCamera cam;
  Frame C, L;
  C.N = n2;
  C.mvKeysUn = C.mvKeys = key_vec(keys2, n2);
  C.mDescriptors = desc_mat(desc2, n2);
  C.mvuRight.assign(n2, -1.0f);
  C.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
  C.mnMinX = bounds[0]; C.mnMaxX = bounds[1]; C.mnMinY = bounds[2]; C.mnMaxY = bounds[3];
  C.mpCamera = &cam;
  C.mTcw = eye_f32(4);
  GridOwner go(keys2, n2, grid);
  go.attach(C);
  MapPoint old;
  C.mvpMapPoints.assign(n2, nullptr);
  for (int i = 0; i < n2; i++) if (blocked && blocked[i]) C.mvpMapPoints[i] = &old;
  L.N = n1;
  L.mvKeysUn = L.mvKeys = key_vec(keys1, n1);
  L.mTcw = eye_f32(4);
  L.mvbOutlier.assign(n1, false);
  std::vector<MapPoint> mps(n1);
  L.mvpMapPoints.assign(n1, nullptr);
  for (int i = 0; i < n1; i++) {
    MapPoint& m = mps[i];
    m.mnId = i;
    m.mObs = (flags[i] & 2) ? 0 : 1;
    m.mWorldPos = vec3_f32(uv[2 * i], uv[2 * i + 1], 1.0f);
    m.mDesc = desc_mat(qdesc + 32 * (size_t)i, 1);
    if (!(flags[i] & 1)) L.mvpMapPoints[i] = &m;
  }
  ORBmatcher matcher(0.9f, check_ori != 0);
  const int k = matcher.SearchByProjection(C, L, th, true);
  for (int i = 0; i < n2; i++) match_train[i] = (C.mvpMapPoints[i] && C.mvpMapPoints[i] != &old) ? (int)C.mvpMapPoints[i]->mnId : -1;
  return k;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo = false, bCoarse) (src/ORBmatcher.cc:
// 965-1206), monocular pinhole keyframes.  Identity rotations, zero translations and camera centre (ep, 1) make the
// reference's own epipole computation (:972-978) yield ep exactly; F12 is handed to the stand-in camera (see
// slam_mock_orb.h on epipolarConstrain).  mpN[i] != 0: the feature already has a map point.
extern "C" int plviref_orb_search_for_triangulation(const cv::KeyPoint* keys1, const unsigned char* desc1, const unsigned char* mp1, int n1,
                                                     const int* fv1_nodes, const int* fv1_start, const int* fv1_feats, int nn1,
                                                     const cv::KeyPoint* keys2, const unsigned char* desc2, const unsigned char* mp2, int n2,
                                                     const int* fv2_nodes, const int* fv2_start, const int* fv2_feats, int nn2,
                                                     const float* F12, float epx, float epy, const float* sf2, const float* sigma2_1,
                                                     const float* sigma2_2, int nlevels, int coarse, int check_ori, int* matches12) {
  GeometricCamera cam1, cam2;
  memcpy(cam1.mF12, F12, sizeof(cam1.mF12));
  KeyFrame K1, K2;
  MapPoint some;
  KeyFrame* ks[2] = {&K1, &K2};
  const cv::KeyPoint* keys[2] = {keys1, keys2};
  const unsigned char* descs[2] = {desc1, desc2};
  const unsigned char* mps[2] = {mp1, mp2};
  const int ns[2] = {n1, n2};
  for (int s = 0; s < 2; s++) {
    KeyFrame& K = *ks[s];
    K.N = ns[s];
    K.mvKeysUn = K.mvKeys = key_vec(keys[s], ns[s]);
    K.mDescriptors = desc_mat(descs[s], ns[s]);
    K.mvuRight.assign(ns[s], -1.0f);
    K.mvpMapPoints.assign(ns[s], nullptr);
    for (int i = 0; i < ns[s]; i++) if (mps[s][i]) K.mvpMapPoints[i] = &some;
    K.mRcw = eye_f32(3);
    K.mtcw = vec3_f32(0, 0, 0);
    K.mOw = vec3_f32(0, 0, 0);
  }
  K1.mOw = vec3_f32(epx, epy, 1.0f);
  K1.mpCamera = &cam1;
  K2.mpCamera = &cam2;
  K1.mvLevelSigma2.assign(sigma2_1, sigma2_1 + nlevels);
  K2.mvLevelSigma2.assign(sigma2_2, sigma2_2 + nlevels);
  K2.mvScaleFactors.assign(sf2, sf2 + nlevels);
  fill_fv(K1.mFeatVec, fv1_nodes, fv1_start, fv1_feats, nn1);
  fill_fv(K2.mFeatVec, fv2_nodes, fv2_start, fv2_feats, nn2);
  cv::Mat F(3, 3, CV_32F);
  for (int i = 0; i < 9; i++) F.at<float>(i / 3, i % 3) = F12[i];
  std::vector<std::pair<size_t, size_t>> pairs;
  ORBmatcher matcher(0.6f, check_ori != 0);
  const int k = matcher.SearchForTriangulation(&K1, &K2, F, pairs, false, coarse != 0);
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  for (auto& p : pairs) matches12[p.first] = (int)p.second;
  return k;
}

namespace {
// keyframe + map points of the projection searches: identity pose, unit pinhole, map point i at (uv[i], 1) with its
// normal along PO (PO . Pn = |PO|^2 >= 0.5 |PO|), unbounded distance invariance, PredictScale = level[i]: every point
// passes the checks before the search, and the reference's own projection yields uv[i] exactly
void fill_projection_case(KeyFrame& K, GeometricCamera& cam, const cv::KeyPoint* keys, const unsigned char* desc, int n,
                          const float* bounds, const float* scale_factors, const float* inv_sigma2, int nlevels,
                          std::vector<MapPoint>& mps, const float* uv, const int* level, const int* flags, const unsigned char* qdesc) {
  K.N = n;
  K.mvKeysUn = K.mvKeys = key_vec(keys, n);
  K.mDescriptors = desc_mat(desc, n);
  K.mvuRight.assign(n, -1.0f);
  K.mvpMapPoints.assign(n, nullptr);
  K.mnMinX = bounds[0]; K.mnMaxX = bounds[1]; K.mnMinY = bounds[2]; K.mnMaxY = bounds[3];
  K.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
  K.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + nlevels);
  K.mpCamera = &cam;
  K.mRcw = eye_f32(3);
  K.mtcw = vec3_f32(0, 0, 0);
  K.mOw = vec3_f32(0, 0, 0);
  K.fx = K.fy = 1; K.cx = K.cy = 0; K.mbf = 0;
  for (size_t i = 0; i < mps.size(); i++) {
    MapPoint& m = mps[i];
    m.mnId = i;
    m.mBad = (flags[i] & 1) != 0;
    m.mWorldPos = vec3_f32(uv[2 * i], uv[2 * i + 1], 1.0f);
    m.mNormal = m.mWorldPos.clone();   // PO . Pn = |PO|^2 >= 0.5 |PO| (|PO| >= 1)
    m.mnPredLevel = level[i];
    m.mDesc = desc_mat(qdesc + 32 * i, 1);
  }
}
}  // namespace

// ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight = false) (src/ORBmatcher.cc:1399-1610), monocular keyframe.
// flags bit0: isBad().  best_idx[i] = keyframe feature the reference fused map point i with (AddObservation), or -1.
extern "C" int plviref_orb_fuse(const cv::KeyPoint* keys, const unsigned char* desc, int n, const float* grid, const float* bounds,
                                const float* scale_factors, const float* inv_sigma2, int nlevels, const float* uv, const int* level,
                                const int* flags, const unsigned char* qdesc, int nq, float th, int* best_idx) {
  GeometricCamera cam;
  KeyFrame K;
  std::vector<MapPoint> mps(nq);
  fill_projection_case(K, cam, keys, desc, n, bounds, scale_factors, inv_sigma2, nlevels, mps, uv, level, flags, qdesc);
  GridOwner go(keys, n, grid);
  go.attach(K);
  std::vector<MapPoint*> ptrs(nq);
  for (int i = 0; i < nq; i++) ptrs[i] = &mps[i];
  ORBmatcher matcher(0.6f, true);
  const int k = matcher.Fuse(&K, ptrs, th, false);
  for (int i = 0; i < nq; i++) best_idx[i] = mps[i].mFusedIdx;
  return k;
}

// ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (src/ORBmatcher.cc:1612-1734), Scw = identity.
extern "C" int plviref_orb_fuse_sim3(const cv::KeyPoint* keys, const unsigned char* desc, int n, const float* grid, const float* bounds,
                                     const float* scale_factors, const float* inv_sigma2, int nlevels, const float* uv,
                                     const int* level, const int* flags, const unsigned char* qdesc, int nq, float th, int* best_idx) {
  GeometricCamera cam;
  KeyFrame K;
  std::vector<MapPoint> mps(nq);
  fill_projection_case(K, cam, keys, desc, n, bounds, scale_factors, inv_sigma2, nlevels, mps, uv, level, flags, qdesc);
  GridOwner go(keys, n, grid);
  go.attach(K);
  std::vector<MapPoint*> ptrs(nq), repl(nq, nullptr);
  for (int i = 0; i < nq; i++) ptrs[i] = &mps[i];
  ORBmatcher matcher(0.6f, true);
  const int k = matcher.Fuse(&K, eye_f32(4), ptrs, th, repl);
  for (int i = 0; i < nq; i++) best_idx[i] = mps[i].mFusedIdx;
  return k;
}

// ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (src/ORBmatcher.cc:473-586), Scw =
// identity.  matched_in[i] != 0: vpMatched[i] already holds a point on entry (a point not among vpPoints).
// match_train[i] = index of the point of vpPoints assigned to feature i by this call, or -1.
extern "C" int plviref_orb_search_by_projection_kf(const cv::KeyPoint* keys, const unsigned char* desc, int n, const unsigned char* matched_in,
                                                    const float* grid, const float* bounds, const float* scale_factors, int nlevels,
                                                    const float* uv, const int* level, const int* flags, const unsigned char* qdesc,
                                                    int nq, int th, float ratio_hamming, int* match_train) {
  GeometricCamera cam;
  KeyFrame K;
  std::vector<MapPoint> mps(nq);
  std::vector<float> inv(nlevels, 1.0f);
  fill_projection_case(K, cam, keys, desc, n, bounds, scale_factors, inv.data(), nlevels, mps, uv, level, flags, qdesc);
  GridOwner go(keys, n, grid);
  go.attach(K);
  MapPoint old;
  std::vector<MapPoint*> ptrs(nq), matched(n, nullptr);
  for (int i = 0; i < nq; i++) ptrs[i] = &mps[i];
  for (int i = 0; i < n; i++) if (matched_in && matched_in[i]) matched[i] = &old;
  ORBmatcher matcher(0.75f, true);
  const int k = matcher.SearchByProjection(&K, eye_f32(4), ptrs, matched, th, ratio_hamming);
  for (int i = 0; i < n; i++) match_train[i] = (matched[i] && matched[i] != &old) ? (int)matched[i]->mnId : -1;
  return k;
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12 = 1, R12 = I, t12 = 0, th) (src/ORBmatcher.cc:1736-1960): both
// keyframes at the identity pose with a unit pinhole; feature i of keyframe A owns a map point at (uvA[i], 1), which
// the reference's own arithmetic projects to uvA[i] in the other keyframe; levelA[i] = its predicted level there;
// flagsA bit0: no map point, bit1: isBad(), bit2 (keyframe 1 only): vpMatches12[i] already set on entry.
// matches12[i1] = keyframe-2 feature whose map point the call stored in vpMatches12[i1], or -1.
extern "C" int plviref_orb_search_by_sim3(const cv::KeyPoint* keys1, const unsigned char* desc1, int n1, const float* uv1, const int* level1,
                                          const int* flags1, const cv::KeyPoint* keys2, const unsigned char* desc2, int n2,
                                          const float* uv2, const int* level2, const int* flags2, const float* grid, const float* bounds,
                                          const float* scale_factors, int nlevels, float th, int* matches12) {
  GeometricCamera cam;
  KeyFrame K1, K2;
  std::vector<MapPoint> m1(n1), m2(n2);
  std::vector<float> inv(nlevels, 1.0f);
  fill_projection_case(K1, cam, keys1, desc1, n1, bounds, scale_factors, inv.data(), nlevels, m1, uv1, level1, flags1, desc1);
  fill_projection_case(K2, cam, keys2, desc2, n2, bounds, scale_factors, inv.data(), nlevels, m2, uv2, level2, flags2, desc2);
  GridOwner g1(keys1, n1, grid), g2(keys2, n2, grid);
  g1.attach(K1);
  g2.attach(K2);
  MapPoint old;
  std::vector<MapPoint*> out(n1, nullptr);
  for (int i = 0; i < n1; i++) {
    m1[i].mBad = (flags1[i] & 2) != 0;
    if (!(flags1[i] & 1)) K1.mvpMapPoints[i] = &m1[i];
    if (flags1[i] & 4) out[i] = &old;
  }
  for (int i = 0; i < n2; i++) {
    m2[i].mBad = (flags2[i] & 2) != 0;
    if (!(flags2[i] & 1)) K2.mvpMapPoints[i] = &m2[i];
  }
  ORBmatcher matcher(0.75f, true);
  const float s12 = 1.0f;
  const int k = matcher.SearchBySim3(&K1, &K2, out, s12, eye_f32(3), vec3_f32(0, 0, 0), th);
  for (int i = 0; i < n1; i++) matches12[i] = (out[i] && out[i] != &old) ? (int)out[i]->mnId : -1;
  return k;
}

// ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:2180-2302, the
// relocalisation overload): identity frame pose, unit pinhole, the map point of keyframe feature i at (uv[i], 1) with
// predicted level[i].  flags bit0: no map point, bit1: isBad(), bit2: in sAlreadyFound.  blocked[i2] != 0: the
// frame feature already holds a map point.  match_train[i2] = keyframe feature assigned by this call, or -1.
extern "C" int plviref_orb_search_by_projection_reloc(const cv::KeyPoint* keys2, const unsigned char* desc2, int n2,
                                                       const unsigned char* blocked, const float* grid, const float* bounds,
                                                       const float* scale_factors, int nlevels, const cv::KeyPoint* keys1, int n1,
                                                       const float* uv, const int* level, const int* flags, const unsigned char* qdesc,
                                                       float th, int orb_dist, int check_ori, int* match_train) {
  GeometricCamera cam;
  Frame C;
  C.N = n2;
  C.mvKeysUn = C.mvKeys = key_vec(keys2, n2);
  C.mDescriptors = desc_mat(desc2, n2);
  C.mvuRight.assign(n2, -1.0f);
  C.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
  C.mnMinX = bounds[0]; C.mnMaxX = bounds[1]; C.mnMinY = bounds[2]; C.mnMaxY = bounds[3];
  C.mpCamera = &cam;
  C.mTcw = eye_f32(4);
  GridOwner go(keys2, n2, grid);
  go.attach(C);
  MapPoint old;
  C.mvpMapPoints.assign(n2, nullptr);
  for (int i = 0; i < n2; i++) if (blocked && blocked[i]) C.mvpMapPoints[i] = &old;
  KeyFrame K;
  K.N = n1;
  K.mvKeysUn = K.mvKeys = key_vec(keys1, n1);
  K.mvpMapPoints.assign(n1, nullptr);
  std::vector<MapPoint> mps(n1);
  std::set<MapPoint*> found;
  for (int i = 0; i < n1; i++) {
    MapPoint& m = mps[i];
    m.mnId = i;
    m.mBad = (flags[i] & 2) != 0;
    m.mWorldPos = vec3_f32(uv[2 * i], uv[2 * i + 1], 1.0f);
    m.mnPredLevel = level[i];
    m.mDesc = desc_mat(qdesc + 32 * (size_t)i, 1);
    if (!(flags[i] & 1)) K.mvpMapPoints[i] = &m;
    if (flags[i] & 4) found.insert(&m);
  }
  ORBmatcher matcher(0.9f, check_ori != 0);
  const int k = matcher.SearchByProjection(C, &K, found, th, orb_dist);
  for (int i = 0; i < n2; i++) match_train[i] = (C.mvpMapPoints[i] && C.mvpMapPoints[i] != &old) ? (int)C.mvpMapPoints[i]->mnId : -1;
  return k;
}

// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2350-2366)
extern "C" int plviref_orb_descriptor_distance(const unsigned char* a, const unsigned char* b) {
  return ORBmatcher::DescriptorDistance(desc_mat(a, 1), desc_mat(b, 1));
}
