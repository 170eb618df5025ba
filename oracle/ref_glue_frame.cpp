// TEST INFRASTRUCTURE -- C entry points over the reference's Frame (src/Frame.cc + include/Frame.h compiled unmodified;
// MapPoint / KeyFrame / MapLine / cameras / IMU types / vocabulary are the stand-ins of cvmini/slam_mock_frame.h;
// ORBextractor.cc, ORBmatcher.cc, gridStructure.cpp, LineIterator.cpp are the reference's own).  oracle/Makefile.ref
// builds it into a library of its own (oracle/_ref/libplvi_ref_frame.so).  Frames are default-constructed and their
// public data members filled in; private member functions are reached through explicit template instantiation (the
// standard's access exemption), so that Frame.cc and Frame.h stay untouched.
#include <cstring>
#include <vector>
#include "Frame.h"          // /root/reference/include
#include "ORBextractor.h"   // /root/reference/include

using namespace ORB_SLAM3;

namespace {
template <typename Tag, typename Tag::type M> struct Rob { friend typename Tag::type get(Tag) { return M; } };
struct TagUndistortKP { typedef void (Frame::*type)(); friend type get(TagUndistortKP); };
struct TagUndistortKL { typedef void (Frame::*type)(); friend type get(TagUndistortKL); };
struct TagAssignGrid { typedef void (Frame::*type)(); friend type get(TagAssignGrid); };
}  // namespace
template struct Rob<TagUndistortKP, &Frame::UndistortKeyPoints>;
template struct Rob<TagUndistortKL, &Frame::UndistortKeyLines>;
template struct Rob<TagAssignGrid, &Frame::AssignFeaturesToGrid>;

namespace {
cv::Mat desc_mat(const unsigned char* d, int n) {
  cv::Mat m(n > 0 ? n : 1, 32, CV_8UC1);
  for (int r = 0; r < n; r++) memcpy(m.ptr(r), d + 32 * (size_t)r, 32);
  return m;
}
void set_bounds_and_grid(const float* bounds) {   // what Frame's constructors compute once (src/Frame.cc:150-170)
  Frame::mnMinX = bounds[0]; Frame::mnMaxX = bounds[1]; Frame::mnMinY = bounds[2]; Frame::mnMaxY = bounds[3];
  Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
  Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
}
}  // namespace

// Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:644-675, 1077-1087): mGrid as a CSR (cell = i * 48 + j).
extern "C" void plviref_frame_assign_grid(const cv::KeyPoint* keys, int n, const float* bounds, int* cell_start, int* items) {
  set_bounds_and_grid(bounds);
  Frame F;
  F.N = n;
  F.Nleft = -1;
  F.mvKeysUn.assign(keys, keys + n);
  (F.*get(TagAssignGrid()))();
  int pos = 0;
  for (int i = 0; i < FRAME_GRID_COLS; i++)
    for (int j = 0; j < FRAME_GRID_ROWS; j++) {
      cell_start[i * FRAME_GRID_ROWS + j] = pos;
      for (size_t v : F.mGrid[i][j]) items[pos++] = (int)v;
    }
  cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = pos;
}

// Frame::GetFeaturesInArea (src/Frame.cc:1006-1075) for nq queries (x, y, r, minLevel, maxLevel): CSR of index lists.
extern "C" void plviref_frame_features_in_area(const cv::KeyPoint* keys, int n, const float* bounds, const float* xyr, const int* levels,
                                               int nq, int* start, int* out, int cap) {
  set_bounds_and_grid(bounds);
  Frame F;
  F.N = n;
  F.Nleft = -1;
  F.mvKeysUn.assign(keys, keys + n);
  (F.*get(TagAssignGrid()))();
  int pos = 0;
  for (int q = 0; q < nq; q++) {
    start[q] = pos;
    const std::vector<size_t> v = F.GetFeaturesInArea(xyr[3 * q], xyr[3 * q + 1], xyr[3 * q + 2], levels[2 * q], levels[2 * q + 1]);
    for (size_t i : v) if (pos < cap) out[pos++] = (int)i;
  }
  start[nq] = pos;
}

// Frame::lineDescriptorMAD (src/Frame.cc:1089-1113) on the kNN-2 distances.
extern "C" void plviref_frame_line_descriptor_mad(const int* d0, const int* d1, int n, double* nn_mad, double* nn12_mad) {
  std::vector<std::vector<cv::DMatch>> m(n, std::vector<cv::DMatch>(2));
  for (int i = 0; i < n; i++) { m[i][0] = cv::DMatch(i, 0, (float)d0[i]); m[i][1] = cv::DMatch(i, 1, (float)d1[i]); }
  Frame F;
  F.lineDescriptorMAD(m, *nn_mad, *nn12_mad);
}

// Frame::UndistortKeyPoints (src/Frame.cc:1124-1157): K = (fx, fy, cx, cy); dist = nd CV_32F coefficients.
extern "C" void plviref_frame_undistort_keypoints(const cv::KeyPoint* keys, int n, const float* K, const float* dist, int nd,
                                                  cv::KeyPoint* out) {
  GeometricCamera cam;
  cam.fx = K[0]; cam.fy = K[1]; cam.cx = K[2]; cam.cy = K[3];
  Frame F;
  F.N = n;
  F.mvKeys.assign(keys, keys + n);
  F.mpCamera = &cam;
  F.mK = cam.toK();
  F.mDistCoef = cv::Mat(nd, 1, CV_32F);
  for (int i = 0; i < nd; i++) F.mDistCoef.at<float>(i) = dist[i];
  (F.*get(TagUndistortKP()))();
  for (int i = 0; i < n; i++) out[i] = F.mvKeysUn[i];
}

// Frame::UndistortKeyLines (src/Frame.cc:1159-1197): the four endpoint coordinates of every KeyLine.
extern "C" void plviref_frame_undistort_keylines(const cv::line_descriptor::KeyLine* kl, int n, const float* K, const float* dist, int nd,
                                                 cv::line_descriptor::KeyLine* out) {
  GeometricCamera cam;
  cam.fx = K[0]; cam.fy = K[1]; cam.cx = K[2]; cam.cy = K[3];
  Frame F;
  F.mvKeys_Line.assign(kl, kl + n);
  F.N_l = n;
  F.mK = cam.toK();
  F.mDistCoef = cv::Mat(nd, 1, CV_32F);
  for (int i = 0; i < nd; i++) F.mDistCoef.at<float>(i) = dist[i];
  (F.*get(TagUndistortKL()))();
  for (int i = 0; i < n; i++) out[i] = F.mvKeysUn_Line[i];
}

// Frame::ComputeStereoMatches (src/Frame.cc:1228-1406) on given left / right keypoints, descriptors and the two image
// pyramids (level images WITHOUT the extractor's border; the function only reads inside them).
extern "C" int plviref_frame_compute_stereo_matches(const cv::KeyPoint* kl, const unsigned char* dl, int nl, const cv::KeyPoint* kr,
                                                    const unsigned char* dr, int nr, const unsigned char* const* pyr_l,
                                                    const unsigned char* const* pyr_r, const int* widths, const int* heights,
                                                    int nlevels, const float* scale_factors, float mb, float mbf, float* u_right,
                                                    float* depth) {
  ORBextractor L(1000, 1.2f, nlevels, 20, 7), R(1000, 1.2f, nlevels, 20, 7);
  for (int l = 0; l < nlevels; l++) {
    cv::Mat a(heights[l], widths[l], CV_8UC1), b(heights[l], widths[l], CV_8UC1);
    for (int y = 0; y < heights[l]; y++) {
      memcpy(a.ptr(y), pyr_l[l] + (size_t)y * widths[l], widths[l]);
      memcpy(b.ptr(y), pyr_r[l] + (size_t)y * widths[l], widths[l]);
    }
    L.mvImagePyramid[l] = a;
    R.mvImagePyramid[l] = b;
  }
  Frame F;
  F.N = nl;
  F.mvKeys.assign(kl, kl + nl);
  F.mvKeysRight.assign(kr, kr + nr);
  F.mDescriptors = desc_mat(dl, nl);
  F.mDescriptorsRight = desc_mat(dr, nr);
  F.mpORBextractorLeft = &L;
  F.mpORBextractorRight = &R;
  F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
  F.mvInvScaleFactors.resize(nlevels);
  for (int l = 0; l < nlevels; l++) F.mvInvScaleFactors[l] = 1.0f / scale_factors[l];
  F.mb = mb;
  F.mbf = mbf;
  F.ComputeStereoMatches();
  int k = 0;
  for (int i = 0; i < nl; i++) { u_right[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; k += F.mvuRight[i] >= 0; }
  return k;
}

#include "LineMatcher.h"   // src/LineMatcher.cpp is part of this library (compiled in this class set, Makefile.ref)

// ---- KeyFrame (src/KeyFrame.cc + include/KeyFrame.h compiled unmodified in the same class set) ----------------------
#include "KeyFrame.h"   // /root/reference/include
#include "Map.h"        // guarded out: the stand-in Map / KeyFrameDatabase of slam_mock_keyframe.h

namespace {
// A keyframe is built the way the reference builds it: KeyFrame(Frame&, Map*, KeyFrameDatabase*) copies the frame's
// keys, grid and bounds (src/KeyFrame.cc:51-106).
struct KeyFrameCase {
  Map map;
  KeyFrameDatabase db;
  Frame F;
  KeyFrame* K = nullptr;
  KeyFrameCase(const cv::KeyPoint* keys, int n, const cv::line_descriptor::KeyLine* kl, int nl, const float* bounds) {
    set_bounds_and_grid(bounds);
    F.N = n;
    F.Nleft = -1;
    F.Nright = -1;
    if (n) { F.mvKeysUn.assign(keys, keys + n); F.mvKeys = F.mvKeysUn; }
    F.mvuRight.assign(n, -1.0f);
    F.mvDepth.assign(n, -1.0f);
    F.mvpMapPoints.assign(n, nullptr);
    if (nl) { F.mvKeys_Line.assign(kl, kl + nl); F.mvKeysUn_Line = F.mvKeys_Line; }
    F.N_l = nl;
    F.mvpMapLines.assign(nl, nullptr);
    F.mTcw = cv::Mat::eye(4, 4, CV_32F);
    (F.*get(TagAssignGrid()))();
    K = new KeyFrame(F, &map, &db);
  }
  ~KeyFrameCase() { delete K; }
};
}  // namespace

// KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:1200-1244) for nq queries (x, y, r): CSR of index lists.
extern "C" void plviref_keyframe_features_in_area(const cv::KeyPoint* keys, int n, const float* bounds, const float* xyr, int nq,
                                                  int* start, int* out, int cap) {
  KeyFrameCase c(keys, n, nullptr, 0, bounds);
  int pos = 0;
  for (int q = 0; q < nq; q++) {
    start[q] = pos;
    const std::vector<size_t> v = c.K->GetFeaturesInArea(xyr[3 * q], xyr[3 * q + 1], xyr[3 * q + 2]);
    for (size_t i : v) if (pos < cap) out[pos++] = (int)i;
  }
  start[nq] = pos;
}

// KeyFrame::GetLinesInArea(x1, y1, x2, y2, r) (src/KeyFrame.cc:1170-1198) for nq queries of 5 floats: CSR of index lists.
extern "C" void plviref_keyframe_lines_in_area(const cv::line_descriptor::KeyLine* kl, int nl, const float* bounds, const float* q5,
                                               int nq, int* start, int* out, int cap) {
  KeyFrameCase c(nullptr, 0, kl, nl, bounds);
  int pos = 0;
  for (int q = 0; q < nq; q++) {
    start[q] = pos;
    const float* p = q5 + 5 * (size_t)q;
    const std::vector<size_t> v = c.K->GetLinesInArea(p[0], p[1], p[2], p[3], p[4]);
    for (size_t i : v) if (pos < cap) out[pos++] = (int)i;
  }
  start[nq] = pos;
}

// KeyFrame::lineDescriptorMAD (src/KeyFrame.cc:411-435) and KeyFrame::IsInImage (:1246-1249).
extern "C" void plviref_keyframe_line_descriptor_mad(const int* d0, const int* d1, int n, double* nn_mad, double* nn12_mad) {
  const float bounds[4] = {0, 752, 0, 480};
  KeyFrameCase c(nullptr, 0, nullptr, 0, bounds);
  std::vector<std::vector<cv::DMatch>> m(n, std::vector<cv::DMatch>(2));
  for (int i = 0; i < n; i++) { m[i][0] = cv::DMatch(i, 0, (float)d0[i]); m[i][1] = cv::DMatch(i, 1, (float)d1[i]); }
  c.K->lineDescriptorMAD(m, *nn_mad, *nn12_mad);
}

// ---- ORBmatcher.cc compiled against THIS class set (the reference's own Frame / KeyFrame, stand-in MapPoint): the three
// tracking searches end to end, i.e. with the reference's own AssignFeaturesToGrid + GetFeaturesInArea underneath
// (in libplvi_ref_orbmatcher.so the stand-in Frame borrows those two from the oracle).  Same arguments as the entry
// points of the same name without "_realframe" in ref_glue_orbmatcher.cpp, bounds instead of the grid record.
#include "ORBmatcher.h"

namespace {
cv::Mat vec3(float x, float y, float z) {
  cv::Mat m(3, 1, CV_32F);
  m.at<float>(0) = x; m.at<float>(1) = y; m.at<float>(2) = z;
  return m;
}
void fill_frame(Frame& F, const cv::KeyPoint* keys, const unsigned char* desc, int n, const float* scale_factors, int nlevels) {
  F.N = n;
  F.Nleft = -1;
  F.Nright = -1;
  F.mvKeysUn.assign(keys, keys + n);
  F.mvKeys = F.mvKeysUn;
  F.mDescriptors = desc_mat(desc, n);
  F.mvuRight.assign(n, -1.0f);
  F.mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
  F.mvpMapPoints.assign(n, nullptr);
  F.mvbOutlier.assign(n, false);
  F.mTcw = cv::Mat::eye(4, 4, CV_32F);
  (F.*get(TagAssignGrid()))();
}
}  // namespace

extern "C" int plviref_orb_search_by_projection_mappoints_realframe(const cv::KeyPoint* keys, const unsigned char* desc, int n,
                                                                     const unsigned char* blocked, const float* bounds,
                                                                     const float* scale_factors, int nlevels, const float* proj,
                                                                     const float* viewcos, const int* level, const int* flags,
                                                                     const unsigned char* qdesc, int nq, float th, float nnratio,
                                                                     int* match_train) {
  set_bounds_and_grid(bounds);
  Frame F;
  fill_frame(F, keys, desc, n, scale_factors, nlevels);
  MapPoint old;
  old.mObs = 1;
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

extern "C" int plviref_orb_search_by_projection_frame_realframe(const cv::KeyPoint* keys2, const unsigned char* desc2, int n2,
                                                                 const unsigned char* blocked, const float* bounds,
                                                                 const float* scale_factors, int nlevels, const cv::KeyPoint* keys1,
                                                                 int n1, const float* uv, const int* flags, const unsigned char* qdesc,
                                                                 float th, int check_ori, int* match_train) {
  set_bounds_and_grid(bounds);
  GeometricCamera cam;
  Frame C, L;
  fill_frame(C, keys2, desc2, n2, scale_factors, nlevels);
  C.mpCamera = &cam;
  MapPoint old;
  for (int i = 0; i < n2; i++) if (blocked && blocked[i]) C.mvpMapPoints[i] = &old;
  L.N = n1;
  L.Nleft = -1;
  L.mvKeysUn.assign(keys1, keys1 + n1);
  L.mvKeys = L.mvKeysUn;
  L.mTcw = cv::Mat::eye(4, 4, CV_32F);
  L.mvbOutlier.assign(n1, false);
  std::vector<MapPoint> mps(n1);
  L.mvpMapPoints.assign(n1, nullptr);
  for (int i = 0; i < n1; i++) {
    MapPoint& m = mps[i];
    m.mnId = i;
    m.mObs = (flags[i] & 2) ? 0 : 1;
    m.mWorldPos = vec3(uv[2 * i], uv[2 * i + 1], 1.0f);
    m.mDesc = desc_mat(qdesc + 32 * (size_t)i, 1);
    if (!(flags[i] & 1)) L.mvpMapPoints[i] = &m;
  }
  ORBmatcher matcher(0.9f, check_ori != 0);
  const int k = matcher.SearchByProjection(C, L, th, true);
  for (int i = 0; i < n2; i++) match_train[i] = (C.mvpMapPoints[i] && C.mvpMapPoints[i] != &old) ? (int)C.mvpMapPoints[i]->mnId : -1;
  return k;
}

extern "C" int plviref_orb_search_for_initialization_realframe(const cv::KeyPoint* keys1, const unsigned char* desc1, int n1,
                                                                const cv::KeyPoint* keys2, const unsigned char* desc2, int n2,
                                                                const float* bounds, float* prev_matched, int window, float nnratio,
                                                                int check_ori, int* matches12) {
  set_bounds_and_grid(bounds);
  const float sf1[1] = {1.0f};
  Frame F1, F2;
  fill_frame(F1, keys1, desc1, n1, sf1, 1);
  fill_frame(F2, keys2, desc2, n2, sf1, 1);
  std::vector<cv::Point2f> prev(n1);
  for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
  std::vector<int> m12;
  ORBmatcher matcher(nnratio, check_ori != 0);
  const int k = matcher.SearchForInitialization(F1, F2, prev, m12, window);
  for (int i = 0; i < n1; i++) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
  return k;
}

// ---- keyframe searches of ORBmatcher.cc on the reference's own KeyFrame class (built by its own constructor from a
// frame; identity pose, unit pinhole: see ref_glue_orbmatcher.cpp for the conventions)
namespace {
struct KeyFrameSearchCase {
  GeometricCamera cam;
  Map map;
  KeyFrameDatabase db;
  Frame F;
  KeyFrame* K = nullptr;
  KeyFrameSearchCase(const cv::KeyPoint* keys, const unsigned char* desc, int n, const float* bounds, const float* scale_factors,
                     const float* inv_sigma2, int nlevels, const std::vector<MapPoint*>* mps = nullptr,
                     const DBoW2::FeatureVector* fv = nullptr) {
    set_bounds_and_grid(bounds);
    Frame::fx = Frame::fy = 1.0f; Frame::cx = Frame::cy = 0.0f; Frame::invfx = Frame::invfy = 1.0f;
    fill_frame(F, keys, desc, n, scale_factors, nlevels);
    F.mvDepth.assign(n, -1.0f);
    F.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + nlevels);
    F.mvLevelSigma2.resize(nlevels);
    for (int l = 0; l < nlevels; l++) F.mvLevelSigma2[l] = 1.0f / inv_sigma2[l];
    F.mnScaleLevels = nlevels;
    F.mpCamera = &cam;
    F.mb = 0; F.mbf = 0;
    if (mps) F.mvpMapPoints = *mps;
    if (fv) F.mFeatVec = *fv;
    K = new KeyFrame(F, &map, &db);
  }
  ~KeyFrameSearchCase() { delete K; }
};
void fill_points(std::vector<MapPoint>& mps, const float* uv, const int* level, const int* flags, const unsigned char* qdesc) {
  for (size_t i = 0; i < mps.size(); i++) {
    MapPoint& m = mps[i];
    m.mnId = i;
    m.mBad = (flags[i] & 1) != 0;
    m.mWorldPos = vec3(uv[2 * i], uv[2 * i + 1], 1.0f);
    m.mNormal = m.mWorldPos.clone();
    m.mnPredLevel = level[i];
    m.mDesc = desc_mat(qdesc + 32 * i, 1);
  }
}
void fill_fv2(DBoW2::FeatureVector& fv, const int* nodes, const int* start, const int* feats, int nn) {
  for (int i = 0; i < nn; i++)
    for (int j = start[i]; j < start[i + 1]; j++) fv.addFeature((DBoW2::NodeId)nodes[i], (unsigned int)feats[j]);
}
}  // namespace

// ORBmatcher::Fuse(pKF, vpMapPoints, th) / Fuse(pKF, Scw = I, vpPoints, th, vpReplacePoint) on the reference's KeyFrame.
extern "C" int plviref_orb_fuse_realkeyframe(const cv::KeyPoint* keys, const unsigned char* desc, int n, const float* bounds,
                                             const float* scale_factors, const float* inv_sigma2, int nlevels, const float* uv,
                                             const int* level, const int* flags, const unsigned char* qdesc, int nq, float th, int sim3,
                                             int* best_idx) {
  KeyFrameSearchCase c(keys, desc, n, bounds, scale_factors, inv_sigma2, nlevels);
  std::vector<MapPoint> mps(nq);
  fill_points(mps, uv, level, flags, qdesc);
  std::vector<MapPoint*> ptrs(nq), repl(nq, nullptr);
  for (int i = 0; i < nq; i++) ptrs[i] = &mps[i];
  ORBmatcher matcher(0.6f, true);
  const int k = sim3 ? matcher.Fuse(c.K, cv::Mat::eye(4, 4, CV_32F), ptrs, th, repl) : matcher.Fuse(c.K, ptrs, th, false);
  for (int i = 0; i < nq; i++) {
    best_idx[i] = mps[i].mFusedIdx;
    if (sim3 && repl[i]) best_idx[i] = repl[i]->mFusedIdx;   // vpReplacePoint[iMP] = the point already sitting on that feature
  }
  return k;
}

// ORBmatcher::SearchByProjection(pKF, Scw = I, vpPoints, vpMatched, th, ratioHamming) on the reference's KeyFrame.
extern "C" int plviref_orb_search_by_projection_kf_realkeyframe(const cv::KeyPoint* keys, const unsigned char* desc, int n,
                                                                 const unsigned char* matched_in, const float* bounds,
                                                                 const float* scale_factors, int nlevels, const float* uv,
                                                                 const int* level, const int* flags, const unsigned char* qdesc, int nq,
                                                                 int th, float ratio_hamming, int* match_train) {
  std::vector<float> inv(nlevels, 1.0f);
  KeyFrameSearchCase c(keys, desc, n, bounds, scale_factors, inv.data(), nlevels);
  std::vector<MapPoint> mps(nq);
  fill_points(mps, uv, level, flags, qdesc);
  MapPoint old;
  std::vector<MapPoint*> ptrs(nq), matched(n, nullptr);
  for (int i = 0; i < nq; i++) ptrs[i] = &mps[i];
  for (int i = 0; i < n; i++) if (matched_in && matched_in[i]) matched[i] = &old;
  ORBmatcher matcher(0.75f, true);
  const int k = matcher.SearchByProjection(c.K, cv::Mat::eye(4, 4, CV_32F), ptrs, matched, th, ratio_hamming);
  for (int i = 0; i < n; i++) match_train[i] = (matched[i] && matched[i] != &old) ? (int)matched[i]->mnId : -1;
  return k;
}

// ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) on the reference's KeyFrame and Frame.
extern "C" int plviref_orb_search_by_bow_kf_f_real(const cv::KeyPoint* keys1, const unsigned char* desc1, const unsigned char* mp1, int n1,
                                                    const int* fv1_nodes, const int* fv1_start, const int* fv1_feats, int nn1,
                                                    const cv::KeyPoint* keys2, const unsigned char* desc2, int n2, const int* fv2_nodes,
                                                    const int* fv2_start, const int* fv2_feats, int nn2, const float* bounds,
                                                    float nnratio, int check_ori, int* match_train) {
  const float one[1] = {1.0f};
  std::vector<MapPoint> mps(n1);
  std::vector<MapPoint*> ptrs(n1, nullptr);
  for (int i = 0; i < n1; i++) {
    mps[i].mnId = i;
    mps[i].mBad = mp1[i] == 2;
    if (mp1[i]) ptrs[i] = &mps[i];
  }
  DBoW2::FeatureVector fv1;
  fill_fv2(fv1, fv1_nodes, fv1_start, fv1_feats, nn1);
  KeyFrameSearchCase c(keys1, desc1, n1, bounds, one, one, 1, &ptrs, &fv1);
  Frame F;
  fill_frame(F, keys2, desc2, n2, one, 1);
  fill_fv2(F.mFeatVec, fv2_nodes, fv2_start, fv2_feats, nn2);
  std::vector<MapPoint*> out;
  ORBmatcher matcher(nnratio, check_ori != 0);
  const int k = matcher.SearchByBoW(c.K, F, out);
  for (int i = 0; i < n2; i++) match_train[i] = out[i] ? (int)out[i]->mnId : -1;
  return k;
}

// =====================================================================================================================
// The consumer's call pattern, end to end on the reference's own Frame class (drop-in proof, tests/test_dropin_gpu.py):
//   Frame::Frame(imGray, timeStamp, extractor, Lineextractor, voc, voc_l, pCamera, distCoef, bf, thDepth, pPrevF, ImuCalib)
//     (src/Frame.cc:537-642: the monocular point-line constructor of Tracking::GrabImageMonocular, src/Tracking.cc:1467):
//     two threads run ORBextractor::operator() and Lineextractor::operator() (:558-561), then UndistortKeyPoints,
//     UndistortKeyLines, ComputeImageBounds, AssignFeaturesToGrid;
//   ORBmatcher::SearchByProjection(mCurrentFrame, mLastFrame, th, bMono) + LineMatcher::match(mLastFrame.mDescriptors_Line,
//     mCurrentFrame.mDescriptors_Line, 0.9, matches_12)   (Tracking::TrackWithMotionModelWithLines, src/Tracking.cc:3957,3990).
// In libplvi_ref_frame.so the extractors and matchers are the reference's own sources; in libplvi_dropin_frame.so they
// are the product's drop-in headers + libplvi_cuda.so.  Frame.cc is the same unmodified file in both.
// =====================================================================================================================
#include "LineExtractor.h"
#include "gridStructure.h"
#include <memory>

#ifndef PLVI_DROPIN
// Monotone allocator (see ref_glue.cpp): DistributeOctTree orders nodes of equal size by HEAP ADDRESS
// (src/ORBextractor.cc:682).  While a tracker frame is being built, operator new inside this library (linked with
// -Bsymbolic-functions) hands out strictly increasing addresses from one lazily committed range, which makes the
// reference's own code realise "node created earlier first" -- the definition the oracle and the CUDA path implement.
// The Frame outlives the call, so the range lives until the next plviref_track_create (one tracker at a time); the two
// extractor threads share an atomic bump pointer (addresses still increase in time order).
#include <sys/mman.h>
#include <atomic>
#include <cstdlib>
#include <new>
namespace {
const size_t kTrackArena = (size_t)64 << 30;
char* g_tbase = nullptr;
std::atomic<char*> g_tcur{nullptr};
std::atomic<bool> g_ton{false};
void track_arena_reset() {
  if (!g_tbase) {
    void* m = mmap(nullptr, kTrackArena, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) { fprintf(stderr, "libplvi_ref_frame: cannot reserve the arena\n"); abort(); }
    g_tbase = (char*)m;
  } else {
    madvise(g_tbase, (size_t)(g_tcur.load() - g_tbase), MADV_DONTNEED);
  }
  g_tcur = g_tbase;
}
inline bool track_arena_owns(void* p) { return g_tbase && (char*)p >= g_tbase && (char*)p < g_tbase + kTrackArena; }
struct TrackArenaScope {
  TrackArenaScope() { g_ton = true; }
  ~TrackArenaScope() { g_ton = false; }
};
}  // namespace
void* operator new(size_t n) {
  if (g_ton.load(std::memory_order_relaxed)) {
    n = (n + 15) & ~(size_t)15;
    char* p = g_tcur.fetch_add((ptrdiff_t)n);
    if (p + n > g_tbase + kTrackArena) { fprintf(stderr, "libplvi_ref_frame: arena exhausted\n"); abort(); }
    return p;
  }
  void* p = malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { if (p && !track_arena_owns(p)) free(p); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }
#else
namespace { inline void track_arena_reset() {} struct TrackArenaScope {}; }
#endif

namespace {
struct TrackCase {
  std::unique_ptr<ORBextractor> orb;
  std::unique_ptr<Lineextractor> line;
  Pinhole cam;
  IMU::Calib calib;
  std::unique_ptr<Frame> cur, last;
  std::vector<MapPoint> points;   // one per keypoint of `last`
};
}  // namespace

extern "C" void* plviref_track_create(int nfeatures, float sf, int nlevels, int ini_th, int min_th, int lsd_nfeatures, int lsd_refine,
                                      float lsd_scale, int line_levels, float line_scale) {
  track_arena_reset();
  TrackCase* t = new TrackCase();
  t->orb.reset(new ORBextractor(nfeatures, sf, nlevels, ini_th, min_th));
  t->line.reset(new Lineextractor(lsd_nfeatures, lsd_refine, lsd_scale, line_levels, line_scale, 0));
  Frame::mbInitialComputations = true;
  return t;
}
extern "C" void plviref_track_destroy(void* h) { delete (TrackCase*)h; }

// One frame through the reference's own constructor.  K = (fx, fy, cx, cy); dist = nd CV_32F coefficients.  Returns N.
extern "C" int plviref_track_frame(void* h, const unsigned char* img, int w, int hh, int stride, const float* K, const float* dist, int nd) {
  TrackCase* t = (TrackCase*)h;
  t->cam.fx = K[0]; t->cam.fy = K[1]; t->cam.cx = K[2]; t->cam.cy = K[3];
  cv::Mat image(hh, w, CV_8UC1, (void*)img, (size_t)stride);
  cv::Mat d(nd, 1, CV_32F);
  for (int i = 0; i < nd; i++) d.at<float>(i) = dist[i];
  t->last = std::move(t->cur);
  TrackArenaScope scope;
  t->cur.reset(new Frame(image, 0.0, t->orb.get(), t->line.get(), (ORBVocabulary*)nullptr, (LineVocabulary*)nullptr, &t->cam, d, 40.0f, 35.0f,
                         (Frame*)nullptr, t->calib));
  return t->cur->N;
}

// Members of the current (which = 0) / last (1) frame.  field: 0 mvKeys, 1 mvKeysUn (28 B each), 2 mDescriptors (32 B rows),
// 3 mvKeys_Line, 4 mvKeysUn_Line (68 B each), 5 mDescriptors_Line, 6 mvKeyLineFunctions (3 doubles each), 7 mGrid as CSR
// ints (64 * 48 + 1 starts, then the items), 8 {mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, HeightInv} floats,
// 9 mvScaleFactors ++ mvInvScaleFactors ++ mvLevelSigma2 ++ mvInvLevelSigma2, 10 the same four vectors of the line
// extractor (first mnScaleLevels_l entries each), 11 {N, N_l, mnScaleLevels, mnScaleLevels_l, monoLeft} ints.
// Returns the number of records (bytes for 7..11).
extern "C" int plviref_track_get(void* h, int which, int field, void* out, int cap_bytes) {
  TrackCase* t = (TrackCase*)h;
  Frame* F = which ? t->last.get() : t->cur.get();
  if (!F) return -1;
  std::vector<unsigned char> buf;
  int count = 0;
  auto put = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; buf.insert(buf.end(), b, b + n); };
  switch (field) {
    case 0: count = (int)F->mvKeys.size(); if (count) put(F->mvKeys.data(), sizeof(cv::KeyPoint) * count); break;
    case 1: count = (int)F->mvKeysUn.size(); if (count) put(F->mvKeysUn.data(), sizeof(cv::KeyPoint) * count); break;
    case 2: count = F->mDescriptors.rows; for (int i = 0; i < count; i++) put(F->mDescriptors.ptr(i), 32); break;
    case 3: count = (int)F->mvKeys_Line.size(); if (count) put(F->mvKeys_Line.data(), sizeof(KeyLine) * count); break;
    case 4: count = (int)F->mvKeysUn_Line.size(); if (count) put(F->mvKeysUn_Line.data(), sizeof(KeyLine) * count); break;
    case 5: count = F->mvKeys_Line.empty() ? 0 : F->mDescriptors_Line.rows; for (int i = 0; i < count; i++) put(F->mDescriptors_Line.ptr(i), 32); break;
    case 6: count = (int)F->mvKeyLineFunctions.size(); for (int i = 0; i < count; i++) { double v[3] = {F->mvKeyLineFunctions[i](0), F->mvKeyLineFunctions[i](1), F->mvKeyLineFunctions[i](2)}; put(v, sizeof(v)); } break;
    case 7: {
      std::vector<int> start, items;
      for (int i = 0; i < FRAME_GRID_COLS; i++)
        for (int j = 0; j < FRAME_GRID_ROWS; j++) { start.push_back((int)items.size()); for (size_t v : F->mGrid[i][j]) items.push_back((int)v); }
      start.push_back((int)items.size());
      put(start.data(), start.size() * sizeof(int));
      if (!items.empty()) put(items.data(), items.size() * sizeof(int));
      count = (int)buf.size();
      break;
    }
    case 8: { const float v[6] = {Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv}; put(v, sizeof(v)); count = (int)buf.size(); break; }
    case 9: for (const std::vector<float>* v : {&F->mvScaleFactors, &F->mvInvScaleFactors, &F->mvLevelSigma2, &F->mvInvLevelSigma2}) put(v->data(), v->size() * sizeof(float)); count = (int)buf.size(); break;
    case 10: for (const std::vector<float>* v : {&F->mvScaleFactors_l, &F->mvInvScaleFactors_l, &F->mvLevelSigma2_l, &F->mvInvLevelSigma2_l}) put(v->data(), (size_t)F->mnScaleLevels_l * sizeof(float)); count = (int)buf.size(); break;
    case 11: { const int v[5] = {F->N, F->N_l, F->mnScaleLevels, F->mnScaleLevels_l, F->monoLeft}; put(v, sizeof(v)); count = (int)buf.size(); break; }
    default: return -1;
  }
  if ((int)buf.size() > cap_bytes) return -2;
  if (!buf.empty()) memcpy(out, buf.data(), buf.size());
  return count;
}

// Tracking::TrackWithMotionModelWithLines' two searches on (cur, last).  Every keypoint of `last` carries a map point at
// K^-1 (u, v, 1) * depth[i] in the last camera (= world) frame; the current pose is [I | t] with t = (tx, ty, tz).
// obs0[i] != 0: the map point has no observations (a match does not block the feature).  Outputs: point_of_cur[i] = the
// last-frame keypoint whose map point CurrentFrame.mvpMapPoints[i] holds (-1: none), line_m12 = matches_12.
extern "C" int plviref_track_search(void* h, float th, int b_mono, const float* depth, const unsigned char* obs0, const float* t3,
                                    float nnratio, int check_ori, float line_nnr, int* point_of_cur, int* line_m12, int* n_line_matches) {
  TrackCase* t = (TrackCase*)h;
  Frame& C = *t->cur;
  Frame& L = *t->last;
  const int n1 = L.N;
  t->points.assign(n1, MapPoint());
  L.mvpMapPoints.assign(n1, nullptr);
  L.mvbOutlier.assign(n1, false);
  for (int i = 0; i < n1; i++) {
    MapPoint& m = t->points[i];
    m.mnId = i;
    m.mObs = (obs0 && obs0[i]) ? 0 : 1;
    const float z = depth ? depth[i] : 1.0f;
    m.mWorldPos = vec3((L.mvKeysUn[i].pt.x - t->cam.cx) / t->cam.fx * z, (L.mvKeysUn[i].pt.y - t->cam.cy) / t->cam.fy * z, z);
    m.mDesc = L.mDescriptors.row(i).clone();
    L.mvpMapPoints[i] = &m;
  }
  L.mTcw = cv::Mat::eye(4, 4, CV_32F);
  C.mTcw = cv::Mat::eye(4, 4, CV_32F);
  for (int k = 0; k < 3; k++) C.mTcw.at<float>(k, 3) = t3 ? t3[k] : 0.0f;
  C.mpCamera = &t->cam;
  C.mvpMapPoints.assign(C.N, nullptr);
  ORBmatcher matcher(nnratio, check_ori != 0);
  const int k = matcher.SearchByProjection(C, L, th, b_mono != 0);
  for (int i = 0; i < C.N; i++) point_of_cur[i] = C.mvpMapPoints[i] ? (int)C.mvpMapPoints[i]->mnId : -1;
  std::vector<int> m12;
  *n_line_matches = 0;
  if (L.mDescriptors_Line.rows >= 2 && C.mDescriptors_Line.rows >= 2 && !L.mvKeys_Line.empty() && !C.mvKeys_Line.empty())
    *n_line_matches = LineMatcher::match(L.mDescriptors_Line, C.mDescriptors_Line, line_nnr, m12);
  for (size_t i = 0; i < m12.size(); i++) line_m12[i] = m12[i];
  return k;
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono = false) on a rectified-stereo frame: mvuRight of the
// current frame given, map points at depth[i], current pose [I | t]: covers the mvuRight > 0 window test and the
// bForward / bBackward level ranges (src/ORBmatcher.cc:1982-2047).  K = (fx, fy, cx, cy).
extern "C" int plviref_orb_search_by_projection_frame_stereo(const cv::KeyPoint* keys2, const unsigned char* desc2, const float* uright2, int n2,
                                                              const unsigned char* blocked, const float* bounds, const float* scale_factors,
                                                              int nlevels, const cv::KeyPoint* keys1, int n1, const float* depth,
                                                              const int* flags, const unsigned char* qdesc, const float* K, float mbf,
                                                              const float* t3, float th, int check_ori, int* match_train) {
  set_bounds_and_grid(bounds);
  GeometricCamera cam;
  cam.fx = K[0]; cam.fy = K[1]; cam.cx = K[2]; cam.cy = K[3];
  Frame C, L;
  fill_frame(C, keys2, desc2, n2, scale_factors, nlevels);
  C.mvuRight.assign(uright2, uright2 + n2);
  C.mpCamera = &cam;
  C.mbf = mbf;
  C.mb = mbf / K[0];
  for (int k = 0; k < 3; k++) C.mTcw.at<float>(k, 3) = t3[k];
  MapPoint old;
  for (int i = 0; i < n2; i++) if (blocked && blocked[i]) C.mvpMapPoints[i] = &old;
  L.N = n1;
  L.Nleft = -1;
  L.mvKeysUn.assign(keys1, keys1 + n1);
  L.mvKeys = L.mvKeysUn;
  L.mTcw = cv::Mat::eye(4, 4, CV_32F);
  L.mvbOutlier.assign(n1, false);
  std::vector<MapPoint> mps(n1);
  L.mvpMapPoints.assign(n1, nullptr);
  for (int i = 0; i < n1; i++) {
    MapPoint& m = mps[i];
    m.mnId = i;
    m.mObs = (flags[i] & 2) ? 0 : 1;
    m.mWorldPos = vec3((keys1[i].pt.x - K[2]) / K[0] * depth[i], (keys1[i].pt.y - K[3]) / K[1] * depth[i], depth[i]);
    m.mDesc = desc_mat(qdesc + 32 * (size_t)i, 1);
    if (!(flags[i] & 1)) L.mvpMapPoints[i] = &m;
  }
  ORBmatcher matcher(0.9f, check_ori != 0);
  const int k = matcher.SearchByProjection(C, L, th, false);
  for (int i = 0; i < n2; i++) match_train[i] = (C.mvpMapPoints[i] && C.mvpMapPoints[i] != &old) ? (int)C.mvpMapPoints[i]->mnId : (C.mvpMapPoints[i] ? -3 : -1);
  return k;
}

// ORBmatcher::Fuse(pKF, vpMapPoints, th) on a keyframe with rectified-stereo observations (mvuRight >= 0: the 3-dof
// chi-square gate 7.8 of src/ORBmatcher.cc:1530-1543).  Map point i at K^-1 (uv, 1) * depth[i]; mbf given.
extern "C" int plviref_orb_fuse_stereo(const cv::KeyPoint* keys, const unsigned char* desc, const float* uright, int n, const float* bounds,
                                       const float* scale_factors, const float* inv_sigma2, int nlevels, const float* uv, const float* depth,
                                       const int* level, const int* flags, const unsigned char* qdesc, int nq, const float* K, float mbf,
                                       float th, int* best_idx) {
  set_bounds_and_grid(bounds);
  GeometricCamera cam;
  cam.fx = K[0]; cam.fy = K[1]; cam.cx = K[2]; cam.cy = K[3];
  Frame::fx = K[0]; Frame::fy = K[1]; Frame::cx = K[2]; Frame::cy = K[3]; Frame::invfx = 1.0f / K[0]; Frame::invfy = 1.0f / K[1];
  Map map;
  KeyFrameDatabase db;
  Frame F;
  fill_frame(F, keys, desc, n, scale_factors, nlevels);
  F.mvuRight.assign(uright, uright + n);
  F.mvDepth.assign(n, -1.0f);
  F.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + nlevels);
  F.mvLevelSigma2.resize(nlevels);
  for (int l = 0; l < nlevels; l++) F.mvLevelSigma2[l] = 1.0f / inv_sigma2[l];
  F.mnScaleLevels = nlevels;
  F.mpCamera = &cam;
  F.mbf = mbf;
  F.mb = mbf / K[0];
  KeyFrame* pKF = new KeyFrame(F, &map, &db);
  std::vector<MapPoint> mps(nq);
  std::vector<MapPoint*> ptrs(nq);
  for (int i = 0; i < nq; i++) {
    MapPoint& m = mps[i];
    m.mnId = i;
    m.mBad = (flags[i] & 1) != 0;
    m.mWorldPos = vec3((uv[2 * i] - K[2]) / K[0] * depth[i], (uv[2 * i + 1] - K[3]) / K[1] * depth[i], depth[i]);
    m.mNormal = m.mWorldPos.clone();
    m.mnPredLevel = level[i];
    m.mDesc = desc_mat(qdesc + 32 * (size_t)i, 1);
    ptrs[i] = &m;
  }
  ORBmatcher matcher(0.6f, true);
  const int k = matcher.Fuse(pKF, ptrs, th, false);
  for (int i = 0; i < nq; i++) best_idx[i] = mps[i].mFusedIdx;
  delete pKF;
  return k;
}

// Frame::ComputeStereoMatches_Lines (src/Frame.cc:1408-1529) unmodified: grid fill, LineMatcher::matchGrid, and the
// disparity / overlap / depth filter that follows the search.  out: per left line {disp_s, disp_e, depth_s, depth_e};
// le_out: mvle_l (the normalised image line through the undistorted end points).
extern "C" int plviref_frame_stereo_lines(const cv::line_descriptor::KeyLine* kl, const unsigned char* dl, int nl,
                                          const cv::line_descriptor::KeyLine* kr, const unsigned char* dr, int nr, double inv_width,
                                          double inv_height, float mbf, float* out, double* le_out) {
  Frame F;
  F.mvKeys_Line.assign(kl, kl + nl);
  F.mvKeysUn_Line = F.mvKeys_Line;
  F.mvKeysRight_Line.assign(kr, kr + nr);
  F.N_l = nl;
  F.mDescriptors_Line = desc_mat(dl, nl);
  F.mDescriptorsRight_Line = desc_mat(dr, nr);
  F.inv_width = inv_width;
  F.inv_height = inv_height;
  F.mbf = mbf;
  F.ComputeStereoMatches_Lines();
  int k = 0;
  for (int i = 0; i < nl; i++) {
    out[4 * i] = F.mvDisparity_l[i].first; out[4 * i + 1] = F.mvDisparity_l[i].second;
    out[4 * i + 2] = F.mvDepth_l[i].first; out[4 * i + 3] = F.mvDepth_l[i].second;
    for (int c = 0; c < 3; c++) le_out[3 * i + c] = F.mvle_l[i](c);
    k += F.mvDepth_l[i].first >= 0;
  }
  return k;
}
