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

// Frame::ComputeStereoMatches_Lines (stereo lines, never reached here) names LineMatcher::matchGrid; LineMatcher.cpp is
// part of libplvi_ref.so (it needs stand-in classes of its own), so this library only carries an aborting stub.
#include "LineMatcher.h"
int LineMatcher::matchGrid(const std::vector<line_2d>&, const cv::Mat&, const GridStructure&, const cv::Mat&,
                           const std::vector<std::pair<double, double>>&, const GridWindow&, std::vector<int>&) {
  fprintf(stderr, "libplvi_ref_frame: LineMatcher::matchGrid is not part of this build\n");
  abort();
}

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
