// TEST INFRASTRUCTURE (see oracle_common.h).  CPU restatement of the Hamming searches:
//   ORBmatcher::DescriptorDistance            src/ORBmatcher.cc:2350-2366
//   ORBmatcher::SearchByProjection(F,F)       src/ORBmatcher.cc:1962-2178  (mono path)
//   ORBmatcher::SearchByProjection(F,MPs)     src/ORBmatcher.cc:44-214     (mono path)
//   ORBmatcher::SearchForInitialization       src/ORBmatcher.cc:706-820
//   ORBmatcher::ComputeThreeMaxima            src/ORBmatcher.cc:2304-2345
//   Frame::AssignFeaturesToGrid/PosInGrid/GetFeaturesInArea  src/Frame.cc:644-675,1006-1087
//   LineMatcher::matchNNR / match / distance / DescriptorDistance  src/LineMatcher.cpp:41-111,173-189,487-499
//   cv::BFMatcher(NORM_HAMMING).knnMatch(k=2): two smallest distances, ties -> lowest train index
#include <algorithm>
#include <climits>

#include "oracle_common.h"

namespace plvio {

static const int GRID_COLS = 64, GRID_ROWS = 48;  // include/Frame.h:47-48
static const int HISTO_LENGTH = 30;

struct Kp {
  float x, y, size, angle, response;
  int octave, class_id;
};

int hamming256(const u8* a, const u8* b) {
  const uint32_t* pa = (const uint32_t*)a;
  const uint32_t* pb = (const uint32_t*)b;
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t v = pa[i] ^ pb[i];
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

// LineMatcher::DescriptorDistance: the >>25 variant (sum of floor(popcount32/2)).
int hamming256_shift25(const u8* a, const u8* b) {
  const uint32_t* pa = (const uint32_t*)a;
  const uint32_t* pb = (const uint32_t*)b;
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t v = pa[i] ^ pb[i];
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 25;
  }
  return dist;
}

struct Grid {
  float minX, minY, invW, invH;
  std::vector<int> cell[GRID_COLS][GRID_ROWS];
  const Kp* keys;
  int n;
};

static void build_grid(Grid& g, const Kp* keys, int n, float minX, float minY, float invW, float invH) {
  g.minX = minX; g.minY = minY; g.invW = invW; g.invH = invH; g.keys = keys; g.n = n;
  for (int i = 0; i < n; i++) {
    int px = (int)std::round((keys[i].x - minX) * invW);
    int py = (int)std::round((keys[i].y - minY) * invH);
    if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
    g.cell[px][py].push_back(i);
  }
}

static void features_in_area(const Grid& g, float x, float y, float r, int minLevel, int maxLevel,
                             std::vector<int>& out) {
  out.clear();
  const int nMinCellX = std::max(0, (int)std::floor((x - g.minX - r) * g.invW));
  if (nMinCellX >= GRID_COLS) return;
  const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - g.minX + r) * g.invW));
  if (nMaxCellX < 0) return;
  const int nMinCellY = std::max(0, (int)std::floor((y - g.minY - r) * g.invH));
  if (nMinCellY >= GRID_ROWS) return;
  const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - g.minY + r) * g.invH));
  if (nMaxCellY < 0) return;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
    for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
      for (int id : g.cell[ix][iy]) {
        const Kp& k = g.keys[id];
        if (bCheckLevels) {
          if (k.octave < minLevel) continue;
          if (maxLevel >= 0 && k.octave > maxLevel) continue;
        }
        const float dx = k.x - x, dy = k.y - y;
        if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(id);
      }
}

static void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

static int rot_bin(float a1, float a2) {
  const float factor = 1.0f / HISTO_LENGTH;  // the reference's quirk: only bins 0..12 are hit
  float rot = a1 - a2;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)std::round(rot * factor);
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

struct Query {  // one projected point
  float u, v, radius;
  int minLevel, maxLevel;
  float angle;
  int flags;  // bit0: skip this query; bit1: its map point has no observations (does not block)
};

}  // namespace plvio

using namespace plvio;

extern "C" {

int plvio_hamming256(const u8* a, const u8* b) { return hamming256(a, b); }
int plvio_hamming256_shift25(const u8* a, const u8* b) { return hamming256_shift25(a, b); }

// Frame::AssignFeaturesToGrid + Frame::GetFeaturesInArea (src/Frame.cc:644-675, 677-763) as a stand-alone query object:
// used by the stand-in Frame of oracle/cvmini/slam_mock_orb.h (the reference's ORBmatcher.cc compiled unmodified; Frame.cc
// itself needs the whole SLAM object graph) and by tests.
void* plvio_grid_create(const Kp* keys, int n, float minX, float minY, float invW, float invH) {
  Grid* g = new Grid();
  build_grid(*g, keys, n, minX, minY, invW, invH);
  return g;
}
void plvio_grid_destroy(void* g) { delete (Grid*)g; }
int plvio_grid_features_in_area(const void* g, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap) {
  std::vector<int> idx;
  features_in_area(*(const Grid*)g, x, y, r, minLevel, maxLevel, idx);
  for (int i = 0; i < (int)idx.size() && i < cap; i++) out[i] = idx[i];
  return (int)idx.size();
}

// SearchByProjection(CurrentFrame, LastFrame, th, bMono=true): queries are the last
// frame's tracked points already projected by the host (u, v, radius=th*scale[octave],
// levels octave-1..octave+1).  match_train[i2] = query index or -1 (mvpMapPoints).
int plvio_search_frame(const Kp* keys, const u8* desc, int n, const u8* blocked, float minX,
                       float minY, float invW, float invH, const Query* q, const u8* qdesc, int nq,
                       int thHigh, int checkOri, int* match_train) {
  Grid g;
  build_grid(g, keys, n, minX, minY, invW, invH);
  std::vector<int> owner(n, -1);
  std::vector<char> blk(n, 0);
  if (blocked) for (int i = 0; i < n; i++) blk[i] = blocked[i];
  std::vector<int> rotHist[HISTO_LENGTH];
  int nmatches = 0;
  std::vector<int> idx;
  for (int i = 0; i < nq; i++) {
    if (q[i].flags & 1) continue;
    features_in_area(g, q[i].u, q[i].v, q[i].radius, q[i].minLevel, q[i].maxLevel, idx);
    if (idx.empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int i2 : idx) {
      if (blk[i2]) continue;
      const int d = hamming256(qdesc + 32 * (size_t)i, desc + 32 * (size_t)i2);
      if (d < bestDist) { bestDist = d; bestIdx2 = i2; }
    }
    if (bestDist <= thHigh) {
      owner[bestIdx2] = i;
      if (!(q[i].flags & 2)) blk[bestIdx2] = 1;
      nmatches++;
      if (checkOri) rotHist[rot_bin(q[i].angle, keys[bestIdx2].angle)].push_back(bestIdx2);
    }
  }
  if (checkOri) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
    for (int b = 0; b < HISTO_LENGTH; b++)
      if (b != i1 && b != i2 && b != i3)
        for (int id : rotHist[b]) { owner[id] = -1; nmatches--; }
  }
  for (int i = 0; i < n; i++) match_train[i] = owner[i];
  return nmatches;
}

// SearchByProjection(F, vpMapPoints, th, ...), mono path: best / second best with the
// level-aware ratio test.  radius = r*th*scale[level]; levels level-1..level.
int plvio_search_mappoints(const Kp* keys, const u8* desc, int n, const u8* blocked, float minX,
                           float minY, float invW, float invH, const Query* q, const u8* qdesc,
                           int nq, int thHigh, float nnratio, int* match_train) {
  Grid g;
  build_grid(g, keys, n, minX, minY, invW, invH);
  std::vector<int> owner(n, -1);
  std::vector<char> blk(n, 0);
  if (blocked) for (int i = 0; i < n; i++) blk[i] = blocked[i];
  int nmatches = 0;
  std::vector<int> idx;
  for (int i = 0; i < nq; i++) {
    if (q[i].flags & 1) continue;
    features_in_area(g, q[i].u, q[i].v, q[i].radius, q[i].minLevel, q[i].maxLevel, idx);
    if (idx.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int i2 : idx) {
      if (blk[i2]) continue;
      const int d = hamming256(qdesc + 32 * (size_t)i, desc + 32 * (size_t)i2);
      if (d < bestDist) {
        bestDist2 = bestDist; bestDist = d; bestLevel2 = bestLevel; bestLevel = keys[i2].octave; bestIdx = i2;
      } else if (d < bestDist2) {
        bestLevel2 = keys[i2].octave; bestDist2 = d;
      }
    }
    if (bestDist <= thHigh) {
      if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
      if (bestLevel != bestLevel2 || bestDist <= nnratio * bestDist2) {
        owner[bestIdx] = i;
        if (!(q[i].flags & 2)) blk[bestIdx] = 1;
        nmatches++;
      }
    }
  }
  for (int i = 0; i < n; i++) match_train[i] = owner[i];
  return nmatches;
}

// SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize): queries are
// ALL F1 keypoints (flags bit0 set for octave > 0), (u,v) = vbPrevMatched, radius =
// windowSize, levels 0..0.  matches12[nq]; prev_matched (u,v pairs) updated in place.
int plvio_search_init(const Kp* keys2, const u8* desc2, int n2, float minX, float minY, float invW,
                      float invH, Query* q, const u8* desc1, int n1, int thLow, float nnratio,
                      int checkOri, int* matches12) {
  Grid g;
  build_grid(g, keys2, n2, minX, minY, invW, invH);
  int nmatches = 0;
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  std::vector<int> matchedDist(n2, INT_MAX), m21(n2, -1);
  std::vector<int> idx;
  for (int i1 = 0; i1 < n1; i1++) {
    if (q[i1].flags & 1) continue;
    features_in_area(g, q[i1].u, q[i1].v, q[i1].radius, q[i1].minLevel, q[i1].maxLevel, idx);
    if (idx.empty()) continue;
    int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
    for (int i2 : idx) {
      const int d = hamming256(desc1 + 32 * (size_t)i1, desc2 + 32 * (size_t)i2);
      if (matchedDist[i2] <= d) continue;
      if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestIdx2 = i2; }
      else if (d < bestDist2) bestDist2 = d;
    }
    if (bestDist <= thLow) {
      if (bestDist < (float)bestDist2 * nnratio) {
        if (m21[bestIdx2] >= 0) { matches12[m21[bestIdx2]] = -1; nmatches--; }
        matches12[i1] = bestIdx2;
        m21[bestIdx2] = i1;
        matchedDist[bestIdx2] = bestDist;
        nmatches++;
        if (checkOri) rotHist[rot_bin(q[i1].angle, keys2[bestIdx2].angle)].push_back(i1);
      }
    }
  }
  if (checkOri) {
    int a = -1, b = -1, c = -1;
    three_maxima(rotHist, HISTO_LENGTH, a, b, c);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == a || i == b || i == c) continue;
      for (int id : rotHist[i])
        if (matches12[id] >= 0) { matches12[id] = -1; nmatches--; }
    }
  }
  for (int i1 = 0; i1 < n1; i1++)
    if (matches12[i1] >= 0) { q[i1].u = keys2[matches12[i1]].x; q[i1].v = keys2[matches12[i1]].y; }
  return nmatches;
}

// SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), mono path (src/ORBmatcher.cc:269-471).
// The caller walks the two DBoW2 feature vectors (host code) and passes, in that order, one
// query per keyframe feature of a common vocabulary node: q.min_level/max_level = [start,end)
// range of the node's frame features inside `items` (the frame's vIndicesF order), flags bit0 =
// no map point / bad.  match_train[i2] = query index or -1 (vpMapPointMatches).
int plvio_search_bow(const Kp* keys, const u8* desc, int n, const int* items, const Query* q, const u8* qdesc,
                     int nq, int thLow, float nnratio, int checkOri, int* match_train) {
  std::vector<int> owner(n, -1);
  std::vector<int> rotHist[HISTO_LENGTH];
  int nmatches = 0;
  for (int i = 0; i < nq; i++) {
    if (q[i].flags & 1) continue;
    int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
    for (int k = q[i].minLevel; k < q[i].maxLevel; k++) {
      const int idx = items[k];
      if (owner[idx] >= 0) continue;
      const int d = hamming256(qdesc + 32 * (size_t)i, desc + 32 * (size_t)idx);
      if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdxF = idx; }
      else if (d < bestDist2) bestDist2 = d;
    }
    if (bestDist1 <= thLow) {
      if ((float)bestDist1 < nnratio * (float)bestDist2) {
        owner[bestIdxF] = i;
        if (checkOri) rotHist[rot_bin(q[i].angle, keys[bestIdxF].angle)].push_back(bestIdxF);
        nmatches++;
      }
    }
  }
  if (checkOri) {
    int a = -1, b = -1, c = -1;
    three_maxima(rotHist, HISTO_LENGTH, a, b, c);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == a || i == b || i == c) continue;
      for (int id : rotHist[i]) { owner[id] = -1; nmatches--; }
    }
  }
  for (int i = 0; i < n; i++) match_train[i] = owner[i];
  return nmatches;
}

// LineMatcher::matchNNR on fresh output: knn-2 + ratio.  Needs n2 >= 2 (the reference
// indexes matches_[idx][1] unconditionally: undefined for fewer train rows; here: no match).
int plvio_match_nnr(const u8* d1, int n1, const u8* d2, int n2, float nnr, int* m12) {
  int matches = 0;
  for (int i = 0; i < n1; i++) {
    m12[i] = -1;
    if (n2 < 2) continue;
    int b0 = INT_MAX, b1 = INT_MAX, i0 = -1;
    for (int j = 0; j < n2; j++) {
      const int d = hamming256(d1 + 32 * (size_t)i, d2 + 32 * (size_t)j);
      if (d < b0) { b1 = b0; b0 = d; i0 = j; }
      else if (d < b1) b1 = d;
    }
    if ((float)b0 < (float)b1 * nnr) { m12[i] = i0; matches++; }
  }
  return matches;
}

// LineMatcher::match(desc1, desc2, nnr, matches_12): both directions + mutual check.
int plvio_line_match(const u8* d1, int n1, const u8* d2, int n2, float nnr, int* m12) {
  std::vector<int> m21(std::max(n2, 1));
  int matches = plvio_match_nnr(d1, n1, d2, n2, nnr, m12);
  plvio_match_nnr(d2, n2, d1, n1, nnr, m21.data());
  for (int i1 = 0; i1 < n1; i1++) {
    int& i2 = m12[i1];
    if (i2 >= 0 && m21[i2] != i1) { i2 = -1; matches--; }
  }
  return matches;
}

}  // extern "C"

// ---- LineMatcher::SerachForInitialize / SearchForTriangulation(KF, KF) (src/LineMatcher.cpp:113-171) with
// Frame/KeyFrame::lineDescriptorMAD (src/Frame.cc:1089-1113, src/KeyFrame.cc:411-435); comparators
// include/LineMatcher.h:56-76 (NN distance ascending, NN12 difference DESCENDING).
// Frame::lineDescriptorMAD / KeyFrame::lineDescriptorMAD on the kNN-2 distances (d0 = nearest, d1 = second nearest of
// every query): 1.4826 x median absolute deviation of d0 and of d1 - d0 (src/Frame.cc:1089-1113).
extern "C" void plvio_line_descriptor_mad(const int* d0, const int* d1, int n, double* nn_mad, double* nn12_mad) {
  struct M { float d0, d1; };
  std::vector<M> lm(n);
  for (int i = 0; i < n; i++) lm[i] = M{(float)d0[i], (float)d1[i]};
  std::vector<M> nn = lm, m12 = lm;
  std::stable_sort(nn.begin(), nn.end(), [](const M& a, const M& b) { return a.d0 < b.d0; });
  const double nn_median = nn[int(nn.size() / 2)].d0;
  for (auto& m : nn) m.d0 = fabsf(m.d0 - nn_median);
  std::stable_sort(nn.begin(), nn.end(), [](const M& a, const M& b) { return a.d0 < b.d0; });
  *nn_mad = 1.4826 * nn[int(nn.size() / 2)].d0;
  std::stable_sort(m12.begin(), m12.end(), [](const M& a, const M& b) { return (a.d1 - a.d0) > (b.d1 - b.d0); });
  const double nn12_median = m12[int(m12.size() / 2)].d1 - m12[int(m12.size() / 2)].d0;
  for (auto& m : m12) m.d0 = fabsf(m.d1 - m.d0 - nn12_median);
  std::stable_sort(m12.begin(), m12.end(), [](const M& a, const M& b) { return a.d0 < b.d0; });
  *nn12_mad = 1.4826 * m12[int(m12.size() / 2)].d0;
}

extern "C" int plvio_line_match_mad(const uint8_t* d1, int n1, const uint8_t* d2, int n2, const uint8_t* has1,
                                    const uint8_t* has2, double factor, int* matches12, double* mad) {
  using namespace plvio;
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  mad[0] = mad[1] = 0.0;
  if (n1 < 1 || n2 < 2) return 0;
  struct M { float d0, d1; int q, t; };
  std::vector<M> lm(n1);
  for (int i = 0; i < n1; i++) {   // knnMatch(k = 2): two smallest distances, ties -> lowest train index
    int b0 = -1, b1 = -1, e0 = 1 << 30, e1 = 1 << 30;
    for (int j = 0; j < n2; j++) {
      const int d = hamming256(d1 + 32 * (size_t)i, d2 + 32 * (size_t)j);
      if (d < e0) { e1 = e0; b1 = b0; e0 = d; b0 = j; }
      else if (d < e1) { e1 = d; b1 = j; }
    }
    (void)b1;
    lm[i] = M{(float)e0, (float)e1, i, b0};
  }
  {
    std::vector<int> e0(n1), e1(n1);
    for (int i = 0; i < n1; i++) { e0[i] = (int)lm[i].d0; e1[i] = (int)lm[i].d1; }
    plvio_line_descriptor_mad(e0.data(), e1.data(), n1, &mad[0], &mad[1]);
  }
  const double th = mad[1] * factor;
  int nmatches = 0;
  for (int i = 0; i < n1; i++) {   // lmatches sorted by queryIdx
    const int qdx = lm[i].q, tdx = lm[i].t;
    if ((has1 && has1[qdx]) || (has2 && has2[tdx])) continue;
    const double dist_12 = lm[i].d1 - lm[i].d0;
    if (dist_12 > th) { matches12[qdx] = tdx; nmatches++; }
  }
  return nmatches;
}

// ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:330-402)
extern "C" int plvio_distinctive_descriptor(const uint8_t* desc, int n) {
  using namespace plvio;
  if (n <= 0) return -1;
  std::vector<std::vector<float>> D(n, std::vector<float>(n, 0.f));
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      const int d = hamming256(desc + 32 * (size_t)i, desc + 32 * (size_t)j);
      D[i][j] = (float)d;
      D[j][i] = (float)d;
    }
  int BestMedian = INT_MAX, BestIdx = 0;
  for (int i = 0; i < n; i++) {
    std::vector<int> v(D[i].begin(), D[i].end());
    std::sort(v.begin(), v.end());
    const int median = v[(size_t)(0.5 * (n - 1))];
    if (median < BestMedian) { BestMedian = median; BestIdx = i; }
  }
  return BestIdx;
}

// ---- ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12) (src/ORBmatcher.cc:823-963), mono path,
// restated on the raw FeatureVectors: fvN_nodes ascending node ids, fvN_start CSR into fvN_feat (feature indices).
// mpN[i] != 0: vpMapPointsN[i] exists and is not bad.  matches12[idx1] = idx2 (the reference stores vpMapPoints2[idx2]).
extern "C" int plvio_search_bow_kfkf(const plvio::Kp* keys1, const uint8_t* desc1, const uint8_t* mp1, int n1, const int* fv1_nodes,
                                     const int* fv1_start, const int* fv1_feat, int nfv1, const plvio::Kp* keys2,
                                     const uint8_t* desc2, const uint8_t* mp2, int n2, const int* fv2_nodes, const int* fv2_start,
                                     const int* fv2_feat, int nfv2, float nnratio, int checkOri, int* matches12) {
  using namespace plvio;
  const int TH_LOW = 50;
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  std::vector<char> matched2(n2, 0);
  std::vector<int> rotHist[HISTO_LENGTH];
  int nmatches = 0;
  int a = 0, b = 0;
  while (a < nfv1 && b < nfv2) {
    if (fv1_nodes[a] == fv2_nodes[b]) {
      for (int p1 = fv1_start[a]; p1 < fv1_start[a + 1]; p1++) {
        const int idx1 = fv1_feat[p1];
        if (!mp1[idx1]) continue;
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int p2 = fv2_start[b]; p2 < fv2_start[b + 1]; p2++) {
          const int idx2 = fv2_feat[p2];
          if (matched2[idx2] || !mp2[idx2]) continue;
          const int dist = hamming256(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 < TH_LOW) {
          if ((float)bestDist1 < nnratio * (float)bestDist2) {
            matches12[idx1] = bestIdx2;
            matched2[bestIdx2] = 1;
            if (checkOri) rotHist[rot_bin(keys1[idx1].angle, keys2[bestIdx2].angle)].push_back(idx1);
            nmatches++;
          }
        }
      }
      a++; b++;
    } else if (fv1_nodes[a] < fv2_nodes[b]) {
      while (a < nfv1 && fv1_nodes[a] < fv2_nodes[b]) a++;     // lower_bound
    } else {
      while (b < nfv2 && fv2_nodes[b] < fv1_nodes[a]) b++;
    }
  }
  if (checkOri) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == i1 || i == i2 || i == i3) continue;
      for (int id : rotHist[i]) { matches12[id] = -1; nmatches--; }
    }
  }
  return nmatches;
}

// The point-to-epipolar-line test of Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:142-156) for a GIVEN F12
// (row-major 3x3 float; the reference builds it as K1^-T [t12]x R12 K2^-1 in front of this test): l = x1' F12,
// dsqr = (l . x2)^2 / (a^2 + b^2) < 3.84 * unc, all in float except the final double comparison.
extern "C" int plvio_epipolar_constrain(float x1, float y1, float x2, float y2, const float* F12, float unc) {
  const float la = x1 * F12[0] + y1 * F12[3] + F12[6];
  const float lb = x1 * F12[1] + y1 * F12[4] + F12[7];
  const float lc = x1 * F12[2] + y1 * F12[5] + F12[8];
  const float num = la * x2 + lb * y2 + lc;
  const float den = la * la + lb * lb;
  if (den == 0) return 0;
  const float dsqr = num * num / den;
  return dsqr < 3.84 * unc;
}

// ---- ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo=false, bCoarse)
// (src/ORBmatcher.cc:965-1206), monocular pinhole path, with Pinhole::epipolarConstrain
// (src/CameraModels/Pinhole.cpp:135-157).  mpN[i] != 0: feature i has a map point.  F12 row-major 3x3 float,
// ep = epipole in image 2, sf2 / sigma2_2 = pKF2->mvScaleFactors / mvLevelSigma2.  matches12[idx1] = idx2 or -1.
extern "C" int plvio_search_triangulation(const plvio::Kp* keys1, const uint8_t* desc1, const uint8_t* mp1, int n1,
                                          const int* fv1_nodes, const int* fv1_start, const int* fv1_feat, int nfv1,
                                          const plvio::Kp* keys2, const uint8_t* desc2, const uint8_t* mp2, int n2,
                                          const int* fv2_nodes, const int* fv2_start, const int* fv2_feat, int nfv2,
                                          const float* F12, float epx, float epy, const float* sf2, const float* sigma2_2,
                                          int coarse, int checkOri, int* matches12) {
  using namespace plvio;
  const int TH_LOW = 50;
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  std::vector<char> vbMatched2(n2, 0);   // never set by the reference
  std::vector<int> rotHist[HISTO_LENGTH];
  int nmatches = 0, a = 0, b = 0;
  while (a < nfv1 && b < nfv2) {
    if (fv1_nodes[a] == fv2_nodes[b]) {
      for (int p1 = fv1_start[a]; p1 < fv1_start[a + 1]; p1++) {
        const int idx1 = fv1_feat[p1];
        if (mp1[idx1]) continue;
        const Kp& kp1 = keys1[idx1];
        int bestDist = TH_LOW, bestIdx2 = -1;
        for (int p2 = fv2_start[b]; p2 < fv2_start[b + 1]; p2++) {
          const int idx2 = fv2_feat[p2];
          if (vbMatched2[idx2] || mp2[idx2]) continue;
          const int dist = hamming256(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
          if (dist > TH_LOW || dist > bestDist) continue;
          const Kp& kp2 = keys2[idx2];
          {   // !bStereo1 && !bStereo2 && !pKF1->mpCamera2
            const float distex = epx - kp2.x;
            const float distey = epy - kp2.y;
            if (distex * distex + distey * distey < 100 * sf2[kp2.octave]) continue;
          }
          const bool ok = coarse != 0 || plvio_epipolar_constrain(kp1.x, kp1.y, kp2.x, kp2.y, F12, sigma2_2[kp2.octave]) != 0;
          if (ok) { bestIdx2 = idx2; bestDist = dist; }
        }
        if (bestIdx2 >= 0) {
          matches12[idx1] = bestIdx2;
          nmatches++;
          if (checkOri) rotHist[rot_bin(kp1.angle, keys2[bestIdx2].angle)].push_back(idx1);
        }
      }
      a++; b++;
    } else if (fv1_nodes[a] < fv2_nodes[b]) {
      while (a < nfv1 && fv1_nodes[a] < fv2_nodes[b]) a++;
    } else {
      while (b < nfv2 && fv2_nodes[b] < fv1_nodes[a]) b++;
    }
  }
  if (checkOri) {
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
      if (i == i1 || i == i2 || i == i3) continue;
      for (int id : rotHist[i]) { matches12[id] = -1; nmatches--; }
    }
  }
  return nmatches;
}


// The per-map-point search shared by ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th) (src/ORBmatcher.cc:1399-1610),
// Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) (:1612-1734) and SearchBySim3 (:1736-1960, both directions)
// [NOT SearchByProjection(KeyFrame*, Scw, ...) (:473-704): that one skips vpMatched[idx] != NULL and claims
// vpMatched[bestIdx] while iterating = plvio_search_frame without the rotation check]: the point is projected by the
// host (u, v, radius = th * mvScaleFactors[nPredictedLevel], levels nPredictedLevel-1 .. nPredictedLevel);
// candidates = KeyFrame::GetFeaturesInArea(u, v, radius) (src/KeyFrame.cc:1200-1244) with the level test applied in
// the loop; optional mono reprojection gate  e2 * mvInvLevelSigma2[kpLevel] > chi2  (Fuse: 5.99, :1546-1552;
// chi2 <= 0: no gate); best = first smallest Hamming distance; accepted when bestDist <= th.  Queries are
// independent (nothing is claimed while iterating); what the caller does with bestIdx (Replace / AddObservation /
// mutual check) is map bookkeeping.  best_idx[q] = feature or -1, best_dist[q] = its distance (256 when no candidate).
extern "C" int plvio_search_in_radius(const plvio::Kp* keys, const uint8_t* desc, int n, float minX, float minY, float invW,
                                      float invH, const plvio::Query* q, const uint8_t* qdesc, int nq,
                                      const float* inv_level_sigma2, double chi2, int th, int* best_idx, int* best_dist) {
  using namespace plvio;
  Grid g;
  build_grid(g, keys, n, minX, minY, invW, invH);
  std::vector<int> idx;
  int found = 0;
  for (int i = 0; i < nq; i++) {
    best_idx[i] = -1;
    best_dist[i] = 256;
    if (q[i].flags & 1) continue;
    features_in_area(g, q[i].u, q[i].v, q[i].radius, -1, -1, idx);
    int bestDist = 256, bestIdx = -1;
    for (int i2 : idx) {
      const Kp& kp = keys[i2];
      if (kp.octave < q[i].minLevel || kp.octave > q[i].maxLevel) continue;
      if (chi2 > 0) {
        const float ex = q[i].u - kp.x, ey = q[i].v - kp.y;
        const float e2 = ex * ex + ey * ey;
        if (e2 * inv_level_sigma2[kp.octave] > chi2) continue;
      }
      const int d = hamming256(qdesc + 32 * (size_t)i, desc + 32 * (size_t)i2);
      if (d < bestDist) { bestDist = d; bestIdx = i2; }
    }
    best_dist[i] = bestDist;
    if (bestIdx >= 0 && bestDist <= th) { best_idx[i] = bestIdx; found++; }
  }
  return found;
}


// The per-map-line search of LineMatcher::Fuse(KeyFrame*, vpMapLines, th) (src/LineMatcher.cpp:373-485): the host
// projects the two endpoints (u1, v1, u2, v2), predicts the level and radius = th * mvScaleFactors[level];
// candidates = KeyFrame::GetLinesInArea(u1, v1, u2, v2, radius) (src/KeyFrame.cc:1170-1198: midpoint distance,
// then the "slope - angle" test as written there), level in [nPredictedLevel - 1, nPredictedLevel], distance =
// LineMatcher::DescriptorDistance (the >> 25 variant, :487-499), first smallest; accepted when <= th_low.
// KeyFrame::GetLinesInArea(x1, y1, x2, y2, r) without level arguments (src/KeyFrame.cc:1170-1198) on 68-byte KeyLine
// records, for the stand-in KeyFrame of oracle/cvmini/slam_mock.h (KeyFrame.cc needs the whole SLAM object graph).
extern "C" int plvio_lines_in_area(const uint8_t* keylines, int n, float x1, float y1, float x2, float y2, float r, int* out) {
  struct KL { float angle; int class_id; int octave; float pt_x, pt_y; float rest[12]; };
  const KL* kl = reinterpret_cast<const KL*>(keylines);
  int k = 0;
  for (int i = 0; i < n; i++) {
    const KL& keyLine = kl[i];
    const float distance = (0.5 * (x1 + x2) - keyLine.pt_x) * (0.5 * (x1 + x2) - keyLine.pt_x) +
                           (0.5 * (y1 + y2) - keyLine.pt_y) * (0.5 * (y1 + y2) - keyLine.pt_y);
    if (distance > r * r) continue;
    const float slope = (y1 - y2) / (x1 - x2) - keyLine.angle;
    if (slope > r * 0.01) continue;
    out[k++] = i;
  }
  return k;
}

// q: 6 floats per query = u1, v1, u2, v2, radius, nPredictedLevel (as float); flags[q] != 0 = skipped.
// keylines: 68-byte KeyLine records (angle at offset 0, octave at 8, pt at 12/16).
extern "C" int plvio_line_fuse_search(const uint8_t* keylines, const uint8_t* desc, int n, const float* q, const uint8_t* flags,
                                      const uint8_t* qdesc, int nq, int th_low, int* best_idx, int* best_dist) {
  using namespace plvio;
  struct KL { float angle; int class_id; int octave; float pt_x, pt_y; float rest[12]; };
  static_assert(sizeof(KL) == 68, "KeyLine layout");
  const KL* kl = reinterpret_cast<const KL*>(keylines);
  int found = 0;
  for (int i = 0; i < nq; i++) {
    best_idx[i] = -1;
    best_dist[i] = INT32_MAX;
    if (flags && flags[i]) continue;
    const float x1 = q[6 * i], y1 = q[6 * i + 1], x2 = q[6 * i + 2], y2 = q[6 * i + 3], r = q[6 * i + 4];
    const int level = (int)q[6 * i + 5];
    int bestDist = INT32_MAX, bestIdx = -1;
    for (int k = 0; k < n; k++) {
      const KL& keyLine = kl[k];
      const float distance = (0.5 * (x1 + x2) - keyLine.pt_x) * (0.5 * (x1 + x2) - keyLine.pt_x) +
                             (0.5 * (y1 + y2) - keyLine.pt_y) * (0.5 * (y1 + y2) - keyLine.pt_y);
      if (distance > r * r) continue;
      const float slope = (y1 - y2) / (x1 - x2) - keyLine.angle;
      if (slope > r * 0.01) continue;
      if (keyLine.octave < level - 1 || keyLine.octave > level) continue;
      const int dist = hamming256_shift25(qdesc + 32 * (size_t)i, desc + 32 * (size_t)k);
      if (dist < bestDist) { bestDist = dist; bestIdx = k; }
    }
    best_dist[i] = bestDist;
    if (bestIdx >= 0 && bestDist <= th_low) { best_idx[i] = bestIdx; found++; }
  }
  return found;
}

// ---- stereo line search of Frame::ComputeStereoMatches_Lines (src/Frame.cc:1421-1448) = grid fill along
// LineIterator (src/LineIterator.cpp:9-52, src/gridStructure.cpp:14-23,49-60) + LineMatcher::matchGrid
// (src/LineMatcher.cpp:191-272).  Restated with one occupancy bitmap per right line (bit x of word y = the line
// passes grid cell (x, y)) in place of the per-cell index lists and the unordered_set of candidates: a right line is
// a candidate of a left line iff its bitmap meets the window of the left start cell or of the left end cell.  The
// candidate order is irrelevant (ties on the best distance fail the ratio test; distances[] / matches_21[] are per
// candidate).  Quirks kept: line_2d holds INT pairs (include/LineMatcher.h:41-42), so the left end points are
// truncated to grid cells before the direction is formed (a left line inside one cell has a NaN direction, which
// passes the |cos| test); cells outside the grid are dropped (GridStructure::at -> out_of_bounds).
// seg = (startPointX, startPointY, endPointX, endPointY) per line; grid_cols, grid_rows <= 64.
static void grid_line_bitmap(double x1, double y1, double x2, double y2, int cols, int rows, uint64_t* occ /*[rows]*/) {
  const bool steep = std::fabs(y2 - y1) > std::fabs(x2 - x1);
  if (steep) { std::swap(x1, y1); std::swap(x2, y2); }
  if (x1 > x2) { std::swap(x1, x2); std::swap(y1, y2); }
  const double dx = x2 - x1, dy = std::fabs(y2 - y1);
  double error = dx / 2.0;
  const int ystep = (y1 < y2) ? 1 : -1;
  int y = (int)y1;
  const int maxX = (int)x2;
  for (int x = (int)x1; x <= maxX; x++) {
    const int cx = steep ? y : x, cy = steep ? x : y;
    if (cx >= 0 && cx < cols && cy >= 0 && cy < rows) occ[cy] |= 1ull << cx;
    error -= dy;
    if (error < 0) { y += ystep; error += dx; }
  }
}

static uint64_t window_cols(int x, int cols, int wl, int wr) {
  const int lo = std::max(0, x - wl), hi = std::min(cols, x + wr + 1);   // [lo, hi)
  if (hi <= lo) return 0;
  const uint64_t upto_hi = hi >= 64 ? ~0ull : ((1ull << hi) - 1);
  return upto_hi & ~((1ull << lo) - 1);
}

extern "C" int plvio_line_match_grid(const float* seg1, const uint8_t* d1, int n1, const float* seg2, const uint8_t* d2,
                                     int n2, double inv_width, double inv_height, int grid_rows, int grid_cols,
                                     int win_left, int win_right, int win_up, int win_down, int* m12) {
  if (grid_rows < 1 || grid_cols < 1 || grid_rows > 64 || grid_cols > 64) return -1;
  std::vector<uint64_t> occ((size_t)std::max(n2, 1) * grid_rows, 0);
  std::vector<double> dir(2 * (size_t)std::max(n2, 1));
  for (int j = 0; j < n2; j++) {
    const float* s = seg2 + 4 * (size_t)j;
    double vx = (s[2] - s[0]) * inv_width, vy = (s[3] - s[1]) * inv_height;
    const double mag = std::sqrt(vx * vx + vy * vy);
    dir[2 * j] = vx / mag;
    dir[2 * j + 1] = vy / mag;
    grid_line_bitmap(s[0] * inv_width, s[1] * inv_height, s[2] * inv_width, s[3] * inv_height, grid_cols, grid_rows,
                     occ.data() + (size_t)j * grid_rows);
  }
  std::vector<int> m21(std::max(n2, 1), -1), dist(std::max(n2, 1), INT_MAX);
  int matches = 0;
  for (int i = 0; i < n1; i++) m12[i] = -1;
  for (int i1 = 0; i1 < n1; i1++) {
    const float* s = seg1 + 4 * (size_t)i1;
    const int sx = (int)(s[0] * inv_width), sy = (int)(s[1] * inv_height);
    const int ex = (int)(s[2] * inv_width), ey = (int)(s[3] * inv_height);
    double vx = ex - sx, vy = ey - sy;
    const double mag = std::sqrt(vx * vx + vy * vy);
    vx /= mag;
    vy /= mag;
    const uint64_t cs = window_cols(sx, grid_cols, win_left, win_right), ce = window_cols(ex, grid_cols, win_left, win_right);
    const int s0 = std::max(0, sy - win_up), s1 = std::min(grid_rows, sy + win_down + 1);
    const int e0 = std::max(0, ey - win_up), e1 = std::min(grid_rows, ey + win_down + 1);
    int best = INT_MAX, best2 = INT_MAX, best_idx = -1;
    for (int i2 = 0; i2 < n2; i2++) {
      const uint64_t* o = occ.data() + (size_t)i2 * grid_rows;
      bool cand = false;
      for (int y = s0; y < s1 && !cand; y++) cand = (o[y] & cs) != 0;
      for (int y = e0; y < e1 && !cand; y++) cand = (o[y] & ce) != 0;
      if (!cand) continue;
      if (std::fabs(vx * dir[2 * i2] + vy * dir[2 * i2 + 1]) < 0.75) continue;
      const int d = hamming256(d1 + 32 * (size_t)i1, d2 + 32 * (size_t)i2);
      if (d < dist[i2]) { dist[i2] = d; m21[i2] = i1; } else continue;
      if (d < best) { best2 = best; best = d; best_idx = i2; }
      else if (d < best2) best2 = d;
    }
    if (best < best2 * 0.9) { m12[i1] = best_idx; matches++; }
  }
  for (int i1 = 0; i1 < n1; i1++) {
    int& i2 = m12[i1];
    if (i2 >= 0 && m21[i2] != i1) { i2 = -1; matches--; }
  }
  return matches;
}
