// Reference-signature LineMatcher members that walk Frame / KeyFrame / MapLine, on top of the C ABI (include/plvi.h).
// Takes the place of the reference's src/LineMatcher.cpp in a SLAM build; the descriptor-only members are inline in
// shim/include/LineMatcher.h.
#include "LineMatcher.h"
#include "KeyFrame.h"

#include <climits>
#include <stdexcept>

namespace ORB_SLAM3 {

const int LineMatcher::TH_HIGH;
const int LineMatcher::TH_LOW;

using plvi_shim::check;
using plvi_shim::MatcherHandle;
using plvi_shim::packed_rows;

// src/LineMatcher.cpp:63-90: the local map lines' descriptors against the frame's, kNN-2 + ratio in both directions,
// mutual matches kept.
int LineMatcher::match(const std::vector<MapLine*>& vpLocalMapLines, Frame& CurrentFrame, float nnr, std::vector<int>& matches_12) {
  const int n1 = (int)vpLocalMapLines.size();
  cv::Mat desc1(n1 > 0 ? n1 : 1, 32, CV_8UC1);
  for (int i = 0; i < n1; ++i) std::memcpy(desc1.ptr(i), vpLocalMapLines[i]->GetDescriptor().ptr(0), 32);
  if (n1 == 0) { matches_12.resize(0, -1); return 0; }
  return match(desc1, CurrentFrame.mDescriptors_Line, nnr, matches_12);
}

namespace {
// kNN-2 with the median-absolute-deviation threshold of Frame / KeyFrame::lineDescriptorMAD (src/Frame.cc:1089-1113):
// pairs (qdx, tdx) in ascending qdx whose distance gap d1 - d0 exceeds factor * nn12_mad
template <class Pair>
int mad_search(const cv::Mat& ldesc1, const cv::Mat& ldesc2, double factor, const std::vector<uint8_t>* has1, const std::vector<uint8_t>* has2,
               std::vector<Pair>& out) {
  out.clear();
  const int n1 = ldesc1.rows, n2 = ldesc2.rows;
  if (n1 <= 0 || n2 < 2) return 0;
  std::vector<int> m12(n1, -1);
  std::vector<uint8_t> ta, tb;
  int nm = 0;
  double mad[2] = {0, 0};
  check(plvi_line_match_mad_host(MatcherHandle::get(), packed_rows(ldesc1, n1, ta), n1, packed_rows(ldesc2, n2, tb), n2,
                                 has1 ? has1->data() : nullptr, has2 ? has2->data() : nullptr, factor, m12.data(), &nm, mad),
        "plvi_line_match_mad");
  for (int i = 0; i < n1; i++)
    if (m12[i] >= 0) out.push_back(Pair(i, m12[i]));
  return nm;
}
}  // namespace

// src/LineMatcher.cpp:113-141
int LineMatcher::SerachForInitialize(Frame& InitialFrame, Frame& CurrentFrame, std::vector<std::pair<int, int> >& LineMatches) {
  return mad_search(InitialFrame.mDescriptors_Line, CurrentFrame.mDescriptors_Line, 0.5, nullptr, nullptr, LineMatches);
}

// src/LineMatcher.cpp:143-171: only pairs in which neither line is a map line yet
int LineMatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<std::pair<size_t, size_t> >& vMatchedPairs) {
  const int n1 = pKF1->mDescriptors_l.rows, n2 = pKF2->mDescriptors_l.rows;
  std::vector<uint8_t> has1(n1 > 0 ? n1 : 1, 0), has2(n2 > 0 ? n2 : 1, 0);
  for (int i = 0; i < n1; i++) has1[i] = pKF1->GetMapLine(i) ? 1 : 0;
  for (int i = 0; i < n2; i++) has2[i] = pKF2->GetMapLine(i) ? 1 : 0;
  return mad_search(pKF1->mDescriptors_l, pKF2->mDescriptors_l, 0.1, &has1, &has2, vMatchedPairs);
}

// src/LineMatcher.cpp:191-272.  The GridStructure arrives filled (Frame::ComputeStereoMatches_Lines walks every right
// line with LineIterator, src/Frame.cc:1427-1440); it is turned into one (first, last) column record per right line and
// grid row -- the cells of a digital line inside one grid row are contiguous -- which is what the kernel keeps in shared
// memory.  A grid whose cells do not have that shape is rejected.
int LineMatcher::matchGrid(const std::vector<line_2d>& lines1, const cv::Mat& desc1, const GridStructure& grid, const cv::Mat& desc2,
                           const std::vector<std::pair<double, double> >& directions2, const GridWindow& w, std::vector<int>& matches_12) {
  if ((int)lines1.size() != desc1.rows) throw std::runtime_error("[matchGrid] Each line needs a corresponding descriptor!");
  const int n1 = desc1.rows, n2 = desc2.rows, rows = grid.rows, cols = grid.cols;
  matches_12.resize(n1, -1);
  if (n1 == 0) return 0;
  std::vector<uint8_t> occ((size_t)(n2 > 0 ? n2 : 1) * rows * 2);
  std::vector<int> cells((size_t)(n2 > 0 ? n2 : 1) * rows, 0);
  for (size_t i = 0; i < occ.size(); i += 2) { occ[i] = 255; occ[i + 1] = 0; }
  GridWindow cellOnly;
  cellOnly.width = std::make_pair(0, 0);
  cellOnly.height = std::make_pair(0, 0);
  std::unordered_set<int> ids;
  for (int x = 0; x < cols; x++)
    for (int y = 0; y < rows; y++) {
      ids.clear();
      grid.get(x, y, cellOnly, ids);
      for (const int i2 : ids) {
        if (i2 < 0 || i2 >= n2) continue;   // the reference skips such entries when it meets them (:239)
        uint8_t* r = &occ[((size_t)i2 * rows + y) * 2];
        if (x < r[0]) r[0] = (uint8_t)x;
        if (x > r[1]) r[1] = (uint8_t)x;
        cells[(size_t)i2 * rows + y]++;
      }
    }
  for (int i2 = 0; i2 < n2; i2++)
    for (int y = 0; y < rows; y++) {
      const uint8_t* r = &occ[((size_t)i2 * rows + y) * 2];
      if (cells[(size_t)i2 * rows + y] && cells[(size_t)i2 * rows + y] != r[1] - r[0] + 1)
        throw std::runtime_error("[matchGrid] grid cells of a line are not contiguous inside a grid row (not filled by getLineCoords)");
    }
  std::vector<int> l1((size_t)n1 * 4);
  for (int i = 0; i < n1; i++) {
    l1[4 * i] = lines1[i].first.first; l1[4 * i + 1] = lines1[i].first.second;
    l1[4 * i + 2] = lines1[i].second.first; l1[4 * i + 3] = lines1[i].second.second;
  }
  std::vector<double> dir((size_t)(n2 > 0 ? n2 : 1) * 2, 0.0);
  for (int i = 0; i < n2 && i < (int)directions2.size(); i++) { dir[2 * i] = directions2[i].first; dir[2 * i + 1] = directions2[i].second; }
  std::vector<int> fresh(n1, -1);
  std::vector<uint8_t> ta, tb;
  int nm = 0;
  check(plvi_line_match_grid_occ_host(MatcherHandle::get(), l1.data(), packed_rows(desc1, n1, ta), n1, occ.data(), dir.data(),
                                      packed_rows(desc2, n2, tb), n2, rows, cols, w.width.first, w.width.second, w.height.first,
                                      w.height.second, fresh.data(), &nm),
        "LineMatcher::matchGrid");
  for (int i = 0; i < n1; i++) matches_12[i] = fresh[i];
  return nm;
}

int LineMatcher::SearchByProjection(Frame&, Frame&, const GridStructure&, const float&, const float&) {
  throw std::runtime_error("LineMatcher::SearchByProjection is dead code in the reference (no call site) and is not provided");
}

// src/LineMatcher.cpp:373-485: every candidate map line is projected (both end points must land inside the image, the
// midpoint's distance and viewing angle are checked), the search over KeyFrame::GetLinesInArea runs as one batch (it
// never skips a feature because of an earlier hit), the bookkeeping follows in the reference's order.
int LineMatcher::Fuse(KeyFrame* pKF, const std::vector<MapLine*>& vpMapLines, const float th) {
  cv::Mat Rcw = pKF->GetRotation();
  cv::Mat tcw = pKF->GetTranslation();
  const float& fx = pKF->fx; const float& fy = pKF->fy; const float& cx = pKF->cx; const float& cy = pKF->cy;
  cv::Mat Ow = pKF->GetCameraCenter();
  std::vector<float> queries;
  std::vector<uint8_t> qdesc;
  std::vector<int> src;
  const int nLines = vpMapLines.size();
  for (int iML = 0; iML < nLines; iML++) {
    MapLine* pML = vpMapLines[iML];
    if (!pML || pML->isBad()) continue;
    Vector6d P = pML->GetWorldPos();
    cv::Mat SP = (cv::Mat_<float>(3, 1) << P(0), P(1), P(2));
    cv::Mat EP = (cv::Mat_<float>(3, 1) << P(3), P(4), P(5));
    const cv::Mat SPc = Rcw * SP + tcw;
    const cv::Mat EPc = Rcw * EP + tcw;
    const float SPcZ = SPc.at<float>(2), EPcZ = EPc.at<float>(2);
    if (SPcZ < 0.0f || EPcZ < 0.0f) continue;
    const float invz1 = 1.0f / SPcZ;
    const float u1 = fx * SPc.at<float>(0) * invz1 + cx;
    const float v1 = fy * SPc.at<float>(1) * invz1 + cy;
    if (u1 < pKF->mnMinX || u1 > pKF->mnMaxX) continue;
    if (v1 < pKF->mnMinY || v1 > pKF->mnMaxY) continue;
    const float invz2 = 1.0f / EPcZ;
    const float u2 = fx * EPc.at<float>(0) * invz2 + cx;
    const float v2 = fy * EPc.at<float>(1) * invz2 + cy;
    if (u2 < pKF->mnMinX || u2 > pKF->mnMaxX) continue;
    if (v2 < pKF->mnMinY || v2 > pKF->mnMaxY) continue;
    const float maxDistance = pML->GetMaxDistanceInvariance();
    const float minDistance = pML->GetMinDistanceInvariance();
    const cv::Mat OM = 0.5 * (SP + EP) - Ow;
    const float dist = cv::norm(OM);
    if (dist < minDistance || dist > maxDistance) continue;
    cv::Mat pn = pML->GetNormal();
    if (OM.dot(pn) < 0.5 * dist) continue;
    const int nPredictedLevel = pML->PredictScale(dist, pKF->mfLogScaleFactor);
    const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
    const float rec[6] = {u1, v1, u2, v2, radius, (float)nPredictedLevel};
    queries.insert(queries.end(), rec, rec + 6);
    const size_t o = qdesc.size();
    qdesc.resize(o + 32);
    std::memcpy(&qdesc[o], pML->GetDescriptor().ptr(0), 32);
    src.push_back(iML);
  }
  const int nq = (int)src.size(), n = (int)pKF->mvKeys_Line.size();
  if (nq == 0 || n == 0) return 0;
  std::vector<int> best(nq, -1), bestDist(nq, INT_MAX);
  std::vector<uint8_t> tmp;
  int nfound = 0;
  check(plvi_line_fuse_search_host(MatcherHandle::get(), reinterpret_cast<const plvi_keyline*>(pKF->mvKeys_Line.data()),
                                   packed_rows(pKF->mDescriptors_l, n, tmp), n, queries.data(), nullptr, qdesc.data(), nq, TH_LOW, best.data(),
                                   bestDist.data(), &nfound),
        "plvi_line_fuse_search");
  int nFused = 0;
  for (int q = 0; q < nq; q++) {
    if (best[q] < 0) continue;
    MapLine* pML = vpMapLines[src[q]];
    if (pML->isBad()) continue;   // a Replace() earlier in this loop may have retired it
    MapLine* pMLinKF = pKF->GetMapLine(best[q]);
    if (pMLinKF) {
      if (!pMLinKF->isBad()) {
        if (pMLinKF->Observations() > pML->Observations()) pML->Replace(pMLinKF);
        else pMLinKF->Replace(pML);
      }
    } else {
      pML->AddObservation(pKF, best[q]);
      pKF->AddMapLine(pML, best[q]);
    }
    nFused++;
  }
  return nFused;
}

}  // namespace ORB_SLAM3
