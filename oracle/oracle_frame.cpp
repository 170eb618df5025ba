// ORACLE - TEST INFRASTRUCTURE ONLY (see oracle_common.h).  CPU restatement of the Frame steps that
// follow the extractors:
//   Frame::UndistortKeyPoints / UndistortKeyLines   src/Frame.cc:1124-1197
//     -> cv::undistortPoints(src, dst, K, distCoeffs, noArray(), P) of OpenCV (not vendored in the
//        reference; 4.x algorithm: normalise, 5 fixed-point iterations of the plumb-bob model in
//        double, project with P).  Pinned against cv2.undistortPoints in tests/test_frame_cpu.py
//        (bit-exact on 5000 points) and against tests/golden/undistort_euroc.npz.
//   Frame::AssignFeaturesToGrid + PosInGrid         src/Frame.cc:644-675,1077-1087
#include <algorithm>
#include <climits>
#include <cmath>
#include <utility>
#include <vector>

#include "oracle_common.h"

namespace plvio {

static void undistort_point(float u, float v, const double* K, const double* k, const double* P, int iters, float* out) {
  const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
  const double ifx = 1. / fx, ify = 1. / fy;
  double x = (u - cx) * ifx, y = (v - cy) * ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < iters; j++) {
    const double r2 = x * x + y * y;
    const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
    if (icdist < 0) {
      x = (u - cx) * ifx;
      y = (v - cy) * ify;
      break;
    }
    const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
    const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
    x = (x0 - deltaX) * icdist;
    y = (y0 - deltaY) * icdist;
  }
  const double xx = P[0] * x + 0.0 * y + P[2];
  const double yy = 0.0 * x + P[1] * y + P[3];
  const double ww = 1. / (0.0 * x + 0.0 * y + 1.0);
  out[0] = (float)(xx * ww);
  out[1] = (float)(yy * ww);
}

}  // namespace plvio

using namespace plvio;

extern "C" {

// xy: n interleaved (x, y) floats; K = (fx, fy, cx, cy); k = 14 distortion coefficients; P = (fx', fy', cx', cy')
void plvio_undistort_points(const float* xy, int n, const double* K, const double* k, const double* P, int iters, float* out) {
  if (k[0] == 0.0) {   // Frame::UndistortKeyPoints: mDistCoef.at<float>(0) == 0.0 -> copy
    for (int i = 0; i < 2 * n; i++) out[i] = xy[i];
    return;
  }
  for (int i = 0; i < n; i++) undistort_point(xy[2 * i], xy[2 * i + 1], K, k, P, iters > 0 ? iters : 5, out + 2 * i);
}

// cv::undistortPoints(src, dst, K, distCoeffs, noArray(), P) itself (no Frame-level shortcut), for the OpenCV stand-in of
// oracle/cvmini (the reference's Frame.cc compiled unmodified calls it): same arithmetic as above.
void plvio_cv_undistort_points(const float* xy, int n, const double* K, const double* k, const double* P, float* out) {
  for (int i = 0; i < n; i++) undistort_point(xy[2 * i], xy[2 * i + 1], K, k, P, 5, out + 2 * i);
}

// mGrid as a CSR: cell (i, j) = items[cell_start[i*48+j] .. cell_start[i*48+j+1])
void plvio_assign_grid(const float* xy, int n, float minX, float minY, float invW, float invH, int* cell_start, int* items) {
  const int COLS = 64, ROWS = 48;
  std::vector<std::vector<int>> cells(COLS * ROWS);
  for (int i = 0; i < n; i++) {
    const int px = (int)std::round((xy[2 * i] - minX) * invW);
    const int py = (int)std::round((xy[2 * i + 1] - minY) * invH);
    if (px < 0 || px >= COLS || py < 0 || py >= ROWS) continue;
    cells[px * ROWS + py].push_back(i);
  }
  int pos = 0;
  for (int c = 0; c < COLS * ROWS; c++) {
    cell_start[c] = pos;
    for (int v : cells[c]) items[pos++] = v;
  }
  cell_start[COLS * ROWS] = pos;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Frame::ComputeStereoMatches (src/Frame.cc:1228-1406): for every left keypoint the best right keypoint of its row
// band by Hamming distance (levels +-1, disparity range), 11x11 SAD refinement over +-5 px on the pyramid level of
// the left keypoint (patches minus their centre pixel, CV_16S, cv::norm NORM_L1), parabola fit, depth = bf /
// disparity, then removal of matches whose SAD is >= 1.5 * 1.4 * median.  pyrL / pyrR: dense level images
// concatenated (level 0 first), lw / lh: level sizes.  An empty match list leaves everything at -1 (the reference
// indexes vDistIdx[0] of an empty vector there: undefined).
// ---------------------------------------------------------------------------------------------------------------
extern "C" int plvio_stereo_matches(const float* kL /* x,y,size,angle,response,octave(int),class_id(int) */, const uint8_t* dL,
                                    int nL, const float* kR, const uint8_t* dR, int nR, const uint8_t* pyrL,
                                    const uint8_t* pyrR, const int* lw, const int* lh, int nlevels, const float* scaleFactors,
                                    const float* invScaleFactors, float mb, float mbf, float* uRight, float* depth) {
  struct KP { float x, y, size, angle, response; int octave, class_id; };
  const KP* L = reinterpret_cast<const KP*>(kL);
  const KP* R = reinterpret_cast<const KP*>(kR);
  std::vector<size_t> off(nlevels + 1, 0);
  for (int l = 0; l < nlevels; l++) off[l + 1] = off[l] + (size_t)lw[l] * lh[l];
  for (int i = 0; i < nL; i++) { uRight[i] = -1.0f; depth[i] = -1.0f; }
  const int TH_HIGH = 100, TH_LOW = 50;
  const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
  const int nRows = lh[0];
  std::vector<std::vector<int>> rows(nRows);
  for (int iR = 0; iR < nR; iR++) {
    const float kpY = R[iR].y;
    const float r = 2.0f * scaleFactors[R[iR].octave];
    const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
    for (int yi = minr; yi <= maxr; yi++)
      if (yi >= 0 && yi < nRows) rows[yi].push_back(iR);
  }
  const float minZ = mb, minD = 0, maxD = mbf / minZ;
  std::vector<std::pair<int, int>> vDistIdx;
  auto hamming = [](const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 32; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
  };
  for (int iL = 0; iL < nL; iL++) {
    const KP& kpL = L[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;
    if (row < 0 || row >= nRows) continue;
    const std::vector<int>& cand = rows[row];
    if (cand.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH;
    int bestIdxR = 0;
    for (int iR : cand) {
      const KP& kpR = R[iR];
      if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
      const float uR = kpR.x;
      if (uR >= minU && uR <= maxU) {
        const int dist = hamming(dL + 32 * (size_t)iL, dR + 32 * (size_t)iR);
        if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
      }
    }
    if (!(bestDist < thOrbDist)) continue;
    const float uR0 = R[bestIdxR].x;
    const float scaleFactor = invScaleFactors[kpL.octave];
    const float scaleduL = std::round(kpL.x * scaleFactor);
    const float scaledvL = std::round(kpL.y * scaleFactor);
    const float scaleduR0 = std::round(uR0 * scaleFactor);
    const int w = 5, Ls = 5;
    const int W = lw[kpL.octave], H = lh[kpL.octave];
    const uint8_t* IL = pyrL + off[kpL.octave];
    const uint8_t* IR = pyrR + off[kpL.octave];
    const int cuL = (int)scaleduL, cvL = (int)scaledvL, cuR = (int)scaleduR0;
    if (cvL - w < 0 || cvL + w >= H || cuL - w < 0 || cuL + w >= W) continue;   // cv::Mat::rowRange / colRange would throw
    const float iniu = scaleduR0 + Ls - w, endu = scaleduR0 + Ls + w + 1;
    if (iniu < 0 || endu >= W) continue;
    if (cuR - Ls - w < 0) continue;                                             // colRange with a negative start would throw
    int bestSad = INT32_MAX, bestincR = 0;
    float vDists[11];
    const int cL = IL[(size_t)cvL * W + cuL];
    for (int incR = -Ls; incR <= Ls; incR++) {
      const int cR = IR[(size_t)cvL * W + cuR + incR];
      int sad = 0;
      for (int dy = -w; dy <= w; dy++)
        for (int dx = -w; dx <= w; dx++) {
          const int a = (int)IL[(size_t)(cvL + dy) * W + cuL + dx] - cL;
          const int b = (int)IR[(size_t)(cvL + dy) * W + cuR + incR + dx] - cR;
          sad += std::abs(a - b);
        }
      const float dist = (float)sad;
      if (dist < bestSad) { bestSad = (int)dist; bestincR = incR; }
      vDists[Ls + incR] = dist;
    }
    if (bestincR == -Ls || bestincR == Ls) continue;
    const float dist1 = vDists[Ls + bestincR - 1], dist2 = vDists[Ls + bestincR], dist3 = vDists[Ls + bestincR + 1];
    const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
    if (deltaR < -1 || deltaR > 1) continue;
    float bestuR = scaleFactors[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
    float disparity = uL - bestuR;
    if (disparity >= minD && disparity < maxD) {
      if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
      depth[iL] = mbf / disparity;
      uRight[iL] = bestuR;
      vDistIdx.push_back(std::pair<int, int>(bestSad, iL));
    }
  }
  if (vDistIdx.empty()) return 0;
  std::sort(vDistIdx.begin(), vDistIdx.end());
  const float median = vDistIdx[vDistIdx.size() / 2].first;
  const float thDist = 1.5f * 1.4f * median;
  int kept = (int)vDistIdx.size();
  for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
    if (vDistIdx[i].first < thDist) break;
    uRight[vDistIdx[i].second] = -1;
    depth[vDistIdx[i].second] = -1;
    kept--;
  }
  return kept;
}

// ---- the disparity / overlap / depth filter that follows the search in Frame::ComputeStereoMatches_Lines
// (src/Frame.cc:1453-1494), Frame::lineSegmentOverlapStereo (:1502-1533), Frame::filterLineSegmentDisparity (:1535-1546)
// and mvle_l (:1495-1500), for one stereo pair.  seg = (startPointX, startPointY, endPointX, endPointY) floats.
namespace plvio {
static double overlap_stereo(double spl_obs, double epl_obs, double spl_proj, double epl_proj) {
  double overlap = 1.f;
  const float lineHorizTh = 0.1;
  if (std::fabs(epl_obs - spl_obs) > lineHorizTh) {
    const double sln = std::min(spl_obs, epl_obs), eln = std::max(spl_obs, epl_obs);
    const double spn = std::min(spl_proj, epl_proj), epn = std::max(spl_proj, epl_proj);
    const double length = eln - spn;
    if ((epn < sln) || (spn > eln)) overlap = 0.f;
    else if ((epn > eln) && (spn < sln)) overlap = eln - sln;
    else overlap = std::min(eln, epn) - std::max(sln, spn);
    if (length > 0.01f) overlap = overlap / length;
    else overlap = 0.f;
    if (overlap > 1.f) overlap = 1.f;
  }
  return overlap;
}
}  // namespace plvio

extern "C" int plvio_line_stereo_depth(const float* seg1, int n1, const float* seg2, int n2, const int* matches12, const float* seg1_un,
                                       float mbf, float* disparity, float* depth, double* le) {
  int k = 0;
  for (int i1 = 0; i1 < n1; i1++) {
    disparity[2 * i1] = disparity[2 * i1 + 1] = -1;
    depth[2 * i1] = depth[2 * i1 + 1] = -1.0f;
    if (le && n2 == 0) le[3 * i1] = le[3 * i1 + 1] = le[3 * i1 + 2] = 0;   // the reference returns before mvle_l is filled (:1419-1420)
    else if (le) {
      const double a0 = seg1_un[4 * i1], a1 = seg1_un[4 * i1 + 1], b0 = seg1_un[4 * i1 + 2], b1 = seg1_un[4 * i1 + 3];
      const double c0 = a1 * 1.0 - 1.0 * b1, c1 = 1.0 * b0 - a0 * 1.0, c2 = a0 * b1 - a1 * b0;   // sp.cross(ep)
      const double nrm = std::sqrt(c0 * c0 + c1 * c1);
      le[3 * i1] = c0 / nrm; le[3 * i1 + 1] = c1 / nrm; le[3 * i1 + 2] = c2 / nrm;
    }
    const int i2 = matches12[i1];
    if (i2 < 0 || i2 >= n2) continue;
    const double xl1 = seg1[4 * i1], yl1 = seg1[4 * i1 + 1], xl2 = seg1[4 * i1 + 2], yl2 = seg1[4 * i1 + 3];
    double xr1 = seg2[4 * i2], yr1 = seg2[4 * i2 + 1], xr2 = seg2[4 * i2 + 2], yr2 = seg2[4 * i2 + 3];
    const double overlap = plvio::overlap_stereo(yl1, yl2, yr1, yr2);
    // the comma initialisers overwrite sp_r first: the second expression already reads the new sp_r (:1468-1469)
    xr1 = (xr1 * (yl1 - yr2) + xr2 * (yr1 - yl1)) / (yr1 - yr2);
    yr1 = yl1;
    xr2 = (xr1 * (yl2 - yr2) + xr2 * (yr1 - yl2)) / (yr1 - yr2);
    yr2 = yl2;
    double disp_s = xl1 - xr1, disp_e = xl2 - xr2;
    const float lsMinDispRatio = 0.7;
    if (std::min(disp_s, disp_e) / std::max(disp_s, disp_e) < lsMinDispRatio) { disp_s = -1.0; disp_e = -1.0; }
    const int minDisp = 1;
    const float lineHorizTh = 0.1, stereoOverlapTh = 0.75;
    if (disp_s >= minDisp && disp_e >= minDisp && std::abs(yl1 - yl2) > lineHorizTh && std::abs(yr1 - yr2) > lineHorizTh &&
        overlap > stereoOverlapTh) {
      disparity[2 * i1] = (float)disp_s; disparity[2 * i1 + 1] = (float)disp_e;
      depth[2 * i1] = mbf / float(disp_s); depth[2 * i1 + 1] = mbf / float(disp_e);
      k++;
    }
  }
  return k;
}
