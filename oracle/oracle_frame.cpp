// ORACLE - TEST INFRASTRUCTURE ONLY (see oracle_common.h).  CPU restatement of the Frame steps that
// follow the extractors:
//   Frame::UndistortKeyPoints / UndistortKeyLines   src/Frame.cc:1124-1197
//     -> cv::undistortPoints(src, dst, K, distCoeffs, noArray(), P) of OpenCV (not vendored in the
//        reference; 4.x algorithm: normalise, 5 fixed-point iterations of the plumb-bob model in
//        double, project with P).  Pinned against cv2.undistortPoints in tests/test_frame_cpu.py
//        (bit-exact on 5000 points) and against tests/golden/undistort_euroc.npz.
//   Frame::AssignFeaturesToGrid + PosInGrid         src/Frame.cc:644-675,1077-1087
#include <cmath>
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
