// TEST INFRASTRUCTURE (see oracle_common.h).  CPU restatement of the line path:
//   Lineextractor::operator()            src/LineExtractor.cc:45-117
//   LSDDetectorC::ComputePyramid/detect  Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:76-138,254-362
//   LineSegmentDetectorImpl (refine=0)   src/LSD/lsd.cpp:412-782,1136-1152
//   BinaryDescriptor (LBD)               Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:76-118,219-261,351-413,540-688,1027-1373
// and of the OpenCV primitives they call (f64 GaussianBlur/resize, pyrDown, Sobel,
// LineIterator count).
//
// Floating-point conventions of this restatement (the reference build's exact choice of
// float vs double libm overloads and FMA contraction is not knowable without building it):
//   * every product/sum is rounded separately (no FMA contraction);
//   * cos/sin/atan2 of float arguments are evaluated in double and rounded to float;
//   * the f64 7x7 Gaussian uses row-sequential / column-symmetric summation, which
//     reproduces cv2 4.13 to <= 1e-13 (cv2 itself is not reproducible to the bit here).
#include <algorithm>
#include <map>

#include "oracle_common.h"
#include "lsd_gauss_table.h"

namespace plvio {

void resize_linear_u8(const u8* src, int sstride, int sw, int sh, u8* dst, int dstride, int dw, int dh);
void gaussian_blur_u8(const u8* src, int sstride, int w, int h, u8* dst, int dstride, const int* k, int ksize);

static const double kPi = 3.14159265358979323846;
static const double NOTDEF = -1024.0;

struct KeyLine {  // 68 bytes, descriptor_custom.hpp:107-146
  float angle; int class_id; int octave; float pt_x, pt_y; float response; float size;
  float startPointX, startPointY, endPointX, endPointY;
  float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
  float lineLength; int numOfPixels;
};

// ---- OpenCV primitives ---------------------------------------------------------------
// cv::getGaussianKernel(n, sigma, CV_64F).  OpenCV >= 4.x builds it with soft-float
// arithmetic; for the reference's fixed LSD setting (n=7, sigma=0.6/(double)0.8f) the
// values probed from cv2 4.13 are used verbatim (and for the other common lsd_scale settings, see
// lsd_gauss_table.h), otherwise the defining formula.
void gaussian_kernel_f64(int n, double sigma, double* k) {
  // values probed from cv2 4.13 for LSD's sigma = 0.6 / (double)lsd_scale at the common lsd_scale settings
  // (oracle/lsd_gauss_table.h, tools/gen_lsd_gauss.py)
  for (const LsdGaussEntry& e : kLsdGaussTable) {
    float sc;
    memcpy(&sc, &e.scale_bits, 4);
    if (e.n == n && sigma == 0.6 / (double)sc) { memcpy(k, e.k, sizeof(double) * n); return; }
  }
  const double scale2X = -0.5 / (sigma * sigma);
  double sum = 0;
  for (int i = 0; i < n; i++) {
    const double x = i - (n - 1) * 0.5;
    k[i] = std::exp(scale2X * x * x);
    sum += k[i];
  }
  sum = 1. / sum;
  for (int i = 0; i < n; i++) k[i] *= sum;
}

// cv::GaussianBlur(CV_64F, ksize, sigma), BORDER_REFLECT_101 (src/LSD/lsd.cpp:455)
void gaussian_blur_f64(const double* src, int w, int h, double* dst, const double* k, int ksize) {
  const int r = ksize / 2;
  std::vector<double> tmp((size_t)w * h), pad(w + 2 * r);
  for (int y = 0; y < h; y++) {
    const double* row = src + (size_t)y * w;
    for (int x = -r; x < w + r; x++) pad[x + r] = row[reflect101(x, w)];
    double* t = &tmp[(size_t)y * w];
    for (int x = 0; x < w; x++) {
      double s = k[0] * pad[x];
      for (int i = 1; i < ksize; i++) s = s + k[i] * pad[x + i];
      t[x] = s;
    }
  }
  std::vector<const double*> up(r + 1), dn(r + 1);
  for (int y = 0; y < h; y++) {
    for (int i = 0; i <= r; i++) {
      dn[i] = &tmp[(size_t)reflect101(y + i, h) * w];
      up[i] = &tmp[(size_t)reflect101(y - i, h) * w];
    }
    double* d = dst + (size_t)y * w;
    for (int x = 0; x < w; x++) {
      double s = k[r] * dn[0][x];
      for (int i = 1; i <= r; i++) s = s + k[r + i] * (dn[i][x] + up[i][x]);
      d[x] = s;
    }
  }
}

// cv::resize(CV_64F, Size(), fx, fy, INTER_LINEAR): float32 weights applied in double,
// horizontal then vertical (src/LSD/lsd.cpp:457)
static void linear_coeffs_f(int ssize, int dsize, double inv_scale, std::vector<int>& ofs, std::vector<float>& a0,
                            std::vector<float>& a1) {
  ofs.resize(dsize); a0.resize(dsize); a1.resize(dsize);
  const double scale = 1.0 / inv_scale;
  for (int d = 0; d < dsize; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = cv_floor(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    ofs[d] = s; a0[d] = 1.f - f; a1[d] = f;
  }
}

void resize_linear_f64(const double* src, int sw, int sh, double* dst, int dw, int dh, double fx, double fy) {
  // cv::resize turns INTER_LINEAR into INTER_AREA when both inverse scales are exactly 2 ("INTER_AREA (fast) also is
  // equal to INTER_LINEAR"): the 2x2 block is summed in raster order and multiplied by 0.25 -- the same value as the
  // bilinear form up to the order of the additions, i.e. not to the last bit.  Probed on cv2 4.13
  // (tests/test_oracle_vs_cv2.py); sizes whose last block would leave the source are not modelled.
  if (fx == 0.5 && fy == 0.5 && 2 * dw <= sw && 2 * dh <= sh) {
    for (int y = 0; y < dh; y++) {
      const double* r0 = src + (size_t)(2 * y) * sw;
      const double* r1 = r0 + sw;
      for (int x = 0; x < dw; x++) dst[(size_t)y * dw + x] = (((r0[2 * x] + r0[2 * x + 1]) + r1[2 * x]) + r1[2 * x + 1]) * 0.25;
    }
    return;
  }
  std::vector<int> xo, yo;
  std::vector<float> xa0, xa1, ya0, ya1;
  linear_coeffs_f(sw, dw, fx, xo, xa0, xa1);
  linear_coeffs_f(sh, dh, fy, yo, ya0, ya1);
  for (int y = 0; y < dh; y++) {
    const double* r0 = src + (size_t)yo[y] * sw;
    const double* r1 = src + (size_t)std::min(yo[y] + 1, sh - 1) * sw;
    for (int x = 0; x < dw; x++) {
      const int s0 = xo[x], s1 = std::min(s0 + 1, sw - 1);
      const double h0 = r0[s0] * xa0[x] + r0[s1] * xa1[x];
      const double h1 = r1[s0] * xa0[x] + r1[s1] * xa1[x];
      dst[(size_t)y * dw + x] = h0 * ya0[y] + h1 * ya1[y];
    }
  }
}

// cv::pyrDown(u8) to (w/2, h/2): [1,4,6,4,1]^2, (v+128)>>8, REFLECT_101
void pyr_down_u8(const u8* src, int w, int h, u8* dst, int dw, int dh) {
  // horizontal [1,4,6,4,1] at even columns for every source row, then vertical at even rows
  std::vector<int> hrow((size_t)h * dw);
  for (int y = 0; y < h; y++) {
    const u8* row = src + (size_t)y * w;
    int* o = &hrow[(size_t)y * dw];
    for (int x = 0; x < dw; x++) {
      const int c = 2 * x;
      if (c >= 2 && c + 2 < w) o[x] = row[c - 2] + 4 * row[c - 1] + 6 * row[c] + 4 * row[c + 1] + row[c + 2];
      else o[x] = row[reflect101(c - 2, w)] + 4 * row[reflect101(c - 1, w)] + 6 * row[c] + 4 * row[reflect101(c + 1, w)] +
                  row[reflect101(c + 2, w)];
    }
  }
  for (int y = 0; y < dh; y++) {
    const int* r0 = &hrow[(size_t)reflect101(2 * y - 2, h) * dw];
    const int* r1 = &hrow[(size_t)reflect101(2 * y - 1, h) * dw];
    const int* r2 = &hrow[(size_t)(2 * y) * dw];
    const int* r3 = &hrow[(size_t)reflect101(2 * y + 1, h) * dw];
    const int* r4 = &hrow[(size_t)reflect101(2 * y + 2, h) * dw];
    for (int x = 0; x < dw; x++)
      dst[(size_t)y * dw + x] = (u8)((r0[x] + 4 * r1[x] + 6 * r2[x] + 4 * r3[x] + r4[x] + 128) >> 8);
  }
}

// cv::Sobel(u8 -> CV_16S, ksize 3), BORDER_REFLECT_101
void sobel3_s16(const u8* src, int w, int h, short* dx, short* dy) {
  for (int y = 0; y < h; y++) {
    const u8* r0 = src + (size_t)reflect101(y - 1, h) * w;
    const u8* r1 = src + (size_t)y * w;
    const u8* r2 = src + (size_t)reflect101(y + 1, h) * w;
    for (int x = 0; x < w; x++) {
      const int xm = x > 0 ? x - 1 : reflect101(-1, w), xp = x + 1 < w ? x + 1 : reflect101(w, w);
      dx[(size_t)y * w + x] = (short)((r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]));
      dy[(size_t)y * w + x] = (short)((r2[xm] - r0[xm]) + 2 * (r2[x] - r0[x]) + (r2[xp] - r0[xp]));
    }
  }
}

static inline float cosf_d(float a) { return glibcm::cosf(a); }   // the reference's cos(float) / sin(float)
static inline float sinf_d(float a) { return glibcm::sinf(a); }

// ---- LSD (src/LSD/lsd.cpp, refine = LSD_REFINE_NONE) -----------------------------------
struct LsdImage {
  int w, h;
  std::vector<double> img, angles, modgrad;
};

// detect(): u8 -> f64, blur, scale (flsd :438-457), then ll_angle (:536-633)
void lsd_prepare(const u8* src, int stride, int w, int h, double scale, double sigma_scale, double quant,
                 double ang_th, LsdImage& L) {
  std::vector<double> image((size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) image[(size_t)y * w + x] = src[(size_t)y * stride + x];
  const double prec = kPi * ang_th / 180;
  const double rho = quant / std::sin(prec);
  if (scale != 1) {
    const double sigma = (scale < 1) ? (sigma_scale / scale) : sigma_scale;
    const unsigned int hk = (unsigned int)(std::ceil(sigma * std::sqrt(2 * 3.0 * std::log(10.0))));
    const int ksize = 1 + 2 * hk;
    std::vector<double> k(ksize), g((size_t)w * h);
    gaussian_kernel_f64(ksize, sigma, k.data());
    gaussian_blur_f64(image.data(), w, h, g.data(), k.data(), ksize);
    L.w = cv_round(w * scale);
    L.h = cv_round(h * scale);
    L.img.resize((size_t)L.w * L.h);
    resize_linear_f64(g.data(), w, h, L.img.data(), L.w, L.h, scale, scale);
  } else {
    L.w = w; L.h = h; L.img = image;
  }
  const int W = L.w, H = L.h;
  L.angles.assign((size_t)W * H, NOTDEF);
  L.modgrad.assign((size_t)W * H, 0.0);
  const double DEG_TO_RADS = kPi / 180;
  for (int y = 0; y < H - 1; ++y)
    for (int x = 0; x < W - 1; ++x) {
      const size_t a = (size_t)y * W + x;
      const double DA = L.img[a + W + 1] - L.img[a];
      const double BC = L.img[a + 1] - L.img[a + W];
      const double gx = DA + BC, gy = DA - BC;
      const double norm = std::sqrt((gx * gx + gy * gy) / 4);
      L.modgrad[a] = norm;
      if (norm <= rho) L.angles[a] = NOTDEF;
      else L.angles[a] = fast_atan2((float)gx, (float)-gy) * DEG_TO_RADS;
    }
}

static inline bool is_aligned(double a, double theta, double prec) {
  if (a == NOTDEF) return false;
  double n_theta = theta - a;
  if (n_theta < 0) n_theta = -n_theta;
  if (n_theta > (3 * kPi) / 2) {
    n_theta -= (2 * kPi);
    if (n_theta < 0) n_theta = -n_theta;
  }
  return n_theta <= prec;
}

static inline double angle_diff(double a, double b) {
  double diff = a - b;
  while (diff <= -kPi) diff += 2 * kPi;
  while (diff > kPi) diff -= 2 * kPi;
  return std::fabs(diff);
}

struct RegPt { int x, y; };

// rect of src/LSD/lsd.cpp (struct rect, :166-176)
struct LsdRect { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };

struct LsdState {
  const LsdImage& L;
  std::vector<u8> used;
  std::vector<RegPt> reg;
  double LOG_NT;
  explicit LsdState(const LsdImage& l) : L(l), used((size_t)l.w * l.h, 0), reg((size_t)l.w * l.h), LOG_NT(0) {}
};

// region_grow (:635-686): breadth-first over the 8-neighbourhood in (yy, xx) scan order; the region angle is
// re-derived from the float sums after every accepted pixel
static void region_grow(LsdState& S, int sx, int sy, int& reg_size, double& reg_angle, double prec) {
  const LsdImage& L = S.L;
  const int W = L.w, H = L.h;
  const double DEG_TO_RADS = kPi / 180;
  const size_t adx = (size_t)sy * W + sx;
  reg_size = 1;
  S.reg[0] = {sx, sy};
  reg_angle = L.angles[adx];
  float sumdx = (float)std::cos(reg_angle), sumdy = (float)std::sin(reg_angle);
  S.used[adx] = 1;
  for (int i = 0; i < reg_size; ++i) {
    const RegPt rp = S.reg[i];
    const int xx_min = std::max(rp.x - 1, 0), xx_max = std::min(rp.x + 1, W - 1);
    const int yy_min = std::max(rp.y - 1, 0), yy_max = std::min(rp.y + 1, H - 1);
    for (int yy = yy_min; yy <= yy_max; ++yy)
      for (int xx = xx_min; xx <= xx_max; ++xx) {
        const size_t c = (size_t)yy * W + xx;
        if (!S.used[c] && is_aligned(L.angles[c], reg_angle, prec)) {
          S.used[c] = 1;
          S.reg[reg_size++] = {xx, yy};
          const double angle = L.angles[c];
          sumdx += cosf_d((float)angle);
          sumdy += sinf_d((float)angle);
          reg_angle = fast_atan2(sumdy, sumdx) * DEG_TO_RADS;
        }
      }
  }
}

// region2rect (:688-744) + get_theta (:746-782)
static void region2rect(const LsdState& S, int reg_size, double reg_angle, double prec, double p, LsdRect& rec) {
  const LsdImage& L = S.L;
  const int W = L.w;
  const double DEG_TO_RADS = kPi / 180;
  double x = 0, y = 0, sum = 0;
  for (int i = 0; i < reg_size; ++i) {
    const double wgt = L.modgrad[(size_t)S.reg[i].y * W + S.reg[i].x];
    x += (double)S.reg[i].x * wgt;
    y += (double)S.reg[i].y * wgt;
    sum += wgt;
  }
  x /= sum;
  y /= sum;
  double Ixx = 0, Iyy = 0, Ixy = 0;
  for (int i = 0; i < reg_size; ++i) {
    const double wgt = L.modgrad[(size_t)S.reg[i].y * W + S.reg[i].x];
    const double dx = (double)S.reg[i].x - x, dy = (double)S.reg[i].y - y;
    Ixx += dy * dy * wgt;
    Iyy += dx * dx * wgt;
    Ixy -= dx * dy * wgt;
  }
  const double lambda = 0.5 * (Ixx + Iyy - std::sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = (std::fabs(Ixx) > std::fabs(Iyy)) ? (double)fast_atan2((float)(lambda - Ixx), (float)Ixy)
                                                    : (double)fast_atan2((float)Ixy, (float)(lambda - Iyy));
  theta *= DEG_TO_RADS;
  if (angle_diff(theta, reg_angle) > prec) theta += kPi;
  const double dx = std::cos(theta), dy = std::sin(theta);
  double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
  for (int i = 0; i < reg_size; ++i) {
    const double rdx = (double)S.reg[i].x - x, rdy = (double)S.reg[i].y - y;
    const double l = rdx * dx + rdy * dy;
    const double w = -rdx * dy + rdy * dx;
    if (l > l_max) l_max = l;
    else if (l < l_min) l_min = l;
    if (w > w_max) w_max = w;
    else if (w < w_min) w_min = w;
  }
  rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy;
  rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
  rec.width = w_max - w_min;
  rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
  if (rec.width < 1.0) rec.width = 1.0;
}

static inline double dist_sq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
static inline double dist_pts(double x1, double y1, double x2, double y2) { return std::sqrt(dist_sq(x1, y1, x2, y2)); }
static inline double angle_diff_signed(double a, double b) {
  double diff = a - b;
  while (diff <= -kPi) diff += 2 * kPi;
  while (diff > kPi) diff -= 2 * kPi;
  return diff;
}
static inline double rect_density(const LsdRect& r, int reg_size) {
  return (double)reg_size / (dist_pts(r.x1, r.y1, r.x2, r.y2) * r.width);
}

// reduce_region_radius (:831-869): shrink the region around its seed until the rectangle is dense enough; removed
// pixels become available again; the removal swaps the last point in (the order of the list changes)
static bool reduce_region_radius(LsdState& S, int& reg_size, double reg_angle, double prec, double p, LsdRect& rec, double density,
                                 double density_th) {
  const int W = S.L.w;
  const double xc = (double)S.reg[0].x, yc = (double)S.reg[0].y;
  const double r1 = dist_sq(xc, yc, rec.x1, rec.y1), r2 = dist_sq(xc, yc, rec.x2, rec.y2);
  double radSq = r1 > r2 ? r1 : r2;
  while (density < density_th) {
    radSq *= 0.75 * 0.75;
    for (int i = 0; i < reg_size; ++i) {
      if (dist_sq(xc, yc, (double)S.reg[i].x, (double)S.reg[i].y) > radSq) {
        S.used[(size_t)S.reg[i].y * W + S.reg[i].x] = 0;
        std::swap(S.reg[i], S.reg[reg_size - 1]);
        --reg_size;
        --i;
      }
    }
    if (reg_size < 2) return false;
    region2rect(S, reg_size, reg_angle, prec, p, rec);
    density = rect_density(rec, reg_size);
  }
  return true;
}

// refine (:784-829): a sparse rectangle is re-grown from the same seed with the tolerance 2 * (standard deviation of the
// angles near the seed), then shrunk
static bool refine_region(LsdState& S, int& reg_size, double reg_angle, double prec, double p, LsdRect& rec, double density_th) {
  const LsdImage& L = S.L;
  const int W = L.w;
  double density = rect_density(rec, reg_size);
  if (density >= density_th) return true;
  const double xc = (double)S.reg[0].x, yc = (double)S.reg[0].y;
  const double ang_c = L.angles[(size_t)S.reg[0].y * W + S.reg[0].x];
  double sum = 0, s_sum = 0;
  int n = 0;
  for (int i = 0; i < reg_size; ++i) {
    const size_t a = (size_t)S.reg[i].y * W + S.reg[i].x;
    S.used[a] = 0;
    if (dist_pts(xc, yc, S.reg[i].x, S.reg[i].y) < rec.width) {
      const double ang_d = angle_diff_signed(L.angles[a], ang_c);
      sum += ang_d;
      s_sum += ang_d * ang_d;
      ++n;
    }
  }
  const double mean_angle = sum / (double)n;
  const double tau = 2.0 * std::sqrt((s_sum - 2.0 * mean_angle * sum) / (double)n + mean_angle * mean_angle);
  region_grow(S, S.reg[0].x, S.reg[0].y, reg_size, reg_angle, tau);
  if (reg_size < 2) return false;
  region2rect(S, reg_size, reg_angle, prec, p, rec);
  density = rect_density(rec, reg_size);
  if (density < density_th) return reduce_region_radius(S, reg_size, reg_angle, prec, p, rec, density, density_th);
  return true;
}

// log_gamma (:70,134-158) and nfa (:1094-1133), incl. the term "(double(n) + 1)" the vendored code has in the place of
// log_gamma(n + 1)
static double log_gamma_f(double x) {
  if (x > 15.0) return 0.918938533204673 + (x - 0.5) * std::log(x) - x + 0.5 * x * std::log(x * std::sinh(1 / x) + 1 / (810.0 * std::pow(x, 6.0)));
  static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * std::log(x + 5.5) - (x + 5.5);
  double b = 0;
  for (int n = 0; n < 7; ++n) {
    a -= std::log(x + (double)n);
    b += q[n] * std::pow(x, (double)n);
  }
  return a + std::log(b);
}

static bool double_equal_rel(double a, double b) {
  if (a == b) return true;
  const double abs_diff = std::fabs(a - b), aa = std::fabs(a), bb = std::fabs(b);
  double abs_max = (aa > bb) ? aa : bb;
  if (abs_max < 2.2250738585072014e-308) abs_max = 2.2250738585072014e-308;
  return (abs_diff / abs_max) <= (100.0 * 2.220446049250313e-16);
}

static double nfa_value(int n, int k, double p, double LOG_NT) {
  if (n == 0 || k == 0) return -LOG_NT;
  if (n == k) return -LOG_NT - (double)n * std::log10(p);
  const double p_term = p / (1 - p);
  const double log1term = ((double)n + 1) - log_gamma_f((double)k + 1) - log_gamma_f((double)(n - k) + 1) + (double)k * std::log(p) +
                          (double)(n - k) * std::log(1.0 - p);
  double term = std::exp(log1term);
  if (double_equal_rel(term, 0)) {
    if (k > n * p) return -log1term / 2.30258509299404568402 - LOG_NT;
    return -LOG_NT;
  }
  double bin_tail = term;
  const double tolerance = 0.1;
  for (int i = k + 1; i <= n; ++i) {
    const double bin_term = (double)(n - i + 1) / (double)i;
    const double mult_term = bin_term * p_term;
    term *= mult_term;
    bin_tail += term;
    if (bin_term < 1) {
      const double err = term * ((1 - std::pow(mult_term, (double)(n - i + 1))) / (1 - mult_term) - 1);
      if (err < tolerance * std::fabs(-std::log10(bin_tail) - LOG_NT) * bin_tail) break;
    }
  }
  return -std::log10(bin_tail) - LOG_NT;
}

// rect_nfa (:975-1092): scan conversion of the rectangle with the vendored code's integer steps (the quotients of the
// corner differences are INTEGER divisions; two of the tests compare a y with tailp's x), rows outside the image skip
// the step update as well
static double rect_nfa(const LsdState& S, const LsdRect& rec) {
  const LsdImage& L = S.L;
  const int W = L.w, H = L.h;
  int total_pts = 0, alg_pts = 0;
  const double half_width = rec.width / 2.0;
  const double dyhw = rec.dy * half_width, dxhw = rec.dx * half_width;
  struct Edge { int x, y; bool taken; };
  Edge e[4] = {{(int)(rec.x1 - dyhw), (int)(rec.y1 + dxhw), false}, {(int)(rec.x2 - dyhw), (int)(rec.y2 + dxhw), false},
               {(int)(rec.x2 + dyhw), (int)(rec.y2 - dxhw), false}, {(int)(rec.x1 + dyhw), (int)(rec.y1 - dxhw), false}};
  std::sort(e, e + 4, [](const Edge& a, const Edge& b) { return a.x == b.x ? a.y < b.y : a.x < b.x; });
  Edge *min_y = &e[0], *max_y = &e[0];
  for (int i = 1; i < 4; ++i) {
    if (min_y->y > e[i].y) min_y = &e[i];
    if (max_y->y < e[i].y) max_y = &e[i];
  }
  min_y->taken = true;
  Edge* leftmost = nullptr;
  for (int i = 0; i < 4; ++i)
    if (!e[i].taken && (!leftmost || leftmost->x > e[i].x)) leftmost = &e[i];
  leftmost->taken = true;
  Edge* rightmost = nullptr;
  for (int i = 0; i < 4; ++i)
    if (!e[i].taken && (!rightmost || rightmost->x < e[i].x)) rightmost = &e[i];
  rightmost->taken = true;
  Edge* tailp = nullptr;
  for (int i = 0; i < 4; ++i)
    if (!e[i].taken && (!tailp || tailp->x > e[i].x)) tailp = &e[i];
  tailp->taken = true;
  const double flstep = (min_y->y != leftmost->y) ? (min_y->x - leftmost->x) / (min_y->y - leftmost->y) : 0;
  const double slstep = (leftmost->y != tailp->x) ? (leftmost->x - tailp->x) / (leftmost->y - tailp->x) : 0;
  const double frstep = (min_y->y != rightmost->y) ? (min_y->x - rightmost->x) / (min_y->y - rightmost->y) : 0;
  const double srstep = (rightmost->y != tailp->x) ? (rightmost->x - tailp->x) / (rightmost->y - tailp->x) : 0;
  double lstep = flstep, rstep = frstep;
  double left_x = min_y->x, right_x = min_y->x;
  for (int y = min_y->y; y <= max_y->y; ++y) {
    if (y < 0 || y >= H) continue;
    for (int x = (int)left_x; x <= (int)right_x; ++x) {
      if (x < 0 || x >= W) continue;
      ++total_pts;
      if (is_aligned(L.angles[(size_t)y * W + x], rec.theta, rec.prec)) ++alg_pts;
    }
    if (y >= leftmost->y) lstep = slstep;
    if (y >= rightmost->y) rstep = srstep;
    left_x += lstep;
    right_x += rstep;
  }
  return nfa_value(total_pts, alg_pts, rec.p, S.LOG_NT);
}

// rect_improve (:871-973)
static double rect_improve(const LsdState& S, LsdRect& rec, double log_eps) {
  const double delta = 0.5, delta_2 = delta / 2.0;
  double log_nfa = rect_nfa(S, rec);
  if (log_nfa > log_eps) return log_nfa;
  LsdRect r = rec;
  for (int n = 0; n < 5; ++n) {
    r.p /= 2;
    r.prec = r.p * kPi;
    const double v = rect_nfa(S, r);
    if (v > log_nfa) { log_nfa = v; rec = r; }
  }
  if (log_nfa > log_eps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.width -= delta;
      const double v = rect_nfa(S, r);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  if (log_nfa > log_eps) return log_nfa;
  for (int side = 0; side < 2; ++side) {
    r = rec;
    for (int n = 0; n < 5; ++n) {
      if ((r.width - delta) >= 0.5) {
        if (side == 0) { r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2; r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2; }
        else { r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2; r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2; }
        r.width -= delta;
        const double v = rect_nfa(S, r);
        if (v > log_nfa) { rec = r; log_nfa = v; }
      }
    }
    if (log_nfa > log_eps) return log_nfa;
  }
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.p /= 2;
      r.prec = r.p * kPi;
      const double v = rect_nfa(S, r);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  return log_nfa;
}

// flsd (:438-534): seeds in raster order over the interior (the vendored code walks the
// coordinate vector, not the gradient-sorted list); returns Vec4f segments.  refine: LSD_REFINE_NONE 0 / STD 1 / ADV 2
// with the thresholds Lineextractor passes (log_eps 1.0, density_th 0.6, src/LineExtractor.cc:60-63).
void lsd_detect(const LsdImage& L, double scale, double ang_th, std::vector<float>& lines,
                std::vector<int>* region_sizes, int refine, double log_eps, double density_th) {
  const int W = L.w, H = L.h;
  const double prec = kPi * ang_th / 180;
  const double p = ang_th / 180;
  LsdState S(L);
  S.LOG_NT = 5 * (std::log10((double)W) + std::log10((double)H)) / 2 + std::log10(11.0);
  const int min_reg_size = (int)(-S.LOG_NT / std::log10(p));
  lines.clear();
  for (int sy = 0; sy < H - 1; sy++)
    for (int sx = 0; sx < W - 1; sx++) {
      const size_t adx = (size_t)sy * W + sx;
      if (S.used[adx] || L.angles[adx] == NOTDEF) continue;
      int reg_size;
      double reg_angle;
      region_grow(S, sx, sy, reg_size, reg_angle, prec);
      if (reg_size < min_reg_size) continue;
      LsdRect rec;
      region2rect(S, reg_size, reg_angle, prec, p, rec);
      if (refine > 0) {
        if (!refine_region(S, reg_size, reg_angle, prec, p, rec, density_th)) continue;
        if (refine >= 2) {
          const double log_nfa = rect_improve(S, rec, log_eps);
          if (log_nfa <= log_eps) continue;
        }
      }
      if (region_sizes) region_sizes->push_back(reg_size);
      double x1 = rec.x1 + 0.5, y1 = rec.y1 + 0.5, x2 = rec.x2 + 0.5, y2 = rec.y2 + 0.5;
      if (scale != 1) { x1 /= scale; y1 /= scale; x2 /= scale; y2 /= scale; }
      lines.push_back((float)x1); lines.push_back((float)y1); lines.push_back((float)x2); lines.push_back((float)y2);
    }
}

// cv::clipLine(Size2l, Point2l&, Point2l&) (OpenCV imgproc/src/drawing.cpp), the clip cv::LineIterator applies
// when an endpoint lies outside the image (an LSD endpoint in (w-1.5, w) rounds to w).  Returns 0 when the line
// misses the image.  Pinned against cv2.clipLine (tests/test_oracle_vs_cv2.py).
int clip_line(int w, int h, long long& x1, long long& y1, long long& x2, long long& y2) {
  const long long right = w - 1, bottom = h - 1;
  if (w <= 0 || h <= 0) return 0;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (long long)((double)(a - y1) * (x2 - x1) / (y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (long long)((double)(a - y2) * (x2 - x1) / (y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (long long)((double)(a - x1) * (y2 - y1) / (x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (long long)((double)(a - x2) * (y2 - y1) / (x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

// cv::LineIterator(img, pt1, pt2, 8).count for integer endpoints
int line_iterator_count(int w, int h, int ax, int ay, int bx, int by) {
  if ((unsigned)ax >= (unsigned)w || (unsigned)bx >= (unsigned)w || (unsigned)ay >= (unsigned)h || (unsigned)by >= (unsigned)h) {
    long long x1 = ax, y1 = ay, x2 = bx, y2 = by;
    if (!clip_line(w, h, x1, y1, x2, y2)) return 0;
    ax = (int)x1; ay = (int)y1; bx = (int)x2; by = (int)y2;
  }
  return std::max(std::abs(bx - ax), std::abs(by - ay)) + 1;
}

// ---- LSDDetectorC::detectImpl KeyLine assembly (LSDDetector_custom.cpp:304-346) -------
static void make_keylines(const std::vector<float>& seg, int octave, int ow, int oh, float lineScale,
                          double min_length, int& class_counter, std::vector<KeyLine>& out) {
  const float octaveScale = (float)std::pow((double)lineScale, (double)octave);
  for (size_t k = 0; k + 3 < seg.size(); k += 4) {
    float e[4] = {seg[k], seg[k + 1], seg[k + 2], seg[k + 3]};
    if (e[0] < 0) e[0] = 0;
    if (e[0] >= ow) e[0] = (float)ow - 1.0f;
    if (e[2] < 0) e[2] = 0;
    if (e[2] >= ow) e[2] = (float)ow - 1.0f;
    if (e[1] < 0) e[1] = 0;
    if (e[1] >= oh) e[1] = (float)oh - 1.0f;
    if (e[3] < 0) e[3] = 0;
    if (e[3] >= oh) e[3] = (float)oh - 1.0f;
    const float ddx = e[0] - e[2], ddy = e[1] - e[3];
    const double length = (float)std::sqrt((double)ddx * ddx + (double)ddy * ddy);
    if (!(length > min_length)) continue;
    KeyLine kl;
    kl.startPointX = e[0] * octaveScale; kl.startPointY = e[1] * octaveScale;
    kl.endPointX = e[2] * octaveScale; kl.endPointY = e[3] * octaveScale;
    kl.sPointInOctaveX = e[0]; kl.sPointInOctaveY = e[1];
    kl.ePointInOctaveX = e[2]; kl.ePointInOctaveY = e[3];
    kl.lineLength = (float)length;
    // cv::LineIterator(img, Point(pt1), Point(pt2)).count, 8-connected (clipped when a rounded endpoint is outside)
    const int ax = cv_roundf(e[0]), ay = cv_roundf(e[1]), bx = cv_roundf(e[2]), by = cv_roundf(e[3]);
    kl.numOfPixels = line_iterator_count(ow, oh, ax, ay, bx, by);
    kl.angle = glibcm::atan2f(kl.endPointY - kl.startPointY, kl.endPointX - kl.startPointX);   // the reference's atan2(float, float)
    kl.class_id = ++class_counter;
    kl.octave = octave;
    kl.size = (kl.endPointX - kl.startPointX) * (kl.endPointY - kl.startPointY);
    kl.response = kl.lineLength / (float)std::max(ow, oh);
    kl.pt_x = (kl.endPointX + kl.startPointX) / 2;
    kl.pt_y = (kl.endPointY + kl.startPointY) / 2;
    out.push_back(kl);
  }
}

// ---- LBD -----------------------------------------------------------------------------
static const int kComb[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
                                 {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
                                 {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

struct LbdWeights {
  double L[21], G[63];
  LbdWeights() {  // BinaryDescriptor ctor (:219-261); note the integer divisions
    double u = (7 * 3 - 1) / 2;
    double sigma = (7 * 2 + 1) / 2;
    double inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < 21; i++) { const double d = i - u; L[i] = std::exp(d * d * inv); }
    u = (9 * 7 - 1) / 2;
    sigma = u;
    inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < 63; i++) { const double d = i - u; G[i] = std::exp(d * d * inv); }
  }
};

// computeLBD for one line (:1027-1373) + binary packing (:402-413,663-667)
void lbd_descriptor(const KeyLine& kl, const short* dxImg, const short* dyImg, int realWidth, int realHeight,
                    float* desVec /*72*/, u8* bin /*32*/) {
  static const LbdWeights Wt;
  const short heightOfLSP = 63, halfHeight = 31;
  const short imageWidth = (short)(realWidth - 1), imageHeight = (short)(realHeight - 1);
  float band[8][9];
  memset(band, 0, sizeof(band));
  const short lengthOfLSP = (short)kl.numOfPixels;
  const short halfWidth = (short)((lengthOfLSP - 1) / 2);
  const float midX = (float)(0.5 * (kl.sPointInOctaveX + kl.ePointInOctaveX));
  const float midY = (float)(0.5 * (kl.sPointInOctaveY + kl.ePointInOctaveY));
  const float dL0 = cosf_d(kl.angle), dL1 = sinf_d(kl.angle);
  const float dO0 = -dL1, dO1 = dL0;
  float t1 = -dL0 * halfWidth, t2 = dL1 * halfHeight;
  float sCorX0 = (t1 + t2) + midX;
  t1 = -dL1 * halfWidth; t2 = dL0 * halfHeight;
  float sCorY0 = (t1 - t2) + midY;
  for (short hID = 0; hID < heightOfLSP; hID++) {
    float sCorX = sCorX0, sCorY = sCorY0;
    float pL = 0, nL = 0, pO = 0, nO = 0;
    for (short wID = 0; wID < lengthOfLSP; wID++) {
      short tc = (short)std::round(sCorX);
      const short xCor = (tc < 0) ? 0 : (tc > imageWidth) ? imageWidth : tc;
      tc = (short)std::round(sCorY);
      const short yCor = (tc < 0) ? 0 : (tc > imageHeight) ? imageHeight : tc;
      const short dx = dxImg[yCor * realWidth + xCor], dy = dyImg[yCor * realWidth + xCor];
      float a = dx * dL0, b = dy * dL1;
      const float gDL = a + b;
      a = dx * dO0; b = dy * dO1;
      const float gDO = a + b;
      if (gDL > 0) pL += gDL; else nL -= gDL;
      if (gDO > 0) pO += gDO; else nO -= gDO;
      sCorX += dL0;
      sCorY += dL1;
    }
    sCorX0 -= dL1;
    sCorY0 += dL0;
    float c = (float)Wt.G[hID];
    pL = c * pL; nL = c * nL;
    const float pL2 = pL * pL, nL2 = nL * nL;
    pO = c * pO; nO = c * nO;
    const float pO2 = pO * pO, nO2 = nO * nO;
    const float rs[8] = {pL, nL, pL2, nL2, pO, nO, pO2, nO2};
    auto add = [&](int b, float cf) {
      for (int q = 0; q < 8; q++) {
        const bool sq = (q == 2 || q == 3 || q == 6 || q == 7);
        const float term = sq ? (cf * cf) * rs[q] : cf * rs[q];
        band[q][b] += term;
      }
    };
    short bandID = (short)(hID / 7);
    add(bandID, (float)Wt.L[hID % 7 + 7]);
    bandID--;
    if (bandID >= 0) add(bandID, (float)Wt.L[hID % 7 + 14]);
    bandID = bandID + 2;
    if (bandID < 9) add(bandID, (float)Wt.L[hID % 7]);
  }
  const float invN2 = (float)(1.0 / (7 * 2.0)), invN3 = (float)(1.0 / (7 * 3.0));
  for (int b = 0; b < 9; b++) {
    const float invN = (b == 0 || b == 8) ? invN2 : invN3;
    const int d = b * 8;
    float temp, m2, tt;
    temp = band[0][b] * invN; desVec[d] = temp; m2 = band[2][b] * invN; tt = temp * temp; desVec[d + 4] = std::sqrt(m2 - tt);
    temp = band[1][b] * invN; desVec[d + 1] = temp; m2 = band[3][b] * invN; tt = temp * temp; desVec[d + 5] = std::sqrt(m2 - tt);
    temp = band[4][b] * invN; desVec[d + 2] = temp; m2 = band[6][b] * invN; tt = temp * temp; desVec[d + 6] = std::sqrt(m2 - tt);
    temp = band[5][b] * invN; desVec[d + 3] = temp; m2 = band[7][b] * invN; tt = temp * temp; desVec[d + 7] = std::sqrt(m2 - tt);
  }
  float tempM = 0, tempS = 0;
  for (int b = 0; b < 9; b++) {
    const float* v = desVec + 8 * b;
    for (int j = 0; j < 4; j++) { const float s = v[j] * v[j]; tempM += s; }
    for (int j = 4; j < 8; j++) { const float s = v[j] * v[j]; tempS += s; }
  }
  tempM = 1 / std::sqrt(tempM);
  tempS = 1 / std::sqrt(tempS);
  for (int b = 0; b < 9; b++) {
    float* v = desVec + 8 * b;
    for (int j = 0; j < 4; j++) v[j] = v[j] * tempM;
    for (int j = 4; j < 8; j++) v[j] = v[j] * tempS;
  }
  for (int i = 0; i < 72; i++)
    if ((double)desVec[i] > 0.4) desVec[i] = (float)0.4;
  float temp = 0;
  for (int i = 0; i < 72; i++) { const float s = desVec[i] * desVec[i]; temp += s; }
  temp = 1 / std::sqrt(temp);
  for (int i = 0; i < 72; i++) desVec[i] = desVec[i] * temp;
  for (int c = 0; c < 32; c++) {
    const float* f1 = desVec + 8 * kComb[c][0];
    const float* f2 = desVec + 8 * kComb[c][1];
    u8 r = 0;
    for (int i = 0; i < 8; i++)
      if (f1[i] > f2[i]) r += (u8)(1 << i);
    bin[c] = r;
  }
}

}  // namespace plvio

using namespace plvio;

extern "C" {

void plvio_gaussian_kernel_f64(int n, double sigma, double* k) { gaussian_kernel_f64(n, sigma, k); }
void plvio_gaussian_blur_f64(const double* src, int w, int h, double* dst, const double* k, int ksize) {
  gaussian_blur_f64(src, w, h, dst, k, ksize);
}
void plvio_resize_linear_f64(const double* src, int sw, int sh, double* dst, int dw, int dh, double fx, double fy) {
  resize_linear_f64(src, sw, sh, dst, dw, dh, fx, fy);
}
void plvio_pyr_down_u8(const u8* src, int w, int h, u8* dst, int dw, int dh) { pyr_down_u8(src, w, h, dst, dw, dh); }
void plvio_sobel3_s16(const u8* src, int w, int h, short* dx, short* dy) { sobel3_s16(src, w, h, dx, dy); }
int plvio_clip_line(int w, int h, int* x1, int* y1, int* x2, int* y2) {
  long long a = *x1, b = *y1, c = *x2, d = *y2;
  const int r = clip_line(w, h, a, b, c, d);
  *x1 = (int)a; *y1 = (int)b; *x2 = (int)c; *y2 = (int)d;
  return r;
}
int plvio_line_iterator_count(int w, int h, int ax, int ay, int bx, int by) { return line_iterator_count(w, h, ax, ay, bx, by); }

// LSD on one u8 octave image.  Outputs (optional): scaled f64 image, angles, modgrad of
// size *sw x *sh; segments as x1,y1,x2,y2 floats (cap = max segments).  Returns count.
int plvio_lsd(const u8* img, int stride, int w, int h, float lsd_scale, int* sw, int* sh, double* scaled,
              double* angles, double* modgrad, float* segs, int cap, int* region_sizes, int refine) {
  LsdImage L;
  lsd_prepare(img, stride, w, h, (double)lsd_scale, 0.6, 2.0, 22.5, L);
  if (sw) *sw = L.w;
  if (sh) *sh = L.h;
  const size_t n = (size_t)L.w * L.h;
  if (scaled) memcpy(scaled, L.img.data(), n * sizeof(double));
  if (angles) memcpy(angles, L.angles.data(), n * sizeof(double));
  if (modgrad) memcpy(modgrad, L.modgrad.data(), n * sizeof(double));
  std::vector<float> lines;
  std::vector<int> rs;
  lsd_detect(L, (double)lsd_scale, 22.5, lines, &rs, refine, 1.0, 0.6);
  const int m = (int)(lines.size() / 4);
  for (int i = 0; i < std::min(m, cap) * 4; i++) segs[i] = lines[i];
  if (region_sizes) for (int i = 0; i < std::min(m, cap); i++) region_sizes[i] = rs[i];
  return m;
}

void plvio_lbd(const plvio::KeyLine* kl, const short* dx, const short* dy, int w, int h, float* des72, u8* bin32) {
  lbd_descriptor(*kl, dx, dy, w, h, des72, bin32);
}

// Lineextractor::operator() (LSD branch).  keylines/desc/lineeq caller-allocated with
// capacity cap (lineeq: 3 doubles per line).  Returns number of lines (0: descriptors
// untouched, like the reference).  Optional debug: raw per-octave segment counts.
int plvio_line_extract(const u8* img, int w, int h, int stride, int lsd_nfeatures, int lsd_refine,
                       float lsd_scale, int nlevels, float scale, plvio::KeyLine* keylines, u8* desc,
                       double* lineeq, int cap, int* raw_counts) {
  if (nlevels < 1 || nlevels > 2 || lsd_refine < 0 || lsd_refine > 2) return -1;
  // LSDDetectorC::ComputePyramid (:76-109)
  std::vector<std::vector<u8>> pyr(nlevels);
  std::vector<int> pw(nlevels), ph(nlevels);
  std::vector<float> sf(nlevels, 1.0f), isf(nlevels, 1.0f);
  for (int l = 0; l < nlevels; l++) {
    if (l > 0) sf[l] = sf[l - 1] * scale;
    isf[l] = 1.0f / sf[l];
    pw[l] = cv_roundf((float)w * isf[l]);
    ph[l] = cv_roundf((float)h * isf[l]);
    pyr[l].resize((size_t)pw[l] * ph[l]);
    if (l == 0) for (int y = 0; y < h; y++) memcpy(&pyr[0][(size_t)y * w], img + (size_t)y * stride, w);
    else resize_linear_u8(pyr[l - 1].data(), pw[l - 1], pw[l - 1], ph[l - 1], pyr[l].data(), pw[l], pw[l], ph[l]);
  }
  const double min_length = 0.025 * (std::min(w, h));
  std::vector<KeyLine> kls;
  int class_counter = -1;
  for (int l = 0; l < nlevels; l++) {
    LsdImage L;
    lsd_prepare(pyr[l].data(), pw[l], pw[l], ph[l], (double)lsd_scale, 0.6, 2.0, 22.5, L);
    std::vector<float> lines;
    lsd_detect(L, (double)lsd_scale, 22.5, lines, nullptr, lsd_refine, 1.0, 0.6);
    if (raw_counts) raw_counts[l] = (int)(lines.size() / 4);
    make_keylines(lines, l, pw[l], ph[l], scale, min_length, class_counter, kls);
  }
  // top-N by response (src/LineExtractor.cc:75-84).  std::sort there is unstable; this
  // oracle DEFINES ties as "earlier detection first" (stable).
  if ((int)kls.size() > lsd_nfeatures && lsd_nfeatures != 0) {
    std::stable_sort(kls.begin(), kls.end(), [](const KeyLine& a, const KeyLine& b) { return a.response > b.response; });
    kls.resize(lsd_nfeatures);
    for (int i = 0; i < lsd_nfeatures; i++) kls[i].class_id = i;
  }
  const int n = (int)kls.size();
  if (n > cap) return -2;
  if (n == 0) return 0;
  // BinaryDescriptor::computeSobel (:374-399) on its own pyramid (:351-371)
  int maxOct = 0;
  for (auto& k : kls) maxOct = std::max(maxOct, k.octave);
  std::vector<std::vector<short>> dxs(maxOct + 1), dys(maxOct + 1);
  std::vector<int> ow(maxOct + 1), oh(maxOct + 1);
  std::vector<u8> cur((size_t)w * h), tmp;
  {
    std::vector<u8> dense((size_t)w * h);
    for (int y = 0; y < h; y++) memcpy(&dense[(size_t)y * w], img + (size_t)y * stride, w);
    static const int k5[5] = {14, 62, 104, 62, 14};
    gaussian_blur_u8(dense.data(), w, w, h, cur.data(), w, k5, 5);
  }
  ow[0] = w; oh[0] = h;
  for (int o = 0; o <= maxOct; o++) {
    if (o > 0) {
      ow[o] = ow[o - 1] / 2; oh[o] = oh[o - 1] / 2;
      tmp.resize((size_t)ow[o] * oh[o]);
      pyr_down_u8(cur.data(), ow[o - 1], oh[o - 1], tmp.data(), ow[o], oh[o]);
      cur.swap(tmp);
    }
    dxs[o].resize((size_t)ow[o] * oh[o]);
    dys[o].resize((size_t)ow[o] * oh[o]);
    sobel3_s16(cur.data(), ow[o], oh[o], dxs[o].data(), dys[o].data());
  }
  for (int i = 0; i < n; i++) {
    float des[72];
    const int o = kls[i].octave;
    lbd_descriptor(kls[i], dxs[o].data(), dys[o].data(), ow[o], oh[o], des, desc + (size_t)i * 32);
    keylines[i] = kls[i];
    // line equation (src/LineExtractor.cc:106-116): l = sp x ep, normalised by |(l0,l1)|
    const double sx = kls[i].startPointX, sy = kls[i].startPointY, ex = kls[i].endPointX, ey = kls[i].endPointY;
    double l0 = sy * 1.0 - 1.0 * ey, l1 = 1.0 * ex - sx * 1.0, l2 = sx * ey - sy * ex;
    const double nrm = std::sqrt(l0 * l0 + l1 * l1);
    lineeq[3 * i] = l0 / nrm; lineeq[3 * i + 1] = l1 / nrm; lineeq[3 * i + 2] = l2 / nrm;
  }
  return n;
}

}  // extern "C"
