// TEST INFRASTRUCTURE (see oracle_common.h).  CPU restatement of
// ORBextractor::operator() -- reference src/ORBextractor.cc:70-145,408-878,1059-1177.
#include "oracle_common.h"

#include <algorithm>
#include <list>

#include "../pl_vi_orbslam3_b200/csrc/orb_pattern_table.h"

namespace plvio {

// ---------------------------------------------------------------------------
// OpenCV primitives restated as integer / float32 models
// ---------------------------------------------------------------------------

float fast_atan2(float y, float x) {
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale;
  const float p3 = -0.3258083974640975f * scale;
  const float p5 = 0.1555786518463281f * scale;
  const float p7 = -0.04432655554792128f * scale;
  float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1: 11-bit fixed-point
// coefficients, horizontal pass kept at full precision, vertical pass
// ((b0*(H0>>4))>>16 + (b1*(H1>>4))>>16 + 2) >> 2.   Call sites:
// src/ORBextractor.cc:1165, Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:98.
static void linear_coeffs(int ssize, int dsize, std::vector<int>& ofs, std::vector<short>& a0,
                          std::vector<short>& a1) {
  ofs.resize(dsize);
  a0.resize(dsize);
  a1.resize(dsize);
  double inv_scale = (double)dsize / ssize;
  double scale = 1.0 / inv_scale;
  for (int d = 0; d < dsize; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = cv_floor(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    ofs[d] = s;
    a0[d] = (short)cv_roundf((1.f - f) * 2048.f);
    a1[d] = (short)cv_roundf(f * 2048.f);
  }
}

void resize_linear_u8(const u8* src, int sstride, int sw, int sh, u8* dst, int dstride, int dw,
                      int dh) {
  std::vector<int> xo, yo;
  std::vector<short> xa0, xa1, ya0, ya1;
  linear_coeffs(sw, dw, xo, xa0, xa1);
  linear_coeffs(sh, dh, yo, ya0, ya1);
  std::vector<int> h0(dw), h1(dw);
  for (int y = 0; y < dh; y++) {
    int sy0 = yo[y], sy1 = std::min(sy0 + 1, sh - 1);
    const u8* r0 = src + (size_t)sy0 * sstride;
    const u8* r1 = src + (size_t)sy1 * sstride;
    for (int x = 0; x < dw; x++) {
      int sx0 = xo[x], sx1 = std::min(sx0 + 1, sw - 1);
      h0[x] = r0[sx0] * xa0[x] + r0[sx1] * xa1[x];
      h1[x] = r1[sx0] * xa0[x] + r1[sx1] * xa1[x];
    }
    int b0 = ya0[y], b1 = ya1[y];
    for (int x = 0; x < dw; x++) {
      int v = (((b0 * (h0[x] >> 4)) >> 16) + ((b1 * (h1[x] >> 4)) >> 16) + 2) >> 2;
      dst[(size_t)y * dstride + x] = (u8)v;
    }
  }
}

// cv::GaussianBlur for CV_8UC1 with a fixed-point kernel (sum 256), border
// REFLECT_101: horizontal pass 8.8, vertical pass 16.16, (v + 32768) >> 16.
// 7x7 sigma=2 -> {18,34,48,56,48,34,18} (src/ORBextractor.cc:1115);
// 5x5 sigma=1 -> {14,62,104,62,14} (binary_descriptor_custom.cpp:359).
void gaussian_blur_u8(const u8* src, int sstride, int w, int h, u8* dst, int dstride,
                      const int* k, int ksize) {
  const int r = ksize / 2;
  std::vector<uint32_t> tmp((size_t)w * h);
  std::vector<int> pad(w + 2 * r);               // one row with its REFLECT_101 border
  for (int y = 0; y < h; y++) {
    const u8* row = src + (size_t)y * sstride;
    for (int x = -r; x < w + r; x++) pad[x + r] = row[reflect101(x, w)];
    uint32_t* t = &tmp[(size_t)y * w];
    for (int x = 0; x < w; x++) {
      uint32_t s = 0;
      for (int i = 0; i < ksize; i++) s += k[i] * pad[x + i];
      t[x] = s;
    }
  }
  std::vector<const uint32_t*> rows(ksize);
  for (int y = 0; y < h; y++) {
    for (int i = 0; i < ksize; i++) rows[i] = &tmp[(size_t)reflect101(y + i - r, h) * w];
    u8* d = dst + (size_t)y * dstride;
    for (int x = 0; x < w; x++) {
      uint32_t s = 0;
      for (int i = 0; i < ksize; i++) s += k[i] * rows[i][x];
      d[x] = (u8)((s + 32768u) >> 16);
    }
  }
}

// FAST-9/16 corner score of one pixel: the largest threshold t for which the pixel
// is still a corner (9 contiguous ring pixels all > p+t or all < p-t), i.e.
// max over the 16 arcs of min(d) over both polarities, minus 1.
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

int fast_score(const u8* p, int stride) {
  int d[25];
  int c = p[0];
  for (int k = 0; k < 16; k++) d[k] = c - p[kRingDy[k] * stride + kRingDx[k]];
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int best = -1000;
  for (int k = 0; k < 16; k++) {
    int mn = d[k], mx = d[k];
    for (int j = 1; j < 9; j++) {
      mn = std::min(mn, d[k + j]);
      mx = std::max(mx, d[k + j]);
    }
    best = std::max(best, std::max(mn, -mx));
  }
  return best - 1;
}

struct RawKp {
  float x, y, response;
};

// cv::FAST(roi, kps, th, nonmaxSuppression=true), TYPE_9_16, on a w x h ROI.
// Scores exist only for pixels at distance >= 3 from the ROI edge that are corners
// at `th`; everything else scores 0; a corner is kept when its score is strictly
// larger than its 8 neighbours'.  Output in row-major order, pt relative to the ROI.
void fast_roi(const u8* roi, int stride, int w, int h, int th, std::vector<RawKp>& out) {
  out.clear();
  if (w < 7 || h < 7) return;
  std::vector<int> sc((size_t)w * h, 0);
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const u8* p = roi + (size_t)y * stride + x;
      // any arc of 9 ring pixels holds at least 2 of the 4 compass pixels: cheap necessary test
      const int c = p[0], hi = c + th, lo = c - th;
      const int v0 = p[3 * stride], v8 = p[-3 * stride], v4 = p[3], v12 = p[-3];
      const int nb = (v0 > hi) + (v8 > hi) + (v4 > hi) + (v12 > hi);
      const int nd = (v0 < lo) + (v8 < lo) + (v4 < lo) + (v12 < lo);
      if (nb < 2 && nd < 2) continue;
      int s = fast_score(p, stride);
      if (s >= th) sc[(size_t)y * w + x] = s;
    }
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      int s = sc[(size_t)y * w + x];
      if (s == 0) continue;
      bool keep = true;
      for (int dy = -1; dy <= 1 && keep; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          if (!dx && !dy) continue;
          if (sc[(size_t)(y + dy) * w + x + dx] >= s) { keep = false; break; }
        }
      if (keep) out.push_back({(float)x, (float)y, (float)s});
    }
}

// ---------------------------------------------------------------------------
// Reference-owned logic
// ---------------------------------------------------------------------------

struct OrbPlan {
  int nlevels;
  std::vector<float> scale, inv_scale;
  std::vector<int> w, h, quota;
  int umax[16];
};

// ORBextractor::ORBextractor (src/ORBextractor.cc:408-468) + level sizes of
// ComputePyramid (:1156-1157).
void make_plan(int w, int h, int nfeatures, float scaleFactor, int nlevels, OrbPlan& p) {
  p.nlevels = nlevels;
  p.scale.assign(nlevels, 1.0f);
  p.inv_scale.assign(nlevels, 1.0f);
  for (int i = 1; i < nlevels; i++) p.scale[i] = p.scale[i - 1] * scaleFactor;
  for (int i = 0; i < nlevels; i++) p.inv_scale[i] = 1.0f / p.scale[i];
  p.w.resize(nlevels);
  p.h.resize(nlevels);
  for (int i = 0; i < nlevels; i++) {
    p.w[i] = cv_roundf((float)w * p.inv_scale[i]);
    p.h[i] = cv_roundf((float)h * p.inv_scale[i]);
  }
  p.quota.resize(nlevels);
  float factor = 1.0f / scaleFactor;
  float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int l = 0; l < nlevels - 1; l++) {
    p.quota[l] = cv_roundf(nDesired);
    sum += p.quota[l];
    nDesired *= factor;
  }
  p.quota[nlevels - 1] = std::max(nfeatures - sum, 0);

  const int HP = 15;
  int v, v0, vmax = cv_floor(HP * std::sqrt(2.f) / 2 + 1);
  int vmin = cv_ceil(HP * std::sqrt(2.f) / 2);
  const double hp2 = HP * HP;
  for (v = 0; v <= vmax; ++v) p.umax[v] = cv_round(std::sqrt(hp2 - v * v));
  for (v = HP, v0 = 0; v >= vmin; --v) {
    while (p.umax[v0] == p.umax[v0 + 1]) ++v0;
    p.umax[v] = v0;
    ++v0;
  }
}

// Grid FAST of ComputeKeyPointsOctTree (src/ORBextractor.cc:763-855): candidates in
// coordinates relative to (minBorderX, minBorderY), reference emission order.
void grid_fast(const u8* img, int stride, int w, int h, int iniTh, int minTh,
               std::vector<RawKp>& out) {
  const float W = 30;
  const int minBX = 16, minBY = 16, maxBX = w - 16, maxBY = h - 16;
  const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
  const int nCols = (int)(width / W), nRows = (int)(height / W);
  out.clear();
  if (nCols <= 0 || nRows <= 0) return;
  const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
  std::vector<RawKp> cell;
  for (int i = 0; i < nRows; i++) {
    const float iniY = (float)(minBY + i * hCell);
    float maxY = iniY + hCell + 6;
    if (iniY >= maxBY - 3) continue;
    if (maxY > maxBY) maxY = (float)maxBY;
    for (int j = 0; j < nCols; j++) {
      const float iniX = (float)(minBX + j * wCell);
      float maxX = iniX + wCell + 6;
      if (iniX >= maxBX - 6) continue;
      if (maxX > maxBX) maxX = (float)maxBX;
      int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
      const u8* roi = img + (size_t)y0 * stride + x0;
      fast_roi(roi, stride, cw, ch, iniTh, cell);
      if (cell.empty()) fast_roi(roi, stride, cw, ch, minTh, cell);
      for (auto& k : cell) out.push_back({k.x + j * wCell, k.y + i * hCell, k.response});
    }
  }
}

// DistributeOctTree (src/ORBextractor.cc:537-761) + ExtractorNode::DivideNode
// (:479-535).  The reference breaks ties between equally populated nodes by heap
// address (sort of pair<int,ExtractorNode*>, :682), which is not reproducible; this
// oracle DEFINES the tie-break as node creation order (a later-created node sorts
// higher).  Everything else (list order, push_front, early break) follows the source.
struct OctNode {
  int ulx, uly, urx, bry;  // UL.x, UL.y, UR.x(=BR.x), BR.y(=BL.y)
  std::vector<int> keys;   // indices into the candidate array, input order kept
  bool no_more = false;
  int seq = 0;
  std::list<OctNode>::iterator lit;
};

static void divide_node(const OctNode& n, const std::vector<RawKp>& c, OctNode ch[4]) {
  const int halfX = (int)std::ceil((float)(n.urx - n.ulx) / 2);
  const int halfY = (int)std::ceil((float)(n.bry - n.uly) / 2);
  const int mx = n.ulx + halfX, my = n.uly + halfY;
  ch[0] = OctNode{n.ulx, n.uly, mx, my};
  ch[1] = OctNode{mx, n.uly, n.urx, my};
  ch[2] = OctNode{n.ulx, my, mx, n.bry};
  ch[3] = OctNode{mx, my, n.urx, n.bry};
  for (int id : n.keys) {
    const RawKp& kp = c[id];
    if (kp.x < (float)mx) {
      if (kp.y < (float)my) ch[0].keys.push_back(id);
      else ch[2].keys.push_back(id);
    } else if (kp.y < (float)my) ch[1].keys.push_back(id);
    else ch[3].keys.push_back(id);
  }
  for (int q = 0; q < 4; q++) ch[q].no_more = ch[q].keys.size() == 1;
}

void distribute_octree(const std::vector<RawKp>& c, int minX, int maxX, int minY, int maxY, int N,
                       std::vector<int>& result) {
  result.clear();
  if (c.empty()) return;  // the reference is never called usefully with no keys
  const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
  if (nIni < 1) return;   // reference divides by zero here (portrait images): undefined
  const float hX = (float)(maxX - minX) / nIni;
  std::list<OctNode> L;
  std::vector<OctNode*> ini(nIni);
  int seq = 0;
  for (int i = 0; i < nIni; i++) {
    OctNode n{(int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), maxY - minY};
    n.seq = seq++;
    L.push_back(n);
    ini[i] = &L.back();
  }
  for (size_t i = 0; i < c.size(); i++) {
    int idx = (int)(c[i].x / hX);
    if (idx >= nIni) idx = nIni - 1;  // cannot happen for in-range x; guards UB
    ini[idx]->keys.push_back((int)i);
  }
  for (auto it = L.begin(); it != L.end();) {
    if (it->keys.size() == 1) { it->no_more = true; ++it; }
    else if (it->keys.empty()) it = L.erase(it);
    else ++it;
  }
  typedef std::pair<int, OctNode*> SP;
  auto by_size_then_seq = [](const SP& a, const SP& b) {
    return a.first != b.first ? a.first < b.first : a.second->seq < b.second->seq;
  };
  auto push_children = [&](OctNode ch[4], std::vector<SP>& cand, int* nToExpand) {
    for (int q = 0; q < 4; q++) {
      if (ch[q].keys.empty()) continue;
      ch[q].seq = seq++;
      L.push_front(ch[q]);
      if (ch[q].keys.size() > 1) {
        if (nToExpand) ++*nToExpand;
        cand.push_back(SP((int)ch[q].keys.size(), &L.front()));
        L.front().lit = L.begin();
      }
    }
  };
  bool finish = false;
  std::vector<SP> cand;
  while (!finish) {
    int prev = (int)L.size();
    int nToExpand = 0;
    cand.clear();
    for (auto it = L.begin(); it != L.end();) {
      if (it->no_more) { ++it; continue; }
      OctNode ch[4];
      divide_node(*it, c, ch);
      push_children(ch, cand, &nToExpand);
      it = L.erase(it);
    }
    if ((int)L.size() >= N || (int)L.size() == prev) {
      finish = true;
    } else if ((int)L.size() + nToExpand * 3 > N) {
      while (!finish) {
        prev = (int)L.size();
        std::vector<SP> pc = cand;
        cand.clear();
        std::sort(pc.begin(), pc.end(), by_size_then_seq);
        for (int j = (int)pc.size() - 1; j >= 0; j--) {
          OctNode ch[4];
          divide_node(*pc[j].second, c, ch);
          push_children(ch, cand, nullptr);
          L.erase(pc[j].second->lit);
          if ((int)L.size() >= N) break;
        }
        if ((int)L.size() >= N || (int)L.size() == prev) finish = true;
      }
    }
  }
  for (auto& n : L) {
    int best = n.keys[0];
    float mr = c[best].response;
    for (size_t k = 1; k < n.keys.size(); k++)
      if (c[n.keys[k]].response > mr) { best = n.keys[k]; mr = c[best].response; }
    result.push_back(best);
  }
}

// IC_Angle (src/ORBextractor.cc:75-102) on the unblurred level image.
float ic_angle(const u8* img, int stride, float px, float py, const int* umax) {
  int m01 = 0, m10 = 0;
  const u8* center = img + (size_t)cv_roundf(py) * stride + cv_roundf(px);
  for (int u = -15; u <= 15; ++u) m10 += u * center[u];
  for (int v = 1; v <= 15; ++v) {
    int v_sum = 0, d = umax[v];
    for (int u = -d; u <= d; ++u) {
      int vp = center[u + v * stride], vm = center[u - v * stride];
      v_sum += vp - vm;
      m10 += u * (vp + vm);
    }
    m01 += v * v_sum;
  }
  return fast_atan2((float)m01, (float)m10);
}

static const signed char kPattern[1024] = PLVI_ORB_PATTERN_VALUES;

// computeOrbDescriptor (src/ORBextractor.cc:106-145) on the blurred level image.
// float32 throughout, products and sum rounded separately (no FMA contraction).
void orb_descriptor(const u8* img, int stride, float px, float py, float angle_deg, u8* desc) {
  const float factorPI = (float)(3.14159265358979323846 / 180.f);
  float angle = angle_deg * factorPI;
  float a = glibcm::cosf(angle), b = glibcm::sinf(angle);
  const u8* center = img + (size_t)cv_roundf(py) * stride + cv_roundf(px);
  for (int i = 0; i < 32; i++) {
    int val = 0;
    for (int bit = 0; bit < 8; bit++) {
      const signed char* q = kPattern + (i * 8 + bit) * 4;
      volatile float xb0 = q[0] * b, ya0 = q[1] * a, xa0 = q[0] * a, yb0 = q[1] * b;
      volatile float xb1 = q[2] * b, ya1 = q[3] * a, xa1 = q[2] * a, yb1 = q[3] * b;
      int t0 = center[cv_roundf(xb0 + ya0) * stride + cv_roundf(xa0 - yb0)];
      int t1 = center[cv_roundf(xb1 + ya1) * stride + cv_roundf(xa1 - yb1)];
      val |= (t0 < t1) << bit;
    }
    desc[i] = (u8)val;
  }
}

}  // namespace plvio

// ---------------------------------------------------------------------------
// C entry points (ctypes)
// ---------------------------------------------------------------------------
using namespace plvio;

extern "C" {

struct plvio_keypoint {  // cv::KeyPoint POD layout, 28 bytes
  float x, y, size, angle, response;
  int octave, class_id;
};

float plvio_fast_atan2(float y, float x) { return fast_atan2(y, x); }

// restated host-libm float functions (oracle_common.h, namespace glibcm), vectorised for the tests
void plvio_glibc_sincosf(const float* x, int n, float* s, float* c) {
  for (int i = 0; i < n; i++) { s[i] = glibcm::sinf(x[i]); c[i] = glibcm::cosf(x[i]); }
}
void plvio_glibc_atan2f(const float* y, const float* x, int n, float* r) {
  for (int i = 0; i < n; i++) r[i] = glibcm::atan2f(y[i], x[i]);
}
// the same functions of the libm this library is linked against
void plvio_host_sincosf(const float* x, int n, float* s, float* c) {
  for (int i = 0; i < n; i++) { s[i] = ::sinf(x[i]); c[i] = ::cosf(x[i]); }
}
void plvio_host_atan2f(const float* y, const float* x, int n, float* r) {
  for (int i = 0; i < n; i++) r[i] = ::atan2f(y[i], x[i]);
}

int plvio_orb_plan(int w, int h, int nfeatures, float sf, int nlevels, int* lw, int* lh,
                   float* scale, int* quota, int* umax) {
  OrbPlan p;
  make_plan(w, h, nfeatures, sf, nlevels, p);
  for (int i = 0; i < nlevels; i++) {
    lw[i] = p.w[i]; lh[i] = p.h[i]; scale[i] = p.scale[i]; quota[i] = p.quota[i];
  }
  if (umax) for (int i = 0; i < 16; i++) umax[i] = p.umax[i];
  return 0;
}

void plvio_resize_linear_u8(const u8* src, int sstride, int sw, int sh, u8* dst, int dstride,
                            int dw, int dh) {
  resize_linear_u8(src, sstride, sw, sh, dst, dstride, dw, dh);
}

void plvio_gaussian_blur7_u8(const u8* src, int sstride, int w, int h, u8* dst, int dstride) {
  static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
  gaussian_blur_u8(src, sstride, w, h, dst, dstride, k, 7);
}

void plvio_gaussian_blur5_u8(const u8* src, int sstride, int w, int h, u8* dst, int dstride) {
  static const int k[5] = {14, 62, 104, 62, 14};
  gaussian_blur_u8(src, sstride, w, h, dst, dstride, k, 5);
}

void plvio_fast_score_map(const u8* img, int stride, int w, int h, int* out) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      out[(size_t)y * w + x] =
          (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) ? 0 : fast_score(img + (size_t)y * stride + x, stride);
}

// returns count; out = [x, y, response] triples (float)
int plvio_fast_roi(const u8* img, int stride, int w, int h, int th, float* out, int cap) {
  std::vector<RawKp> k;
  fast_roi(img, stride, w, h, th, k);
  int n = std::min((int)k.size(), cap);
  for (int i = 0; i < n; i++) { out[3 * i] = k[i].x; out[3 * i + 1] = k[i].y; out[3 * i + 2] = k[i].response; }
  return (int)k.size();
}

int plvio_grid_fast(const u8* img, int stride, int w, int h, int iniTh, int minTh, float* out,
                    int cap) {
  std::vector<RawKp> k;
  grid_fast(img, stride, w, h, iniTh, minTh, k);
  int n = std::min((int)k.size(), cap);
  for (int i = 0; i < n; i++) { out[3 * i] = k[i].x; out[3 * i + 1] = k[i].y; out[3 * i + 2] = k[i].response; }
  return (int)k.size();
}

// cands = [x,y,response] triples relative to (minX,minY); out_idx = selected
// candidate indices in reference output order.  returns count.
int plvio_distribute_octree(const float* cands, int n, int minX, int maxX, int minY, int maxY,
                            int N, int* out_idx, int cap) {
  std::vector<RawKp> c(n);
  for (int i = 0; i < n; i++) c[i] = {cands[3 * i], cands[3 * i + 1], cands[3 * i + 2]};
  std::vector<int> r;
  distribute_octree(c, minX, maxX, minY, maxY, N, r);
  int m = std::min((int)r.size(), cap);
  for (int i = 0; i < m; i++) out_idx[i] = r[i];
  return (int)r.size();
}

float plvio_ic_angle(const u8* img, int stride, float x, float y) {
  OrbPlan p;
  make_plan(64, 64, 100, 1.2f, 1, p);
  return ic_angle(img, stride, x, y, p.umax);
}

void plvio_orb_descriptor(const u8* img, int stride, float x, float y, float angle, u8* desc) {
  orb_descriptor(img, stride, x, y, angle, desc);
}

// Whole ORBextractor::operator() (src/ORBextractor.cc:1068-1150).
//  img: w x h u8 with row stride `stride`.
//  kps/desc: caller-allocated with capacity `cap` rows; returns number of keypoints
//  (or -1 on empty image, like the reference); *mono_index receives the return value
//  of the reference operator().  Optional debug outputs (may be NULL):
//  pyr_out: concatenated dense level images (level 0 first); blur_out: same for the
//  blurred levels; lvl_counts[nlevels]: keypoints per level.
int plvio_orb_extract(const u8* img, int w, int h, int stride, int nfeatures, float sf,
                      int nlevels, int iniTh, int minTh, int lap0, int lap1,
                      plvio_keypoint* kps, u8* desc, int cap, int* mono_index, u8* pyr_out,
                      u8* blur_out, int* lvl_counts) {
  if (!img || w <= 0 || h <= 0) return -1;
  OrbPlan p;
  make_plan(w, h, nfeatures, sf, nlevels, p);
  std::vector<std::vector<u8>> pyr(nlevels);
  pyr[0].resize((size_t)w * h);
  for (int y = 0; y < h; y++) memcpy(&pyr[0][(size_t)y * w], img + (size_t)y * stride, w);
  for (int l = 1; l < nlevels; l++) {
    pyr[l].resize((size_t)p.w[l] * p.h[l]);
    resize_linear_u8(pyr[l - 1].data(), p.w[l - 1], p.w[l - 1], p.h[l - 1], pyr[l].data(), p.w[l],
                     p.w[l], p.h[l]);
  }
  std::vector<std::vector<plvio_keypoint>> all(nlevels);
  for (int l = 0; l < nlevels; l++) {
    int lw = p.w[l], lh = p.h[l];
    std::vector<RawKp> cands;
    grid_fast(pyr[l].data(), lw, lw, lh, iniTh, minTh, cands);
    std::vector<int> sel;
    distribute_octree(cands, 16, lw - 16, 16, lh - 16, p.quota[l], sel);
    const int scaledPatch = (int)(31 * p.scale[l]);
    for (int id : sel) {
      plvio_keypoint k;
      k.x = cands[id].x + 16;
      k.y = cands[id].y + 16;
      k.size = (float)scaledPatch;
      k.response = cands[id].response;
      k.octave = l;
      k.class_id = -1;
      k.angle = ic_angle(pyr[l].data(), lw, k.x, k.y, p.umax);
      all[l].push_back(k);
    }
    if (lvl_counts) lvl_counts[l] = (int)all[l].size();
  }
  int n = 0;
  for (int l = 0; l < nlevels; l++) n += (int)all[l].size();
  if (pyr_out) {
    size_t o = 0;
    for (int l = 0; l < nlevels; l++) { memcpy(pyr_out + o, pyr[l].data(), pyr[l].size()); o += pyr[l].size(); }
  }
  if (n > cap) return -2;
  static const int k7[7] = {18, 34, 48, 56, 48, 34, 18};
  int mono = 0, stereo = n - 1;
  size_t bo = 0;
  for (int l = 0; l < nlevels; l++) {
    int lw = p.w[l], lh = p.h[l];
    std::vector<u8> blur((size_t)lw * lh);
    gaussian_blur_u8(pyr[l].data(), lw, lw, lh, blur.data(), lw, k7, 7);
    if (blur_out) { memcpy(blur_out + bo, blur.data(), blur.size()); bo += blur.size(); }
    for (auto& k : all[l]) {
      u8 d[32];
      orb_descriptor(blur.data(), lw, k.x, k.y, k.angle, d);
      if (l != 0) { k.x *= p.scale[l]; k.y *= p.scale[l]; }
      int slot;
      if (k.x >= (float)lap0 && k.x <= (float)lap1) slot = stereo--;
      else slot = mono++;
      kps[slot] = k;
      memcpy(desc + (size_t)slot * 32, d, 32);
    }
  }
  if (mono_index) *mono_index = mono;
  return n;
}

}  // extern "C"
