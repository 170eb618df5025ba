// Line feature kernels for sm_100a: LSD (refine 0) + KeyLine assembly + LBD.
//
//   k_lsd_pre         u8 -> f64 Gaussian blur + lsd_scale bilinear resize +
//                     ll_angle gradient/angle/"available" bitmap   src/LSD/lsd.cpp:415-457,536-585
//   k_lsd_grow        seed scan (raster order) + region_grow       src/LSD/lsd.cpp:476-487,635-686,1136-1152
//   k_lsd_rect        region2rect + get_theta                      src/LSD/lsd.cpp:688-782, 506-521
//   k_line_assemble   LSDDetectorC::detectImpl KeyLine fields,     LSDDetector_custom.cpp:304-346
//                     top-N by response                            src/LineExtractor.cc:75-84
//   k_gauss5 / k_pyrdown / k_sobel   LBD pyramid + Sobel           binary_descriptor_custom.cpp:351-399
//   k_lbd_rows/_fold  computeLBD + binary packing + line equation  binary_descriptor_custom.cpp:1027-1373,402-413,663-667;
//                                                                  src/LineExtractor.cc:106-116
//
// Region growing is sequential by construction (raster-order seeds coupled through the
// `used` marks, a float running mean angle updated after every accepted pixel), so one
// warp owns one (frame, octave): the "available" bitmap lives in shared memory, the 3x3
// neighbourhoods of up to three queued pixels are fetched by 27 lanes at once, and only
// the accept/update chain is serial.  Parallelism comes from frames x octaves.
// All f64/f32 arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction)
// so results follow oracle/oracle_line.cpp operation by operation.
#include "plvi_internal.cuh"
#include "line_internal.cuh"

namespace plvi {

// BORDER_REFLECT_101 for -len < p < 2 * len - 1 (one reflection: halos are <= 8 pixels, images >= 2 * halo + 2)
__device__ __forceinline__ int reflect101_l(int p, int len) {
  p = p < 0 ? -p : p;
  return p >= len ? 2 * (len - 1) - p : p;
}

__device__ __forceinline__ float fast_atan2_dev(float y, float x) {
  const float sc = 57.29577951308232f;
  const float p1 = __fmul_rn(0.9997878412794807f, sc), p3 = __fmul_rn(-0.3258083974640975f, sc);
  const float p5 = __fmul_rn(0.1555786518463281f, sc), p7 = __fmul_rn(-0.04432655554792128f, sc);
  const float eps = 2.220446049250313e-16f;
  const float ax = fabsf(x), ay = fabsf(y);
  // one division and one polynomial for both octants (same operations as the two branches of cv::fastAtan2)
  const bool steep = !(ax >= ay);
  const float c = __fdiv_rn(steep ? ax : ay, __fadd_rn(steep ? ay : ax, eps));
  const float c2 = __fmul_rn(c, c);
  float a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  if (steep) a = __fsub_rn(90.f, a);
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

#define D2R 0.017453292519943295  // CV_PI / 180
#define PI_D 3.14159265358979323846

// ---------------------------------------------------------------------------------------
// k_lsd_pre: everything flsd does before the seed loop (src/LSD/lsd.cpp:415-457, 536-585) for one tile of the scaled
// image, with no intermediate in global memory: u8 -> f64 cv::GaussianBlur (row pass in sequential tap order, column
// pass with the symmetric pairs added first, BORDER_REFLECT_101), cv::resize INTER_LINEAR with float32 weights applied
// in double (horizontal, then vertical), then ll_angle: 2x2 gradient, level-line angle, cos / sin records of a pixel
// and the "available" bitmap.  CTA = 32 x preTH scaled pixels (32 columns = one bitmap word); the tile's source window
// (+ hk halo) lives in shared memory: u8 window -> row-filtered f64 -> column-filtered f64 -> scaled tile (+1 halo).
// HK = half width of the Gaussian (3 for lsd_scale 0.8; 0: lsd_scale == 1, the image itself is the working image).
// ---------------------------------------------------------------------------------------
// cos/sin of a level-line angle a in [0, 2 pi): a = k * (2 pi / 1024) + r, table values for the
// grid point (correctly rounded doubles from the host libm) and short Taylor series for |r| <=
// pi/1024.  Accurate to a few ulp of double, i.e. far below the float rounding that follows.
#define TRIG_N 1024
__device__ __forceinline__ void sincos_tab(double a, const double2* __restrict__ tab, double& s, double& c) {
  const double step = 2.0 * PI_D / TRIG_N;
  int k = __double2int_rn(a * (TRIG_N / (2.0 * PI_D)));
  const double r = a - (double)k * step;
  const double2 t = __ldg(tab + (k & (TRIG_N - 1)));   // {cos, sin} of k * step
  const double r2 = r * r;
  const double cr = 1.0 + r2 * (-0.5 + r2 * (1.0 / 24 + r2 * (-1.0 / 720 + r2 * (1.0 / 40320))));
  const double sr = r * (1.0 + r2 * (-1.0 / 6 + r2 * (1.0 / 120 + r2 * (-1.0 / 5040))));
  c = t.x * cr - t.y * sr;
  s = t.y * cr + t.x * sr;
}

#define PRE_TW 32

template <int HK>
__global__ void __launch_bounds__(256) k_lsd_pre(const __grid_constant__ LineGeom g, int oct, const u8* __restrict__ src,
                                                 int spitch, size_t sfs, const LineTab* __restrict__ tabs, LineBufs b) {
  extern __shared__ __align__(16) double smem_d[];
  const LineOct& O = g.o[oct];
  const int TH = g.preTH, SW = O.preSW, RH = O.preRH;
  const int f = blockIdx.z, tid = threadIdx.x;
  const int x0 = blockIdx.x * PRE_TW, y0 = blockIdx.y * TH;
  double* rowf = smem_d;                                      // [RH + 2 HK + 1][SW]  row-filtered source window (+ 1 spare row)
  double* sblur = rowf + (RH + 2 * HK + 1) * SW;              // [RH][SW]             + column pass
  double* ssc = smem_d;                                       // [TH + 1][PRE_TW + 2] scaled tile; reuses rowf after the column pass
  const int UW = (SW + 2 * HK + 11) & ~3;                     // u8 row stride (multiple of 4: 32-bit stores)
  u8* su8 = reinterpret_cast<u8*>(sblur + RH * SW);           // [RH + 2 HK][UW]
  const u8* img = src + (size_t)f * sfs;
  const int SSW = PRE_TW + 2;
  if (x0 >= O.sw) {   // sw is a multiple of 32: the last word of every bitmap row is padding only
    for (int ty = tid; ty < TH; ty += 256)
      if (y0 + ty < O.sh) b.bitmap[(size_t)f * g.bmTotal + O.bmOff + (size_t)(y0 + ty) * O.wpr + (x0 >> 5)] = 0u;
    return;
  }

  if (HK == 0) {
    // lsd_scale == 1: scaled_image = image (src/LSD/lsd.cpp:461)
    for (int e = tid; e < (TH + 1) * (PRE_TW + 1); e += 256) {
      const int ty = e / (PRE_TW + 1), tx = e - ty * (PRE_TW + 1);
      const int sx = x0 + tx, sy = y0 + ty;
      double v = 0.0;
      if (sx < O.sw && sy < O.sh) {
        v = (double)__ldg(img + (size_t)sy * spitch + sx);
        if (b.scaledDbg && tx < PRE_TW && ty < TH) b.scaledDbg[(size_t)f * g.pxTotal + O.pxOff + (size_t)sy * O.sw + sx] = v;
      }
      ssc[ty * SSW + tx] = v;
    }
  } else {
    const LineTab* xt = tabs + O.xtabOff;
    const LineTab* yt = tabs + O.ytabOff;
    // source window of the tile: columns sx0 .. sx0 + nsx - 1, rows sy0 .. sy0 + nsy - 1 (<= SW, RH: checked on the host)
    const int sx0 = __ldg(&xt[x0].ofs), sx1 = min(__ldg(&xt[min(x0 + PRE_TW, O.sw - 1)].ofs) + 1, O.w - 1);
    const int sy0 = __ldg(&yt[y0].ofs), sy1 = min(__ldg(&yt[min(y0 + TH, O.sh - 1)].ofs) + 1, O.h - 1);
    const int nsx = sx1 - sx0 + 1, nsy = sy1 - sy0 + 1;
    const int G = (nsx + 3) >> 2;                 // groups of 4 row-filter outputs per window row
    const int ncol = 4 * G + 2 * HK;              // u8 columns the row pass reads
    const int nrow = nsy + 2 * HK;
    // ---- u8 window (+ HK halo, reflected at the image border)
    const int wx0 = sx0 - HK, wy0 = sy0 - HK;
    int xoff = wx0 & 3;
    const int ax0 = wx0 - xoff, nw = (ncol + xoff + 3) >> 2;
    const bool wordsOk = ((spitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(img) & 3) == 0) && wx0 >= 0 && ax0 + 4 * nw <= O.w;
    if (wordsOk) {   // 32-bit loads from the 4-aligned column below the window
      const float inw = 1.f / (float)nw;
      for (int e = tid; e < nrow * nw; e += 256) {
        const int r = (int)(((float)e + 0.5f) * inw), wv = e - r * nw;   // e / nw (e < 2^16, nw < 2^8: the float quotient cannot cross an integer)
        const int yy = reflect101_l(wy0 + r, O.h);
        *reinterpret_cast<uint32_t*>(su8 + r * UW + 4 * wv) = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)yy * spitch + ax0) + wv);
      }
    } else {
      xoff = 0;
      for (int e = tid; e < nrow * ncol; e += 256) {
        const int r = e / ncol, c = e - r * ncol;
        const int yy = reflect101_l(wy0 + r, O.h);
        const int xx = reflect101_l(min(wx0 + c, O.w + HK), O.w);
        su8[r * UW + c] = __ldg(img + (size_t)yy * spitch + xx);
      }
    }
    __syncthreads();
    // ---- row pass: thread = 4 adjacent outputs of one window row (4 + 2 HK inputs converted once)
    const float iG = 1.f / (float)G;
    for (int e = tid; e < nrow * G; e += 256) {
      const int r = (int)(((float)e + 0.5f) * iG), c4 = (e - r * G) * 4;
      const u8* q = su8 + r * UW + xoff + c4;
      double v[4 + 2 * HK];
#pragma unroll
      for (int i = 0; i < 4 + 2 * HK; i++) v[i] = (double)q[i];
      double o[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        double sacc = __dmul_rn(g.kern[0], v[k]);
#pragma unroll
        for (int i = 1; i <= 2 * HK; i++) sacc = __dadd_rn(sacc, __dmul_rn(g.kern[i], v[k + i]));
        o[k] = sacc;
      }
      double2* out = reinterpret_cast<double2*>(rowf + r * SW + c4);
      out[0] = make_double2(o[0], o[1]);
      out[1] = make_double2(o[2], o[3]);
    }
    __syncthreads();
    // ---- column pass: thread = 2 vertically adjacent outputs of one column
    {
      const int npair = (nsy + 1) >> 1;
      const float insx = 1.f / (float)nsx;
      for (int e = tid; e < npair * nsx; e += 256) {
        const int rp = (int)(((float)e + 0.5f) * insx), c = e - rp * nsx;
        const int r = 2 * rp;
        const double* q = rowf + r * SW + c;    // window row r + i  <->  image row sy0 + r - HK + i
        double v[2 * HK + 2];
#pragma unroll
        for (int i = 0; i < 2 * HK + 2; i++) v[i] = q[i * SW];   // the last row may be the spare one (its output is dropped)
#pragma unroll
        for (int k = 0; k < 2; k++) {
          if (r + k >= nsy) break;
          double sacc = __dmul_rn(g.kern[HK], v[HK + k]);
#pragma unroll
          for (int i = 1; i <= HK; i++) sacc = __dadd_rn(sacc, __dmul_rn(g.kern[HK + i], __dadd_rn(v[HK + k + i], v[HK + k - i])));
          sblur[(r + k) * SW + c] = sacc;
        }
      }
    }
    __syncthreads();
    // ---- bilinear resize: the (PRE_TW + 1) x (TH + 1) scaled pixels of the tile and its halo
    for (int e = tid; e < (PRE_TW + 1) * (TH + 1); e += 256) {
      const int ty = e / (PRE_TW + 1), tx = e - ty * (PRE_TW + 1);
      const int sx = x0 + tx, sy = y0 + ty;
      double v = 0.0;
      if (sx < O.sw && sy < O.sh) {
        const LineTab X = xt[sx], Y = yt[sy];
        const int xa = X.ofs - sx0, xb = min(X.ofs + 1, O.w - 1) - sx0;
        const int ya = Y.ofs - sy0, yb = min(Y.ofs + 1, O.h - 1) - sy0;
        if (g.lsdScale == 0.5) {
          // cv::resize switches INTER_LINEAR to the 2x2 INTER_AREA sum when both inverse scales are exactly 2
          v = __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(sblur[ya * SW + xa], sblur[ya * SW + xb]), sblur[yb * SW + xa]),
                                  sblur[yb * SW + xb]), 0.25);
        } else {
          const double h0 = __dadd_rn(__dmul_rn(sblur[ya * SW + xa], (double)X.a0), __dmul_rn(sblur[ya * SW + xb], (double)X.a1));
          const double h1 = __dadd_rn(__dmul_rn(sblur[yb * SW + xa], (double)X.a0), __dmul_rn(sblur[yb * SW + xb], (double)X.a1));
          v = __dadd_rn(__dmul_rn(h0, (double)Y.a0), __dmul_rn(h1, (double)Y.a1));
        }
        if (b.scaledDbg && tx < PRE_TW && ty < TH)
          b.scaledDbg[(size_t)f * g.pxTotal + O.pxOff + (size_t)sy * O.sw + sx] = v;
      }
      ssc[ty * SSW + tx] = v;      // rowf is dead: every thread passed the barrier after the column pass
    }
  }
  __syncthreads();
  // ---- ll_angle: warp = one row of the tile at a time, lane = column (the ballot is the row's bitmap word)
  const int tx = tid & 31;
  const int sx = x0 + tx;
  for (int ty = tid >> 5; ty < TH; ty += 8) {
    const int sy = y0 + ty;
    if (sy >= O.sh) break;          // warp-uniform
    bool avail = false;
    if (sx < O.sw) {
      float angDeg = -1024.f;
      float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
      double norm = 0.0;
      if (sx < O.sw - 1 && sy < O.sh - 1) {
        const double DA = __dsub_rn(ssc[(ty + 1) * SSW + tx + 1], ssc[ty * SSW + tx]);
        const double BC = __dsub_rn(ssc[ty * SSW + tx + 1], ssc[(ty + 1) * SSW + tx]);
        const double gx = __dadd_rn(DA, BC), gy = __dsub_rn(DA, BC);
        // (gx^2 + gy^2) / 4: the division by a power of two is exact, so the multiplication is identical
        norm = __dsqrt_rn(__dmul_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)), 0.25));
        if (!(norm <= g.rho)) {
          angDeg = fast_atan2_dev((float)gx, (float)(-gy));
          const double a = __dmul_rn((double)angDeg, D2R);
          double sd, cd;
          sincos_tab(a, b.trig, sd, cd);
          // what a region accumulates is cos(float(angle)), sin(float(angle)) evaluated by the host libm's cosf / sinf
          float cf, sf;
          glibc_sincosf((float)a, sf, cf);
          cs = make_float4(cf, sf, (float)cd, (float)sd);
          avail = true;
        }
      }
      const size_t p = (size_t)f * g.pxTotal + O.pxOff + (size_t)sy * O.sw + sx;
      b.ang[p] = angDeg;
      b.cs[p] = make_float2(cs.x, cs.y);
      b.seed[p] = make_float2(cs.z, cs.w);
      b.mod[p] = norm;
    }
    const unsigned m = __ballot_sync(0xffffffffu, avail);
    if (tx == 0) b.bitmap[(size_t)f * g.bmTotal + O.bmOff + (size_t)sy * O.wpr + (x0 >> 5)] = m;
  }
}

// ---------------------------------------------------------------------------------------
// k_lsd_grow: one warp per (octave, frame).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_aligned_dev(double a, double theta, double prec) {
  double n = __dsub_rn(theta, a);
  if (n < 0) n = -n;
  if (n > (3 * PI_D) / 2) {
    n = __dsub_rn(n, 2 * PI_D);
    if (n < 0) n = -n;
  }
  return n <= prec;
}

#define GROW_WPB SPEC_WPB   // independent warps per block in k_lsd_spec (line_internal.cuh)
#ifndef SPEC_MINB
// resident blocks per SM the register allocation of k_lsd_spec allows for: 16 x 64 threads = 64 registers (20 bytes of
// spills).  Alone the kernel is slower than with its natural 78 registers (56 vs 50 ms per 4096 frames), but a third of
// the register file stays free for the ORB kernels that run beside it: the whole step gains 3 % (12: 23.5 k, 14: 23.1 k,
// 16: 24.2 k, 18: 23.1 k, 20: 21.2 k frames/s)
#define SPEC_MINB 16
#endif
#define SPEC_MINB_SMALL 12   // 85 registers allowed: the natural allocation (78)
#ifndef COMMIT_WPB
#define COMMIT_WPB 4  // ... and in k_lsd_commit
#endif
#ifndef COMMIT_WARPS
#define COMMIT_WARPS 28
#endif
#define COMMIT_BPS (COMMIT_WARPS / COMMIT_WPB)   // resident blocks per SM: 28 warps of 72 registers
#define GROW_RQ 512   // shared ring holding the most recent region pixels (BFS frontier)
#define GROW_K 32     // bitmap rows kept in shared memory (sliding window below the seed row)

// "Available" (gradient defined & not used) bitmap of one (frame, octave).  Rows above the
// current seed row hold no available pixel any more (every earlier pixel in raster order
// has been a seed or was absorbed), so only rows [top, top + GROW_K) are cached in shared
// memory; the rare region that reaches further down works on the global copy directly.
template <int K> struct GrowBitmapT {
  unsigned* sm;    // [K][wpr]
  unsigned* gm;    // [H][wpr]
  int wpr, top;
  __device__ __forceinline__ unsigned word(int y, int wi) const {
    return (y - top < K) ? sm[(y & (K - 1)) * wpr + wi] : __ldcg(gm + y * wpr + wi);
  }
  __device__ __forceinline__ bool test(int x, int y) const { return (word(y, x >> 5) >> (x & 31)) & 1u; }
  __device__ __forceinline__ void clear(int x, int y) {   // one lane only
    const unsigned m = ~(1u << (x & 31));
    if (y - top < K) sm[(y & (K - 1)) * wpr + (x >> 5)] &= m;
    else { unsigned* p = gm + y * wpr + (x >> 5); __stcg(p, __ldcg(p) & m); }
  }
};
typedef GrowBitmapT<GROW_K> GrowBitmap;

// One neighbour per lane: lane = 8 * e + k, e = queue entry of the batch (0..3), k = 3x3
// neighbour index without the centre, in the reference's (yy, xx) scan order.
struct GrowBatch {
  unsigned mask;     // lanes holding an available neighbour (warp-uniform)
  int cpk;           // neighbour coordinates x | y << 16
  int bw;            // bitmap word of the neighbour: >= 0 index into the shared window, < 0: ~index into the global copy
  unsigned bbit;     // its bit
  float2 rec;        // cos, sin of the neighbour's float angle
};

template <int K> __device__ __forceinline__ bool grow_bit(const GrowBitmapT<K>& bm, int bw, unsigned bbit) {
  return ((bw >= 0 ? bm.sm[bw] : __ldcg(bm.gm + ~bw)) & bbit) != 0u;
}
template <int K> __device__ __forceinline__ void grow_clear(GrowBitmapT<K>& bm, int bw, unsigned bbit) {   // one lane only
  if (bw >= 0) bm.sm[bw] &= ~bbit;
  else { unsigned* p = bm.gm + ~bw; __stcg(p, __ldcg(p) & ~bbit); }
}

// Neighbourhood fetch for queue entries [i, i+nb): lane (e, k) tests the availability bit of
// its neighbour and, if set, issues the 16-byte record load (consumed later => the load
// latency overlaps the accept chain of the previous batch).
template <int K> __device__ __forceinline__ GrowBatch grow_fetch(const GrowBitmapT<K>& bm, const unsigned* ring, const unsigned* reg,
                                                int regSize, int i, int nb, int e, int ndx, int ndy, int W, int H,
                                                const float2* __restrict__ rec) {
  GrowBatch g;
  g.cpk = 0; g.bw = 0; g.bbit = 0u;
  g.rec = make_float2(0.f, 0.f);
  bool cand = false;
  if (e < nb) {
    const int idx = i + e;
    const unsigned p = (regSize - idx <= GROW_RQ) ? ring[idx & (GROW_RQ - 1)] : __ldcg(reg + idx);
    const int cx = (int)(p & 0xffff) + ndx, cy = (int)(p >> 16) + ndy;
    // x = -1 is rejected; x = W lands on a padding bit of the bitmap row (always 0); rows above
    // the window top hold no available pixel
    if (cy >= bm.top && cy < H && cx >= 0) {
      g.bw = (cy - bm.top < K) ? (cy & (K - 1)) * bm.wpr + (cx >> 5) : ~(cy * bm.wpr + (cx >> 5));
      g.bbit = 1u << (cx & 31);
      if (grow_bit(bm, g.bw, g.bbit)) {
        cand = true;
        g.cpk = cx | (cy << 16);
        g.rec = __ldg(rec + cy * W + cx);
      }
    }
  }
  g.mask = __ballot_sync(0xffffffffu, cand);
  return g;
}

__global__ void __launch_bounds__(32) k_lsd_grow(const __grid_constant__ LineGeom g, LineBufs b, int onlyFlagged) {
  extern __shared__ unsigned smem_u[];
  const int oct = blockIdx.x, f = blockIdx.y, lane = threadIdx.x;
  if (oct >= g.noct) return;
  // band-run path: only the (frame, octave) problems that asked for the serial fallback
  if (onlyFlagged && !b.brFlags[((size_t)f * 2 + oct) * BR_FLAGS]) return;
  const LineOct& O = g.o[oct];
  const int W = O.sw, H = O.sh, wpr = O.wpr;
  unsigned* ring = smem_u;
  GrowBitmap bm;
  bm.sm = smem_u + GROW_RQ;
  bm.gm = b.bitmap + (size_t)f * g.bmTotal + O.bmOff;
  bm.wpr = wpr;
  bm.top = 0;
  for (int i = lane; i < min(GROW_K, H) * wpr; i += 32) bm.sm[i] = __ldcg(bm.gm + i);
  __syncwarp();
  const size_t pbase = (size_t)f * g.pxTotal + O.pxOff;
  const float2* __restrict__ rec = b.cs + pbase;
  const float* __restrict__ ang = b.ang + pbase;
  const float2* __restrict__ seedcs = b.seed + pbase;
  unsigned* regAll = b.reg + (size_t)f * g.regTotal + O.regOff;
  LineRegion* rtab = b.regTab + (size_t)f * g.segTotal + O.segOff;
  const double prec = g.prec;
  const float kHi = g.alignHi2, kLo = g.alignLo2;
  int regBase = 0, nreg = 0;
  bool overflow = false;
  const int e = lane >> 3, k8 = lane & 7;
  const int nidx = k8 < 4 ? k8 : k8 + 1;
  const int ndx = nidx % 3 - 1, ndy = nidx / 3 - 1;
  const unsigned laneBit = 1u << lane;

  for (int row = 0; row < H - 1; row++) {
    // slide the shared window: rows [top, row) are exhausted, rows up to row + GROW_K enter
    if (row > bm.top) {
      const int r0 = bm.top + GROW_K, r1 = min(row + GROW_K, H);
      for (int r = r0; r < r1; r++)
        for (int wv = lane; wv < wpr; wv += 32) bm.sm[(r & (GROW_K - 1)) * wpr + wv] = __ldcg(bm.gm + r * wpr + wv);
      bm.top = row;
      __syncwarp();
    }
    for (int c0 = 0; c0 < wpr; c0 += 32) {
      while (true) {
        // next seed: first available pixel in raster order (src/LSD/lsd.cpp:476-479)
        const int wi = c0 + lane;
        const int rowBase = (row & (GROW_K - 1)) * wpr;
        const unsigned word = wi < wpr ? bm.sm[rowBase + wi] : 0u;
        const unsigned nz = __ballot_sync(0xffffffffu, word != 0u);
        if (!nz) break;
        const int wl = __ffs(nz) - 1;
        const unsigned sw_ = __shfl_sync(0xffffffffu, word, wl);
        const int bit = __ffs(sw_) - 1;
        const int sx = (c0 + wl) * 32 + bit, sy = row;
        const int sp = sy * W + sx;
        const float sang = __ldg(ang + sp);
        const float2 scs = __ldg(seedcs + sp);
        unsigned* reg = regAll + regBase;          // pixel list of the region being grown
        if (lane == 0) {
          bm.sm[rowBase + c0 + wl] = sw_ & ~(1u << bit);
          ring[0] = (unsigned)sx | ((unsigned)sy << 16);
        }
        __syncwarp();
        // region_grow (src/LSD/lsd.cpp:635-686).  The region angle is a pure function of the
        // float sums (fastAtan2(sumdy, sumdx)), so it is only evaluated when the candidate that
        // is next in scan order falls inside the band where the dot-product test cannot decide.
        const double seedAngle = __dmul_rn((double)sang, D2R);
        float sumdx = scs.x, sumdy = scs.y;
        bool fresh = true;   // no pixel accepted yet: reg_angle is the seed's own angle
        int regSize = 1;
        int flushed = 0;     // ring entries [flushed, regSize) are not in global memory yet
        int i = 0, nb = 1;
        GrowBatch cur = grow_fetch(bm, ring, reg, regSize, 0, 1, e, ndx, ndy, W, H, rec);
        bool curDirty = false;   // a pixel was accepted after cur's candidates were fetched
        while (nb > 0) {
          const int ni = i + nb, nnb = min(4, regSize - ni);
          GrowBatch nxt = grow_fetch(bm, ring, reg, regSize, ni, nnb, e, ndx, ndy, W, H, rec);
          // candidates fetched ahead may have been absorbed by the previous batch in the meantime (only then is their
          // availability tested again)
          unsigned pm = cur.mask;
          if (pm && curDirty) pm = __ballot_sync(0xffffffffu, (pm & laneBit) && grow_bit(bm, cur.bw, cur.bbit));
          const int sizeBefore = regSize;
          float n2 = __fmaf_rn(sumdx, sumdx, sumdy * sumdy);
          while (pm) {
            // every pending candidate against the current region direction at once
            const float dot = __fmaf_rn(sumdx, cur.rec.x, sumdy * cur.rec.y);
            const float d2 = dot * dot;
            const bool poss = (pm & laneBit) && dot > 0.f && d2 > kLo * n2;
            const unsigned possm = __ballot_sync(0xffffffffu, poss);
            if (!possm) break;                       // nobody can be aligned: batch done
            const int l = __ffs(possm) - 1;          // first in (entry, yy, xx) order
            const unsigned surem = __ballot_sync(0xffffffffu, poss && d2 >= kHi * n2);
            pm &= ~((2u << l) - 1u);                 // l and everything before it is decided now
            if (!((surem >> l) & 1u)) {              // undecided band: the exact test of the reference
              const double regAngle = fresh ? seedAngle : __dmul_rn((double)fast_atan2_dev(sumdy, sumdx), D2R);
              const int lpk = __shfl_sync(0xffffffffu, cur.cpk, l);
              const float la = __ldg(ang + (lpk >> 16) * W + (lpk & 0xffff));
              if (!is_aligned_dev(__dmul_rn((double)la, D2R), regAngle, prec)) continue;
            }
            // accept lane l's pixel
            const float qc = __shfl_sync(0xffffffffu, cur.rec.x, l), qs = __shfl_sync(0xffffffffu, cur.rec.y, l);
            const int qpk = __shfl_sync(0xffffffffu, cur.cpk, l);
            pm &= ~__ballot_sync(0xffffffffu, cur.cpk == qpk);   // the same pixel seen from another entry
            if (lane == l) {
              grow_clear(bm, cur.bw, cur.bbit);
              ring[regSize & (GROW_RQ - 1)] = (unsigned)qpk;
            }
            regSize++;
            fresh = false;
            sumdx = __fadd_rn(sumdx, qc);
            sumdy = __fadd_rn(sumdy, qs);
            n2 = __fmaf_rn(sumdx, sumdx, sumdy * sumdy);
          }
          __syncwarp();
          // spill the ring to the region's global pixel list in coalesced chunks (needed by
          // k_lsd_rect, and by grow_fetch when the frontier outgrows the ring)
          while (regSize - flushed >= 32) {
            __stcg(reg + flushed + lane, ring[(flushed + lane) & (GROW_RQ - 1)]);
            flushed += 32;
          }
          i = ni;
          if (nnb > 0) {
            cur = nxt;
            nb = nnb;
            curDirty = regSize != sizeBefore;   // fetched before this batch's accepts
          } else {
            nb = min(4, regSize - i);
            if (nb > 0) cur = grow_fetch(bm, ring, reg, regSize, i, nb, e, ndx, ndy, W, H, rec);
            curDirty = false;
          }
        }
        if (regSize >= O.minRegSize) {
          if (nreg < O.segCap) {
            if (flushed + lane < regSize) __stcg(reg + flushed + lane, ring[(flushed + lane) & (GROW_RQ - 1)]);
            if (lane == 0) {
              LineRegion r;
              r.start = regBase;
              r.size = regSize;
              r.angle = fresh ? seedAngle : __dmul_rn((double)fast_atan2_dev(sumdy, sumdx), D2R);
              rtab[nreg] = r;
            }
            nreg++;
            regBase += regSize;
          } else {
            overflow = true;
          }
        }
        __syncwarp();
      }
    }
  }
  if (lane == 0) b.regCount[f * 2 + oct] = overflow ? -1 : nreg;
}

// ---------------------------------------------------------------------------------------
// Band-speculative region growing.  The reference grows regions one after another in raster
// order of their seeds (src/LSD/lsd.cpp:476-487) - a serial chain over the whole image.  Here
//   k_lsd_spec_init  gives every band (O.bandRows working rows) a private copy of the initial
//                    availability bitmap (rows above the band are exhausted by definition),
//   k_lsd_spec       grows, one thread per band, the regions seeded in that band as if the
//                    band were the first one: 32 independent serial chains per warp,
//   k_lsd_commit     walks the true seeds in raster order (one warp per frame and octave).  A
//                    speculative region is adopted when it starts at the true next seed, all
//                    of its pixels are still available and no pixel it saw as "used" is in
//                    fact available ("phantom": consumed only by a discarded speculation of
//                    the same band); its growth then saw the same availability and the same
//                    angles as the serial algorithm, so it is the region the reference grows.
//                    Everything else is grown serially as in k_lsd_grow.
// ---------------------------------------------------------------------------------------
// Band boundaries of one (frame, octave): the kernel time of k_lsd_spec is the LONGEST chain of the batch, and a chain
// is as long as the number of available pixels it walks (each is a seed or is expanded once).  Bands of equal rows
// put 2-3 times the mean load on the textured part of a frame; here the rows are cut so that every band holds the same
// number of available pixels (any partition of the rows is exact: k_lsd_commit checks every region).
__global__ void __launch_bounds__(256) k_lsd_band_split(const __grid_constant__ LineGeom g, LineBufs b) {
  const int oct = blockIdx.x, f = blockIdx.y;
  if (oct >= g.noct) return;
  const LineOct& O = g.o[oct];
  const int H = O.sh, wpr = O.wpr, nb = O.nbands;
  __shared__ int cum[1025];
  __shared__ int wsum[8];
  const unsigned* __restrict__ bm = b.bitmap + (size_t)f * g.bmTotal + O.bmOff;
  int* tab = b.bandRow + (size_t)f * (g.tasksPerFrame + 2) + O.taskOff + oct;
  if (H > 1024 || !b.eqLoad) {   // switched off (or taller than the scan below holds): bands of equal rows
    for (int k = threadIdx.x; k <= nb; k += 256) tab[k] = k == nb ? H : min(k * O.bandRows, H);
    return;
  }
  // rows per thread (consecutive), seed rows are [0, H - 1)
  const int rpt = (H + 255) / 256;
  int loc[4];
  int tot = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int y = threadIdx.x * rpt + k;
    int c = 0;
    if (k < rpt && y < H - 1)
      for (int w = 0; w < wpr; w++) c += __popc(bm[y * wpr + w]);
    loc[k] = c;
    tot += c;
  }
  // block-wide exclusive scan of the per-thread totals
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = tot;
  for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < wid; w++) base += wsum[w];
  int run = base + inc - tot;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int y = threadIdx.x * rpt + k;
    if (k < rpt && y < H) cum[y] = run;    // available pixels in rows [0, y)
    run += loc[k];
  }
  __syncthreads();
  int total = 0;
  for (int w = 0; w < 8; w++) total += wsum[w];
  // band k starts at the first row y with cum[y] >= k * total / nb (monotone in k; empty bands are allowed)
  for (int k = threadIdx.x; k <= nb; k += 256) {
    int y;
    if (k == 0) y = 0;
    else if (k == nb) y = H;
    else {
      const long long target = (long long)k * total / nb;
      int lo = 0, hi = H - 1;              // smallest y in [0, H - 1] with cum[y] >= target (cum[H - 1] = total of rows < H - 1)
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (cum[mid] >= target) hi = mid; else lo = mid + 1; }
      y = lo;
    }
    tab[k] = y;
  }
}

// one block per (band, frame): 20 blocks of ~15 words per thread instead of 160 of ~2 (the kernel was bound by the rate
// at which blocks start, not by the 0.5 MB per frame it moves)
__global__ void __launch_bounds__(256) k_lsd_spec_init(const __grid_constant__ LineGeom g, LineBufs b) {
  const int t = blockIdx.x, f = blockIdx.y;
  const int oct = (g.noct > 1 && t >= g.o[1].taskOff) ? 1 : 0;
  const LineOct& O = g.o[oct];
  const int j = t - O.taskOff;
  if (j >= O.nbands) return;
  const int r0 = b.bandRow[(size_t)f * (g.tasksPerFrame + 2) + t + oct];
  const int nw = (O.sh - r0) * O.wpr;
  const unsigned* __restrict__ src = b.bitmap + (size_t)f * g.bmTotal + O.bmOff + r0 * O.wpr;
  unsigned* __restrict__ dst = b.specBm + (size_t)f * g.specBmTotal + O.specBmOff + (size_t)j * O.wpr * O.sh + r0 * O.wpr;
  for (int i0 = threadIdx.x; i0 < nw; i0 += 256 * 4) {   // four independent words per thread in flight
    unsigned v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = i0 + 256 * u < nw ? __ldg(src + i0 + 256 * u) : 0u;
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i0 + 256 * u < nw) dst[i0 + 256 * u] = v[u];
  }
  if (j == 0) {   // the commit pass starts without phantom pixels
    unsigned* ph = b.phantom + (size_t)f * g.bmTotal + O.bmOff;
    for (int i = threadIdx.x; i < nw; i += 256) ph[i] = 0u;
  }
}

#define SPEC_SCAN_WORDS 8
#define SPEC_RING 32

// One thread per (band, frame); the 32 lanes of a warp hold the same band of 32 consecutive
// frames (similar content => similar amount of work).  Single flat loop: every iteration
// expands one queue entry of the lane's current region, so lanes with regions of different
// sizes stay converged.  Same tests, in the same order, as k_lsd_grow.
// EQLOAD: band rows from the table of k_lsd_band_split; else bands of equal rows (the first row is then a warp-uniform
// value, which keeps a few comparisons per iteration on the uniform datapath: 3 ms per 4096 frames)
template <bool EQLOAD, int MINB>
__global__ void __launch_bounds__(32 * GROW_WPB, MINB) k_lsd_spec(const __grid_constant__ LineGeom g, LineBufs b, int n) {
  const int t = blockIdx.x;
  const int f = blockIdx.y * (32 * GROW_WPB) + threadIdx.x;
  const int oct = (g.noct > 1 && t >= g.o[1].taskOff) ? 1 : 0;
  const LineOct& O = g.o[oct];
  const int j = t - O.taskOff;
  if (threadIdx.x == 0) atomicAdd(b.specStart, 1);   // plvi_line_stage_counter: this block is resident
  if (f >= n || j >= O.nbands) return;
  const int W = O.sw, H = O.sh, wpr = O.wpr;
  int r0 = j * O.bandRows, r1 = min(r0 + O.bandRows, H - 1);
  if (EQLOAD) {
    const int* tab = b.bandRow + (size_t)f * (g.tasksPerFrame + 2) + t + oct;
    r0 = tab[0]; r1 = min(tab[1], H - 1);
  }
  unsigned* P = b.specBm + (size_t)f * g.specBmTotal + O.specBmOff + (size_t)j * wpr * H;
  unsigned* list = b.reg + (size_t)f * g.regTotal + O.regOff + (size_t)W * H + (size_t)j * O.bandPxCap;
  uint4* recs = reinterpret_cast<uint4*>(b.specRec + (size_t)f * g.specRecTotal + O.specRecOff + (size_t)j * O.bandRecCap);
  const size_t pbase = (size_t)f * g.pxTotal + O.pxOff;
  const float2* __restrict__ rec = b.cs + pbase;
  const float* __restrict__ ang = b.ang + pbase;
  const float2* __restrict__ seedcs = b.seed + pbase;
  const int pxCap = O.bandPxCap, recCap = O.bandRecCap, minReg = O.minRegSize;
  const double prec = g.prec;
  const float kHi = g.alignHi2, kLo = g.alignLo2;

  // the private bitmap is read and written through L2 only (__ldcg / __stcg): its words are re-read right after
  // they were stored, and L1 (write-through, shared with 500+ other chains) never keeps them anyway
  __shared__ unsigned sring_all[GROW_WPB][SPEC_RING * 32];   // the last SPEC_RING pixels of each lane's region (BFS frontier)
  unsigned* sring = sring_all[threadIdx.x >> 5];
  const unsigned sringAddr = (unsigned)__cvta_generic_to_shared(sring + (threadIdx.x & 31));   // this lane's column of the ring
  const int lane = threadIdx.x & 31;
  int row = r0, wi = 0;
  unsigned word = (r0 < r1) ? __ldcg(P + r0 * wpr) : 0u;
  int nrec = 0, npx = 0, base = 0, size = 0, i = 0;
  unsigned seedpk = 0u;
  float sang = 0.f, sumdx = 0.f, sumdy = 0.f, n2 = 0.f;
  bool fresh = true;
  bool done = r0 >= r1;
  while (!done) {
    bool isNew = false;
    float sangNew = 0.f;
    float2 scsNew = make_float2(0.f, 0.f);
    if (i == size) {
      if (size > 0) {   // close the finished region
        const float ang = fresh ? sang : (size >= minReg ? fast_atan2_dev(sumdy, sumdx) : 0.f);
        recs[nrec++] = make_uint4(seedpk, (unsigned)base, (unsigned)size, __float_as_uint(ang));
        npx += size;
        size = 0; i = 0;
        word = __ldcg(P + row * wpr + wi);   // the region may have taken pixels of the seed's own word
      }
      int guard = 0;
      while (word == 0u && guard < SPEC_SCAN_WORDS) {
        if (++wi == wpr) { wi = 0; ++row; }
        if (row >= r1) break;
        word = __ldcg(P + row * wpr + wi);
        ++guard;
      }
      if (row >= r1 || nrec >= recCap) { done = true; continue; }
      if (word == 0u) continue;
      const int bit = __ffs(word) - 1;
      const int sx = wi * 32 + bit;
      seedpk = (unsigned)sx | ((unsigned)row << 16);
      word &= ~(1u << bit);
      __stcg(P + row * wpr + wi, word);
      base = npx;
      list[base] = seedpk;
      size = 1;
      const int sp = row * W + sx;
      // consumed after the neighbourhood loads below have been issued
      sangNew = __ldg(ang + sp);
      scsNew = __ldg(seedcs + sp);
      isNew = true;
    }
    if (base + size + 8 > pxCap) { done = true; continue; }   // list full: the unfinished region is dropped
    // expand queue entry i
    const unsigned p = (i == 0) ? seedpk
                                : ((size - i <= SPEC_RING) ? sring[(i & (SPEC_RING - 1)) * 32 + lane] : list[base + i]);
    const int ex = (int)(p & 0xffff), ey = (int)(p >> 16);
    const int xm = ex - 1;
    const int wa = max(xm, 0) >> 5;
    const int sh = xm - (wa << 5);   // -1 .. 31: bit of column ex - 1 in word wa
    const bool needHi = sh >= 30 && wa + 1 < wpr;
    unsigned lo[3], hi[3];
    unsigned m9 = 0u;
    // row pointers are formed once (and kept opaque, so that the compiler does not re-derive every address from the
    // kernel parameters): the loads below take constant offsets from them
    unsigned* Pw0 = P + ((ey - 1) * wpr + wa);   // word of column ex - 1 in row ey - 1
    unsigned* Pw1 = Pw0 + wpr;
    unsigned* Pw2 = Pw1 + wpr;
    const float2* rq0 = rec + ((ey - 1) * W + xm);
    const float2* rq1 = rq0 + W;
    const float2* rq2 = rq1 + W;
    asm volatile("" : "+l"(Pw0), "+l"(Pw1), "+l"(Pw2), "+l"(rq0), "+l"(rq1), "+l"(rq2));
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const int y = ey - 1 + r;
      const bool ok = y >= r0 && y < H;
      unsigned* const Pw = r == 0 ? Pw0 : (r == 1 ? Pw1 : Pw2);
      lo[r] = ok ? __ldcg(Pw) : 0u;
      hi[r] = (ok && needHi) ? __ldcg(Pw + 1) : 0u;
      const unsigned long long comb = ((unsigned long long)hi[r] << 32) | lo[r];
      const unsigned three = sh >= 0 ? ((unsigned)(comb >> sh) & 7u) : ((lo[r] << 1) & 7u);
      m9 |= three << (3 * r);
    }
    float2 rk[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
      rk[k] = make_float2(0.f, 0.f);
      const float2* const rq = k < 3 ? rq0 : (k < 6 ? rq1 : rq2);
      if (k != 4 && ((m9 >> k) & 1u)) rk[k] = __ldg(rq + k % 3);
    }
    if (isNew) {
      sang = sangNew;
      sumdx = scsNew.x; sumdy = scsNew.y;
      n2 = __fmaf_rn(sumdx, sumdx, sumdy * sumdy);
      fresh = true;
    }
    unsigned acc = 0u;
    unsigned* lp = list + (base + size);          // next free slot of the pixel list
    asm volatile("" : "+l"(lp));
    float tLo = kLo * n2, tHi = kHi * n2;
#pragma unroll
    for (int k = 0; k < 9; k++) {
      if (k == 4) continue;
      // a neighbour that is not available has rk = (0, 0): dot = 0 fails the first test
      const float dot = __fmaf_rn(sumdx, rk[k].x, sumdy * rk[k].y);
      const float d2 = dot * dot;
      bool take = dot > 0.f && d2 > tLo;
      if (take && !(d2 >= tHi)) {   // undecided band of the cheap test: the reference's own comparison (rare)
        const double regAngle = __dmul_rn((double)(fresh ? sang : fast_atan2_dev(sumdy, sumdx)), D2R);
        const float la = __ldg(ang + ((ey + k / 3 - 1) * W + (ex + k % 3 - 1)));
        take = is_aligned_dev(__dmul_rn((double)la, D2R), regAngle, prec);
      }
      // the accept is straight-line code (two predicated stores, selects): no divergent branch around the common path
      const unsigned q = p + (unsigned)((k / 3 - 1) * 65536 + (k % 3 - 1));
      const int tk = take ? 1 : 0;
      asm volatile("{ .reg .pred pt; setp.ne.s32 pt, %0, 0; @pt st.global.u32 [%1], %2; @pt st.shared.u32 [%3], %2; }"
                   :: "r"(tk), "l"(lp), "r"(q), "r"(sringAddr + ((size & (SPEC_RING - 1)) << 7)) : "memory");
      acc |= (unsigned)tk << k;
      lp += tk;
      size += tk;
      fresh = fresh && !take;
      const float nx = __fadd_rn(sumdx, rk[k].x), ny = __fadd_rn(sumdy, rk[k].y);
      sumdx = take ? nx : sumdx;
      sumdy = take ? ny : sumdy;
      n2 = __fmaf_rn(sumdx, sumdx, sumdy * sumdy);
      tLo = kLo * n2; tHi = kHi * n2;
    }
    if (acc) {
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const unsigned a3 = (acc >> (3 * r)) & 7u;
        if (a3) {
          const unsigned long long mk = sh >= 0 ? ((unsigned long long)a3 << sh) : (unsigned long long)(a3 >> 1);
          unsigned* const Pw = r == 0 ? Pw0 : (r == 1 ? Pw1 : Pw2);
          __stcg(Pw, lo[r] & ~(unsigned)mk);
          if ((unsigned)(mk >> 32)) __stcg(Pw + 1, hi[r] & ~(unsigned)(mk >> 32));
        }
      }
    }
    i++;
  }
  b.specCnt[(size_t)f * g.tasksPerFrame + t] = nrec;
}

// Pixels that a discarded speculative region had consumed in its band's private bitmap but
// that are still available in the true state.  A later speculative region of that band saw
// them as "used"; it can only be adopted if none of them touches it.
template <int K> struct PhantomMapT {
  unsigned* sm;      // [K][wpr] shared window, slides with the availability window
  unsigned* gm;      // [H][wpr] global copy (rows below the window)
  unsigned* rm = nullptr;   // [K] per window row: bit min(w, 31) set if word w of the row holds a phantom pixel (nullptr: not kept)
  bool any;          // warp-uniform: some phantom pixel exists
  // warp-uniform bounding box (grown by one pixel) of the phantom pixels marked for the CURRENT band: only a
  // discarded speculation of the same band can have hidden a pixel from a speculative region, and most regions lie
  // outside the box -- their pixels skip the neighbourhood test
  int bx0, bx1, by0, by1;
  __device__ __forceinline__ void box_reset() { bx0 = 1 << 20; bx1 = -(1 << 20); by0 = 1 << 20; by1 = -(1 << 20); }
  __device__ __forceinline__ void box_add(int x0, int x1, int y0, int y1) {   // lane-local extents of newly marked pixels
    bx0 = min(bx0, __reduce_min_sync(0xffffffffu, x0) - 1);
    bx1 = max(bx1, __reduce_max_sync(0xffffffffu, x1) + 1);
    by0 = min(by0, __reduce_min_sync(0xffffffffu, y0) - 1);
    by1 = max(by1, __reduce_max_sync(0xffffffffu, y1) + 1);
  }
  __device__ __forceinline__ bool in_box(int x, int y) const { return x >= bx0 && x <= bx1 && y >= by0 && y <= by1; }
  __device__ __forceinline__ unsigned word(const GrowBitmapT<K>& bm, int y, int wi) const {
    return (y - bm.top < K) ? sm[(y & (K - 1)) * bm.wpr + wi] : __ldcg(gm + y * bm.wpr + wi);
  }
  __device__ __forceinline__ void mark(const GrowBitmapT<K>& bm, int x, int y) {
    const unsigned bit = 1u << (x & 31);
    if (y - bm.top < K) {
      atomicOr(&sm[(y & (K - 1)) * bm.wpr + (x >> 5)], bit);
      if (rm) atomicOr(&rm[y & (K - 1)], 1u << min(x >> 5, 31));
    } else atomicOr(gm + y * bm.wpr + (x >> 5), bit);
  }
  // any available phantom pixel in the 3x3 neighbourhood of (x, y)?  (y >= bm.top; rows above the window top hold
  // no available pixel).  Same instruction path for every lane: the three rows are read unconditionally (a row
  // outside [top, H) is replaced by row y and masked out) and phantom & available is formed before the 3-bit extract.
  __device__ __forceinline__ bool near(const GrowBitmapT<K>& bm, int x, int y, int H) const {
    const int xm = x - 1;
    const int wa = max(xm, 0) >> 5;
    // quick reject from the row masks while the three rows lie inside the window (a row above the window top reads the
    // mask of a row further down: a superset, never a miss)
    if (rm && y + 1 - bm.top < K) {
      const unsigned wm = (1u << min(wa, 31)) | (1u << min((x + 1) >> 5, 31));
      if (!((rm[(y - 1) & (K - 1)] | rm[y & (K - 1)] | rm[(y + 1) & (K - 1)]) & wm)) return false;
    }
    const int sh = xm - (wa << 5);   // -1 .. 31
    const bool two = sh >= 30 && wa + 1 < bm.wpr;
    unsigned hit = 0u;
#pragma unroll
    for (int r = -1; r <= 1; r++) {
      const int yy = y + r;
      const bool ok = yy >= bm.top && yy < H;
      const int yc = ok ? yy : y;
      unsigned long long pa = (unsigned long long)(word(bm, yc, wa) & bm.word(yc, wa));
      if (two) pa |= (unsigned long long)(word(bm, yc, wa + 1) & bm.word(yc, wa + 1)) << 32;
      const unsigned m3 = sh >= 0 ? ((unsigned)(pa >> sh) & 7u) : (((unsigned)pa << 1) & 7u);
      hit |= ok ? m3 : 0u;
    }
    return hit != 0u;
  }
};
typedef PhantomMapT<GROW_K> PhantomMap;

template <int K> __device__ __forceinline__ void grow_clear_atomic(GrowBitmapT<K>& bm, int x, int y) {
  const unsigned m = ~(1u << (x & 31));
  if (y - bm.top < K) atomicAnd(&bm.sm[(y & (K - 1)) * bm.wpr + (x >> 5)], m);
  else atomicAnd(bm.gm + y * bm.wpr + (x >> 5), m);
}

__global__ void __launch_bounds__(32 * COMMIT_WPB, COMMIT_BPS) k_lsd_commit(const __grid_constant__ LineGeom g, LineBufs b, int n,
                                                              int smemWordsPerWarp) {
  // COMMIT_WPB independent warps per block (consecutive frames of one octave): single-warp blocks
  // would fill the SM's 32 block slots and keep the kernels of the other streams out
  extern __shared__ unsigned smem_all[];
  unsigned* smem_u = smem_all + (threadIdx.x >> 5) * smemWordsPerWarp;
  // longest jobs first: all octave-0 blocks (four times the pixels) precede the octave-1 blocks in launch order
  const int nblk = (n + COMMIT_WPB - 1) / COMMIT_WPB;
  const int oct = (int)blockIdx.x / nblk, f = ((int)blockIdx.x - oct * nblk) * COMMIT_WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (oct >= g.noct || f >= n) return;
  const LineOct& O = g.o[oct];
  const int W = O.sw, H = O.sh, wpr = O.wpr;
  unsigned* ring = smem_u;
  GrowBitmap bm;
  bm.sm = smem_u + GROW_RQ;
  bm.gm = b.bitmap + (size_t)f * g.bmTotal + O.bmOff;
  bm.wpr = wpr;
  bm.top = 0;
  PhantomMap ph;
  ph.gm = b.phantom + (size_t)f * g.bmTotal + O.bmOff;
  ph.sm = bm.sm + GROW_K * wpr;
  ph.rm = ph.sm + GROW_K * wpr;
  ph.any = false;
  ph.box_reset();
  for (int i = lane; i < min(GROW_K, H) * wpr; i += 32) { bm.sm[i] = __ldcg(bm.gm + i); ph.sm[i] = 0u; }
  ph.rm[lane & (GROW_K - 1)] = 0u;
  __syncwarp();
  const size_t pbase = (size_t)f * g.pxTotal + O.pxOff;
  const float2* __restrict__ rec = b.cs + pbase;
  const float* __restrict__ ang = b.ang + pbase;
  const float2* __restrict__ seedcs = b.seed + pbase;
  unsigned* regAll = b.reg + (size_t)f * g.regTotal + O.regOff;
  LineRegion* rtab = b.regTab + (size_t)f * g.segTotal + O.segOff;
  const double prec = g.prec;
  const float kHi = g.alignHi2, kLo = g.alignLo2;
  int regBase = 0, nreg = 0;
  bool overflow = false;
  const int e = lane >> 3, k8 = lane & 7;
  const int nidx = k8 < 4 ? k8 : k8 + 1;
  const int ndx = nidx % 3 - 1, ndy = nidx / 3 - 1;
  const unsigned laneBit = 1u << lane;

  // speculative records of the band the seed scan is in; `cur` / `curPix` are fetched one
  // region ahead (record and the first 32 pixels of its list)
  int band = -1, bandEnd = 0, bp = 0, bcnt = 0, runStart = 0, bandList = 0;
  const uint4* brecs = nullptr;
  const unsigned* blist = nullptr;
  uint4 cur = make_uint4(0u, 0u, 0u, 0u);
  unsigned curPix = 0u;

  for (int row = 0; row < H - 1; row++) {
    if (row == bandEnd) {
      do {   // bands of equal load (k_lsd_band_split): a band without rows is skipped
        band++;
        bandEnd = b.bandRow[(size_t)f * (g.tasksPerFrame + 2) + O.taskOff + oct + band + 1];
      } while (bandEnd <= row && band + 1 < O.nbands);
      bp = 0; runStart = 0;
      bcnt = b.specCnt[(size_t)f * g.tasksPerFrame + O.taskOff + band];
      brecs = reinterpret_cast<const uint4*>(b.specRec + (size_t)f * g.specRecTotal + O.specRecOff + (size_t)band * O.bandRecCap);
      bandList = W * H + band * O.bandPxCap;
      blist = regAll + bandList;
      if (bcnt > 0) { cur = __ldcg(brecs); curPix = __ldcg(blist + lane); }
      ph.box_reset();   // phantoms of the previous band cannot have influenced this band's speculation
    }
    // slide the shared windows: rows [top, row) are exhausted, rows up to row + GROW_K enter
    if (row > bm.top) {
      const int r0 = bm.top + GROW_K, r1 = min(row + GROW_K, H);
      for (int r = r0; r < r1; r++) {
        unsigned rmask = 0u;
        for (int wv = lane; wv < wpr; wv += 32) {
          bm.sm[(r & (GROW_K - 1)) * wpr + wv] = __ldcg(bm.gm + r * wpr + wv);
          const unsigned pw = ph.any ? __ldcg(ph.gm + r * wpr + wv) : 0u;
          ph.sm[(r & (GROW_K - 1)) * wpr + wv] = pw;
          if (pw) rmask |= 1u << min(wv, 31);
        }
        rmask = __reduce_or_sync(0xffffffffu, rmask);
        if (lane == 0) ph.rm[r & (GROW_K - 1)] = rmask;
      }
      bm.top = row;
      __syncwarp();
    }
    for (int c0 = 0; c0 < wpr; c0 += 32) {
      while (true) {
        // next seed: first available pixel in raster order (src/LSD/lsd.cpp:476-479)
        const int wi = c0 + lane;
        const int rowBase = (row & (GROW_K - 1)) * wpr;
        const unsigned word = wi < wpr ? bm.sm[rowBase + wi] : 0u;
        const unsigned nz = __ballot_sync(0xffffffffu, word != 0u);
        if (!nz) break;
        const int wl = __ffs(nz) - 1;
        const unsigned sw_ = __shfl_sync(0xffffffffu, word, wl);
        const int bit = __ffs(sw_) - 1;
        const int sx = (c0 + wl) * 32 + bit, sy = row;
        const unsigned spk = (unsigned)sx | ((unsigned)sy << 16);

        // speculative regions seeded before this pixel never happened: their pixels that are
        // still available become phantoms
        while (bp < bcnt && cur.x < spk) {
          bool marked = false;
          int mx0 = 1 << 20, mx1 = -(1 << 20), my0 = 1 << 20, my1 = -(1 << 20);
          for (int i0 = 0; i0 < (int)cur.z; i0 += 32) {
            const int idx = i0 + lane;
            if (idx < (int)cur.z) {
              const unsigned q = i0 == 0 ? curPix : __ldcg(blist + cur.y + idx);
              const int qx = q & 0xffff, qy = q >> 16;
              if (qy >= bm.top && bm.test(qx, qy)) {
                ph.mark(bm, qx, qy); marked = true;
                mx0 = min(mx0, qx); mx1 = max(mx1, qx); my0 = min(my0, qy); my1 = max(my1, qy);
              }
            }
          }
          if (__any_sync(0xffffffffu, marked)) { ph.any = true; ph.box_add(mx0, mx1, my0, my1); }
          runStart += (int)cur.z;
          bp++;
          if (bp < bcnt) { cur = __ldcg(brecs + bp); curPix = __ldcg(blist + runStart + lane); }
          __syncwarp();
        }
        bool haveSpec = bp < bcnt && cur.x == spk;
        uint4 sr = cur;
        unsigned srPix = curPix;
        if (haveSpec) {
          runStart += (int)cur.z;
          bp++;
          if (bp < bcnt) { cur = __ldcg(brecs + bp); curPix = __ldcg(blist + runStart + lane); }
          bool ok = true;
          for (int i0 = 0; i0 < (int)sr.z && ok; i0 += 32) {
            const int idx = i0 + lane;
            bool good = true;
            if (idx < (int)sr.z) {
              const unsigned q = i0 == 0 ? srPix : __ldcg(blist + sr.y + idx);
              const int qx = q & 0xffff, qy = q >> 16;
              good = bm.test(qx, qy) && !(ph.in_box(qx, qy) && ph.near(bm, qx, qy, H));
            }
            ok = __all_sync(0xffffffffu, good);
          }
          if (ok) {   // adopt: this is the region the serial algorithm grows from this seed
            for (int i0 = 0; i0 < (int)sr.z; i0 += 32) {
              const int idx = i0 + lane;
              if (idx < (int)sr.z) {
                const unsigned q = i0 == 0 ? srPix : __ldcg(blist + sr.y + idx);
                grow_clear_atomic(bm, q & 0xffff, q >> 16);
              }
            }
            if ((int)sr.z >= O.minRegSize) {
              if (nreg < O.segCap) {
                if (lane == 0) {
                  LineRegion r;
                  r.start = bandList + (int)sr.y;
                  r.size = (int)sr.z;
                  r.angle = __dmul_rn((double)__uint_as_float(sr.w), D2R);
                  rtab[nreg] = r;
                }
                nreg++;
              } else {
                overflow = true;
              }
            }
            __syncwarp();
            continue;
          }
        }

        const int sp = sy * W + sx;
        const float sang = __ldg(ang + sp);
        const float2 scs = __ldg(seedcs + sp);
        unsigned* reg = regAll + regBase;          // pixel list of the region being grown
        if (lane == 0) {
          bm.sm[rowBase + c0 + wl] = sw_ & ~(1u << bit);
          ring[0] = spk;
        }
        __syncwarp();
        // region_grow (src/LSD/lsd.cpp:635-686), see k_lsd_grow
        const double seedAngle = __dmul_rn((double)sang, D2R);
        float sumdx = scs.x, sumdy = scs.y;
        bool fresh = true;
        int regSize = 1;
        int flushed = 0;
        int i = 0, nb = 1;
        GrowBatch cb = grow_fetch(bm, ring, reg, regSize, 0, 1, e, ndx, ndy, W, H, rec);
        bool cbDirty = false;   // a pixel was accepted after cb's candidates were fetched (only then are they tested again)
        while (nb > 0) {
          const int ni = i + nb, nnb = min(4, regSize - ni);
          GrowBatch nxt = grow_fetch(bm, ring, reg, regSize, ni, nnb, e, ndx, ndy, W, H, rec);
          unsigned pm = cb.mask;
          if (pm && cbDirty) pm = __ballot_sync(0xffffffffu, (pm & laneBit) && grow_bit(bm, cb.bw, cb.bbit));
          const int sizeBefore = regSize;
          float n2 = __fmaf_rn(sumdx, sumdx, sumdy * sumdy);
          while (pm) {
            const float dot = __fmaf_rn(sumdx, cb.rec.x, sumdy * cb.rec.y);
            const float d2 = dot * dot;
            const bool poss = (pm & laneBit) && dot > 0.f && d2 > kLo * n2;
            const unsigned possm = __ballot_sync(0xffffffffu, poss);
            if (!possm) break;
            const int l = __ffs(possm) - 1;
            const unsigned surem = __ballot_sync(0xffffffffu, poss && d2 >= kHi * n2);
            pm &= ~((2u << l) - 1u);
            if (!((surem >> l) & 1u)) {
              const double regAngle = fresh ? seedAngle : __dmul_rn((double)fast_atan2_dev(sumdy, sumdx), D2R);
              const int lpk = __shfl_sync(0xffffffffu, cb.cpk, l);
              const float la = __ldg(ang + (lpk >> 16) * W + (lpk & 0xffff));
              if (!is_aligned_dev(__dmul_rn((double)la, D2R), regAngle, prec)) continue;
            }
            const float qc = __shfl_sync(0xffffffffu, cb.rec.x, l), qs = __shfl_sync(0xffffffffu, cb.rec.y, l);
            const int qpk = __shfl_sync(0xffffffffu, cb.cpk, l);
            pm &= ~__ballot_sync(0xffffffffu, cb.cpk == qpk);
            if (lane == l) {
              grow_clear(bm, cb.bw, cb.bbit);
              ring[regSize & (GROW_RQ - 1)] = (unsigned)qpk;
            }
            regSize++;
            fresh = false;
            sumdx = __fadd_rn(sumdx, qc);
            sumdy = __fadd_rn(sumdy, qs);
            n2 = __fmaf_rn(sumdx, sumdx, sumdy * sumdy);
          }
          __syncwarp();
          while (regSize - flushed >= 32) {
            __stcg(reg + flushed + lane, ring[(flushed + lane) & (GROW_RQ - 1)]);
            flushed += 32;
          }
          i = ni;
          if (nnb > 0) {
            cb = nxt;
            nb = nnb;
            cbDirty = regSize != sizeBefore;   // fetched before this batch's accepts
          } else {
            nb = min(4, regSize - i);
            if (nb > 0) cb = grow_fetch(bm, ring, reg, regSize, i, nb, e, ndx, ndy, W, H, rec);
            cbDirty = false;
          }
        }
        if (regSize >= O.minRegSize) {
          if (nreg < O.segCap) {
            if (flushed + lane < regSize) __stcg(reg + flushed + lane, ring[(flushed + lane) & (GROW_RQ - 1)]);
            if (lane == 0) {
              LineRegion r;
              r.start = regBase;
              r.size = regSize;
              r.angle = fresh ? seedAngle : __dmul_rn((double)fast_atan2_dev(sumdy, sumdx), D2R);
              rtab[nreg] = r;
            }
            nreg++;
            regBase += regSize;
          } else {
            overflow = true;
          }
        }
        __syncwarp();
        if (haveSpec) {   // the discarded speculation of this seed: what it had taken beyond the true region
          bool marked = false;
          int mx0 = 1 << 20, mx1 = -(1 << 20), my0 = 1 << 20, my1 = -(1 << 20);
          for (int i0 = 0; i0 < (int)sr.z; i0 += 32) {
            const int idx = i0 + lane;
            if (idx < (int)sr.z) {
              const unsigned q = i0 == 0 ? srPix : __ldcg(blist + sr.y + idx);
              const int qx = q & 0xffff, qy = q >> 16;
              if (qy >= bm.top && bm.test(qx, qy)) {
                ph.mark(bm, qx, qy); marked = true;
                mx0 = min(mx0, qx); mx1 = max(mx1, qx); my0 = min(my0, qy); my1 = max(my1, qy);
              }
            }
          }
          if (__any_sync(0xffffffffu, marked)) { ph.any = true; ph.box_add(mx0, mx1, my0, my1); }
          __syncwarp();
        }
      }
    }
  }
  if (lane == 0) b.regCount[f * 2 + oct] = overflow ? -1 : nreg;
}

// ---------------------------------------------------------------------------------------
// Band-run region growing: the exact region sequence of the serial algorithm for SMALL batches, where the
// serial chains above leave the GPU idle (one frame = 20 chains).  The working image is cut into many bands
// (O.brBands); one warp owns one band and grows, with the serial building block of k_lsd_grow, every region
// seeded in its rows -- from an INPUT availability bitmap In_b that is a guess of the true state at the moment
// the raster scan reaches the band:   In_b = I_0 & ~(M_0 | ... | M_{b-1}),   M_j = pixels band j's latest run consumed
// (k_lsd_band_compose).  A band whose input changed runs again (k_lsd_band_run); its previous regions serve as
// speculation records and are adopted under the three conditions of k_lsd_commit (true next seed, all pixels
// still available, no phantom pixel in the 3x3 neighbourhood; the phantom map starts as In_new & ~In_old), so a
// re-run only re-grows what the changed input touches.  Rounds repeat until no input changes.  Exactness: band 0
// always runs on I_0, so its first result is the true one; by induction band b's input is the true state once
// bands 0..b-1 are final, and a round without any change is a fixed point in which every band ran on the true
// state.  At most brBands + 1 rounds are needed; if the rounds launched did not reach the fixed point (or a
// band overflowed its record / pixel capacity) the octave falls back to the serial kernel.
// ---------------------------------------------------------------------------------------
#define BR_ST 8   // ints of per-band state: nrec[0], nrec[1], cur, dirty, hasPrev
// bitmap rows of a band's shared-memory window (template parameter of k_lsd_band_run): 256 while the batch leaves
// the warp an SM's shared memory to itself (a window that holds nearly every region keeps the availability tests off
// the L2 round trip), 32 when many frames are in flight and resident warps per SM count more
#define BR_K_BIG 256
#ifndef BR_PREFETCH
// L1 prefetches of the per-pixel records in k_lsd_band_run: 2 / 1 = the lower neighbours of every accepted pixel (both
// sides / one line) + the band's rows in the prologue, 0 = the prologue only, -1 = none.  Measured on one frame (B200,
// end of round 2): 4.35 / 4.34 / 4.23 / 4.19 ms -- the loads they spare cost less than the instructions they add
#define BR_PREFETCH -1
#endif
#define BR_K_SMALL 32

// thread = one bitmap word (row, w) of one (frame, octave); walks the bands that contain the row in band order
__global__ void __launch_bounds__(256) k_lsd_band_compose(const __grid_constant__ LineGeom g, LineBufs b, int round, int check) {
  const int oct = blockIdx.y, f = blockIdx.z;
  if (oct >= g.noct) return;
  const LineOct& O = g.o[oct];
  int* flags = b.brFlags + ((size_t)f * 2 + oct) * BR_FLAGS;
  if (flags[0] || flags[1]) return;                       // serial fallback requested / fixed point reached
  if (round > 1 && flags[2 + round - 1] == 0) {           // nothing ran in the previous round: fixed point
    __syncthreads();                                      // (every thread has read the flags before one sets converged)
    if (threadIdx.x == 0 && blockIdx.x == 0) flags[1] = 1;
    return;
  }
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int nwords = O.sh * O.wpr;
  if (i >= nwords) return;
  const int row = i / O.wpr;
  const unsigned I0 = __ldg(b.bitmap + (size_t)f * g.bmTotal + O.bmOff + i);
  const size_t base = (size_t)f * g.brBmTotal + O.brBmOff + i;
  int* st = b.brState + ((size_t)f * g.brBandsPerFrame + O.brBandOff) * BR_ST;
  unsigned acc = 0u;
  bool anyChange = false;
  const int jmax = min(O.brBands, row / O.brRows + 1);   // the bands whose rows start at or above this row
  if (round == 1) {
    // nothing is known about what the bands consume yet
    for (int j = 0; j < jmax; j++) {
      const size_t idx = base + (size_t)j * nwords;
      b.brIn[idx] = I0;
      b.brPh[idx] = 0u;
      if (i == 0 || row == j * O.brRows) {   // one thread per band is enough; several writing the same values is harmless
        st[j * BR_ST + 0] = 0; st[j * BR_ST + 1] = 0; st[j * BR_ST + 2] = 0; st[j * BR_ST + 3] = 1; st[j * BR_ST + 4] = 0; st[j * BR_ST + 6] = 0;
      }
    }
    anyChange = jmax > 0;
  } else {
    // the words of four bands are requested together: the chain through acc is ALU work only (one L2 round trip per
    // four bands instead of one per band; the thread of the last row walks all 64 bands)
    for (int j0 = 0; j0 < jmax; j0 += 4) {
      unsigned oldv[4], wkv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const size_t idx = base + (size_t)min(j0 + u, jmax - 1) * nwords;
        oldv[u] = __ldcg(b.brIn + idx);
        wkv[u] = __ldcg(b.brWk + idx);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = j0 + u;
        if (j >= jmax) break;
        const size_t idx = base + (size_t)j * nwords;
        const unsigned nin = I0 & ~acc;
        const unsigned old = oldv[u];
        if (nin != old) {
          if (check) { flags[0] = 1; return; }
          b.brPh[idx] = nin & ~old;      // available now, was not when the band last ran
          b.brIn[idx] = nin;
          st[j * BR_ST + 3] = 1;
          anyChange = true;
        } else if (!check) {
          b.brPh[idx] = 0u;
        }
        acc |= old & ~wkv[u];            // what band j's latest run took (from the input that run saw)
      }
    }
  }
  if (anyChange && !check) flags[2 + round] = 1;
}

// (the small-window instantiation runs with many frames in flight: it keeps the 72 registers = 28 warps per SM it had
// before the prologue held eight words per lane in flight)
template <int BR_K> __global__ void __launch_bounds__(32, BR_K == BR_K_BIG ? 1 : 28) k_lsd_band_run(const __grid_constant__ LineGeom g, LineBufs b) {
  extern __shared__ unsigned smem_u[];
  const int t = blockIdx.x, f = blockIdx.y, lane = threadIdx.x;
  const int oct = (g.noct > 1 && t >= g.o[1].brBandOff) ? 1 : 0;
  const LineOct& O = g.o[oct];
  const int j = t - O.brBandOff;
  if (j >= O.brBands) return;
  int* flags = b.brFlags + ((size_t)f * 2 + oct) * BR_FLAGS;
  int* st = b.brState + ((size_t)f * g.brBandsPerFrame + t) * BR_ST;
  if (flags[0] || flags[1] || !st[3]) return;
  const long long tStart = clock64();
  const int W = O.sw, H = O.sh, wpr = O.wpr;
  const int r0 = j * O.brRows, r1 = min(r0 + O.brRows, H - 1);
  const int nwords = H * wpr;
  const size_t bmBase = (size_t)f * g.brBmTotal + O.brBmOff + (size_t)j * nwords;
  const unsigned* In = b.brIn + bmBase;
  unsigned* Wk = b.brWk + bmBase;
  unsigned* Ph = b.brPh + bmBase;
  const int cur = st[2], nxt = cur ^ 1;
  const int bcnt = st[4] ? st[cur] : 0;
  const uint4* brecs = b.brRec + ((size_t)f * 2 + cur) * g.brRecTotal + O.brRecOff + (size_t)j * O.brRecCap;
  uint4* nrecs = b.brRec + ((size_t)f * 2 + nxt) * g.brRecTotal + O.brRecOff + (size_t)j * O.brRecCap;
  const unsigned* blist = b.brList + ((size_t)f * 2 + cur) * g.brListTotal + O.brListOff + (size_t)j * O.brPxCap;
  unsigned* nlist = b.brList + ((size_t)f * 2 + nxt) * g.brListTotal + O.brListOff + (size_t)j * O.brPxCap;
  const int pxCap = O.brPxCap, recCap = O.brRecCap;

  const size_t pbase = (size_t)f * g.pxTotal + O.pxOff;
  const float2* __restrict__ rec = b.cs + pbase;
  const float* __restrict__ ang = b.ang + pbase;
  const float2* __restrict__ seedcs = b.seed + pbase;
#if BR_PREFETCH >= 0
  {
    // The warp has an SM sub-partition (and most of its L1) to itself: pull the per-pixel records of the band's own
    // rows and of the rows right below into L1 while the bitmaps are copied, so that the dependent loads of the
    // growth (seed angle -> neighbour records -> ...) hit L1 instead of paying an L2 round trip each.
    const int pr1 = min(r1 + 8, H);
    const char* c0 = reinterpret_cast<const char*>(rec + (size_t)r0 * W);
    const char* c1 = reinterpret_cast<const char*>(rec + (size_t)pr1 * W);
    for (const char* q = c0 + lane * 128; q < c1; q += 32 * 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
    const char* a0 = reinterpret_cast<const char*>(ang + (size_t)r0 * W);
    const char* a1 = reinterpret_cast<const char*>(ang + (size_t)min(r1 + 2, H) * W);
    for (const char* q = a0 + lane * 128; q < a1; q += 32 * 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
    const char* s0 = reinterpret_cast<const char*>(seedcs + (size_t)r0 * W);
    const char* s1 = reinterpret_cast<const char*>(seedcs + (size_t)r1 * W);
    for (const char* q = s0 + lane * 128; q < s1; q += 32 * 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
  }
#endif
  // working copy of the input (rows above the band hold nothing by definition), the first BR_K rows of it and of the
  // initial phantom map into the shared windows (a ring of BR_K rows: linear from the band's first row, wrapping once),
  // and: is there any initial phantom?  Eight (four) independent words per lane are in flight; one pass over the input.
  unsigned* ring = smem_u;
  GrowBitmapT<BR_K> bm;
  bm.sm = smem_u + GROW_RQ;
  bm.gm = Wk;
  bm.wpr = wpr;
  bm.top = r0;
  PhantomMapT<BR_K> ph;
  ph.gm = Ph;
  ph.sm = bm.sm + BR_K * wpr;
  bool anyPh = false;
  {
    const int first = r0 * wpr;
    const int winWords = (min(r0 + BR_K, H) - r0) * wpr, ringWords = BR_K * wpr;
    const int sbase = (r0 & (BR_K - 1)) * wpr - first;
    constexpr int U = BR_K == BR_K_BIG ? 8 : 4;
    for (int i0 = first + lane; i0 < nwords; i0 += 32 * U) {
      unsigned v[U], q[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int i = i0 + 32 * u;
        v[u] = i < nwords ? __ldcg(In + i) : 0u;
        q[u] = (i < nwords && bcnt > 0) ? __ldcg(Ph + i) : 0u;   // a first run has no record a phantom could invalidate
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int i = i0 + 32 * u;
        if (i < nwords) Wk[i] = v[u];
        anyPh |= q[u] != 0u;
        if (i - first < winWords) {
          int sidx = sbase + i;
          if (sidx >= ringWords) sidx -= ringWords;
          bm.sm[sidx] = v[u];
          ph.sm[sidx] = q[u];     // all zero when no phantom exists
        }
      }
    }
  }
  ph.any = bcnt > 0 && __any_sync(0xffffffffu, anyPh);
  __syncwarp();
  const double prec = g.prec;
  const float kHi = g.alignHi2, kLo = g.alignLo2;
  const int e = lane >> 3, k8 = lane & 7;
  const int nidx = k8 < 4 ? k8 : k8 + 1;
  const int ndx = nidx % 3 - 1, ndy = nidx / 3 - 1;
  const unsigned laneBit = 1u << lane;
  int bp = 0, runStart = 0;           // previous-run records: next record, start of its pixel list
  int nnew = 0, npx = 0;              // this run's records / pixels
  bool overflow = false;
  uint4 curRec = make_uint4(0u, 0u, 0u, 0u);
  unsigned curPix = 0u;
  if (bcnt > 0) { curRec = __ldcg(brecs); curPix = __ldcg(blist + lane); }

  for (int row = r0; row < r1 && !overflow; row++) {
    if (row > bm.top) {   // slide the shared windows
      const int ra = bm.top + BR_K, rb = min(row + BR_K, H);
      for (int r = ra; r < rb; r++)
        for (int wv = lane; wv < wpr; wv += 32) {
          bm.sm[(r & (BR_K - 1)) * wpr + wv] = __ldcg(bm.gm + r * wpr + wv);
          ph.sm[(r & (BR_K - 1)) * wpr + wv] = ph.any ? __ldcg(ph.gm + r * wpr + wv) : 0u;
        }
      bm.top = row;
      __syncwarp();
    }
    for (int c0 = 0; c0 < wpr && !overflow; c0 += 32) {
      while (true) {
        const int wi = c0 + lane;
        const int rowBase = (row & (BR_K - 1)) * wpr;
        const unsigned word = wi < wpr ? bm.sm[rowBase + wi] : 0u;
        const unsigned nz = __ballot_sync(0xffffffffu, word != 0u);
        if (!nz) break;
        const int wl = __ffs(nz) - 1;
        const unsigned sw_ = __shfl_sync(0xffffffffu, word, wl);
        const int bit = __ffs(sw_) - 1;
        const int sx = (c0 + wl) * 32 + bit, sy = row;
        const unsigned spk = (unsigned)sx | ((unsigned)sy << 16);
        if (nnew >= recCap || npx + 64 > pxCap) { overflow = true; break; }

        // records of the previous run seeded before this pixel never happen now: what they took and is still
        // available becomes phantom
        while (bp < bcnt && curRec.x < spk) {
          bool marked = false;
          for (int i0 = 0; i0 < (int)curRec.z; i0 += 32) {
            const int idx = i0 + lane;
            if (idx < (int)curRec.z) {
              const unsigned q = i0 == 0 ? curPix : __ldcg(blist + curRec.y + idx);
              const int qx = q & 0xffff, qy = q >> 16;
              if (qy >= bm.top && bm.test(qx, qy)) { ph.mark(bm, qx, qy); marked = true; }
            }
          }
          if (__any_sync(0xffffffffu, marked)) ph.any = true;
          runStart += (int)curRec.z;
          bp++;
          if (bp < bcnt) { curRec = __ldcg(brecs + bp); curPix = __ldcg(blist + runStart + lane); }
          __syncwarp();
        }
        const bool haveRec = bp < bcnt && curRec.x == spk;
        const uint4 sr = curRec;
        const unsigned srPix = curPix;
        if (haveRec) {
          runStart += (int)curRec.z;
          bp++;
          if (bp < bcnt) { curRec = __ldcg(brecs + bp); curPix = __ldcg(blist + runStart + lane); }
          bool ok = npx + (int)sr.z + 64 <= pxCap;
          for (int i0 = 0; i0 < (int)sr.z && ok; i0 += 32) {
            const int idx = i0 + lane;
            bool good = true;
            if (idx < (int)sr.z) {
              const unsigned q = i0 == 0 ? srPix : __ldcg(blist + sr.y + idx);
              const int qx = q & 0xffff, qy = q >> 16;
              good = bm.test(qx, qy) && !(ph.any && ph.near(bm, qx, qy, H));
            }
            ok = __all_sync(0xffffffffu, good);
          }
          if (ok) {   // adopt: the region the serial algorithm grows from this seed; carried into this run's lists
            for (int i0 = 0; i0 < (int)sr.z; i0 += 32) {
              const int idx = i0 + lane;
              if (idx < (int)sr.z) {
                const unsigned q = i0 == 0 ? srPix : __ldcg(blist + sr.y + idx);
                grow_clear_atomic(bm, q & 0xffff, q >> 16);
                __stcg(nlist + npx + idx, q);
              }
            }
            if (lane == 0) nrecs[nnew] = make_uint4(sr.x, (unsigned)npx, sr.z, sr.w);
            nnew++;
            npx += (int)sr.z;
            __syncwarp();
            continue;
          }
        }

        const int sp = sy * W + sx;
        const float sang = __ldg(ang + sp);
        const float2 scs = __ldg(seedcs + sp);
        unsigned* reg = nlist + npx;
        if (lane == 0) {
          bm.sm[rowBase + c0 + wl] = sw_ & ~(1u << bit);
          ring[0] = spk;
        }
        __syncwarp();
        // region_grow (src/LSD/lsd.cpp:635-686), see k_lsd_grow
        float sumdx = scs.x, sumdy = scs.y;
        bool fresh = true;
        int regSize = 1;
        int flushed = 0;
        int i = 0, nb = 1;
        GrowBatch cb = grow_fetch(bm, ring, reg, regSize, 0, 1, e, ndx, ndy, W, H, rec);
        bool cbDirty = false;   // a pixel was accepted after cb's candidates were fetched (only then are they tested again)
        while (nb > 0) {
          if (npx + regSize + 64 > pxCap) { overflow = true; break; }
          const int ni = i + nb, nnb = min(4, regSize - ni);
          GrowBatch nxtb = grow_fetch(bm, ring, reg, regSize, ni, nnb, e, ndx, ndy, W, H, rec);
          unsigned pm = cb.mask;
          if (pm && cbDirty) pm = __ballot_sync(0xffffffffu, (pm & laneBit) && grow_bit(bm, cb.bw, cb.bbit));
          const int sizeBefore = regSize;
          float n2 = __fmaf_rn(sumdx, sumdx, sumdy * sumdy);
          while (pm) {
            const float dot = __fmaf_rn(sumdx, cb.rec.x, sumdy * cb.rec.y);
            const float d2 = dot * dot;
            const bool poss = (pm & laneBit) && dot > 0.f && d2 > kLo * n2;
            const unsigned possm = __ballot_sync(0xffffffffu, poss);
            if (!possm) break;
            const int l = __ffs(possm) - 1;
            const unsigned surem = __ballot_sync(0xffffffffu, poss && d2 >= kHi * n2);
            pm &= ~((2u << l) - 1u);
            if (!((surem >> l) & 1u)) {
              const double regAngle = __dmul_rn((double)(fresh ? sang : fast_atan2_dev(sumdy, sumdx)), D2R);
              const int lpk = __shfl_sync(0xffffffffu, cb.cpk, l);
              const float la = __ldg(ang + (lpk >> 16) * W + (lpk & 0xffff));
              if (!is_aligned_dev(__dmul_rn((double)la, D2R), regAngle, prec)) continue;
            }
            const float qc = __shfl_sync(0xffffffffu, cb.rec.x, l), qs = __shfl_sync(0xffffffffu, cb.rec.y, l);
            const int qpk = __shfl_sync(0xffffffffu, cb.cpk, l);
            pm &= ~__ballot_sync(0xffffffffu, cb.cpk == qpk);
            if (lane == l) {
              grow_clear(bm, cb.bw, cb.bbit);
              ring[regSize & (GROW_RQ - 1)] = (unsigned)qpk;
              // this pixel is expanded a few queue entries from now: its lower neighbours' records are what that
              // expansion waits for (rows up to its own are in L1 already when it was reached from above)
              const int py = (qpk >> 16) + 1, px = qpk & 0xffff;
#if BR_PREFETCH == 2
              if (py < H) {
                const float2* pr = rec + py * W + px;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pr - 1));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pr + 1));
              }
#elif BR_PREFETCH == 1
              if (py < H) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + py * W + px));
#endif
            }
            regSize++;
            fresh = false;
            sumdx = __fadd_rn(sumdx, qc);
            sumdy = __fadd_rn(sumdy, qs);
            n2 = __fmaf_rn(sumdx, sumdx, sumdy * sumdy);
          }
          __syncwarp();
          while (regSize - flushed >= 32) {
            __stcg(reg + flushed + lane, ring[(flushed + lane) & (GROW_RQ - 1)]);
            flushed += 32;
          }
          i = ni;
          if (nnb > 0) {
            cb = nxtb;
            nb = nnb;
            cbDirty = regSize != sizeBefore;   // fetched before this batch's accepts
          } else {
            nb = min(4, regSize - i);
            if (nb > 0) cb = grow_fetch(bm, ring, reg, regSize, i, nb, e, ndx, ndy, W, H, rec);
            cbDirty = false;
          }
        }
        if (overflow) break;
        if (flushed + lane < regSize) __stcg(reg + flushed + lane, ring[(flushed + lane) & (GROW_RQ - 1)]);
        if (lane == 0) {
          const float a = fresh ? sang : (regSize >= O.minRegSize ? fast_atan2_dev(sumdy, sumdx) : 0.f);
          nrecs[nnew] = make_uint4(spk, (unsigned)npx, (unsigned)regSize, __float_as_uint(a));
        }
        nnew++;
        npx += regSize;
        __syncwarp();
        if (haveRec) {   // the record that failed: what it had taken beyond the region grown now
          bool marked = false;
          for (int i0 = 0; i0 < (int)sr.z; i0 += 32) {
            const int idx = i0 + lane;
            if (idx < (int)sr.z) {
              const unsigned q = i0 == 0 ? srPix : __ldcg(blist + sr.y + idx);
              const int qx = q & 0xffff, qy = q >> 16;
              if (qy >= bm.top && bm.test(qx, qy)) { ph.mark(bm, qx, qy); marked = true; }
            }
          }
          if (__any_sync(0xffffffffu, marked)) ph.any = true;
          __syncwarp();
        }
      }
    }
  }
  __syncwarp();
  if (overflow) {
    if (lane == 0) { flags[0] = 1; st[3] = 0; }
    return;
  }
  // output bitmap: the band's own rows are exhausted; rows of the shared window go back to global memory
  for (int i = r0 * wpr + lane; i < r1 * wpr; i += 32) Wk[i] = 0u;
  {
    const int ra = max(r1, bm.top), rb = min(bm.top + BR_K, H);   // rb - ra <= BR_K: the ring wraps at most once
    const int sbase = (ra & (BR_K - 1)) * wpr - ra * wpr, ringWords = BR_K * wpr;
    for (int i = ra * wpr + lane; i < rb * wpr; i += 32) {
      int sidx = sbase + i;
      if (sidx >= ringWords) sidx -= ringWords;
      Wk[i] = bm.sm[sidx];
    }
  }
  if (lane == 0) { st[nxt] = nnew; st[2] = nxt; st[3] = 0; st[4] = 1; st[5] = npx; st[6] += 1; st[7] = (int)(clock64() - tStart); }
}

// one CTA per (frame, octave): the regions of all bands, in band order, into the octave's region table
__global__ void __launch_bounds__(256) k_lsd_band_gather(const __grid_constant__ LineGeom g, LineBufs b) {
  const int oct = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
  if (oct >= g.noct) return;
  const LineOct& O = g.o[oct];
  const int* flags = b.brFlags + ((size_t)f * 2 + oct) * BR_FLAGS;
  if (flags[0]) return;   // the serial kernel produces this octave
  __shared__ int s_base, s_wsum[8], s_over;
  if (tid == 0) { s_base = 0; s_over = 0; }
  __syncthreads();
  LineRegion* rtab = b.regTab + (size_t)f * g.segTotal + O.segOff;
  const int* st = b.brState + ((size_t)f * g.brBandsPerFrame + O.brBandOff) * BR_ST;
  for (int j = 0; j < O.brBands; j++) {
    const int cur = st[j * BR_ST + 2], n = st[j * BR_ST + cur];
    const uint4* recs = b.brRec + ((size_t)f * 2 + cur) * g.brRecTotal + O.brRecOff + (size_t)j * O.brRecCap;
    const size_t listOff = (size_t)cur * g.brListTotal + O.brListOff + (size_t)j * O.brPxCap;   // relative to the frame's lists
    for (int i0 = 0; i0 < n; i0 += 256) {
      const int i = i0 + tid;
      uint4 r = make_uint4(0u, 0u, 0u, 0u);
      bool keep = false;
      if (i < n) { r = recs[i]; keep = (int)r.z >= O.minRegSize; }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      const int lane = tid & 31, wid = tid >> 5;
      if (lane == 0) s_wsum[wid] = __popc(m);
      __syncthreads();
      int before = s_base;
      for (int w = 0; w < wid; w++) before += s_wsum[w];
      const int slot = before + __popc(m & ((1u << lane) - 1u));
      if (keep) {
        if (slot < O.segCap) {
          LineRegion R;
          R.start = (int)(listOff + r.y);
          R.size = (int)r.z;
          R.angle = __dmul_rn((double)__uint_as_float(r.w), D2R);
          rtab[slot] = R;
        } else {
          s_over = 1;
        }
      }
      __syncthreads();
      if (tid == 0) { int tot = 0; for (int w = 0; w < 8; w++) tot += s_wsum[w]; s_base += tot; }
      __syncthreads();
    }
  }
  if (tid == 0) b.regCount[f * 2 + oct] = s_over ? -1 : s_base;
}

// ---------------------------------------------------------------------------------------
// k_lsd_rect: warp per region.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// reductions inside groups of G consecutive lanes (xor butterfly: every lane of a group ends with the same value)
template <int G> __device__ __forceinline__ double group_sum_d(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int G> __device__ __forceinline__ double group_min_d(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int G> __device__ __forceinline__ double group_max_d(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#ifndef RECT_G
#define RECT_G 8    // 32: 7.6 ms, 16: 5.35 ms, 8: 4.8 ms per 4096 frames
#endif
// G lanes per region (32 / G regions per warp at a time): most regions are a few dozen pixels, and the fixed part of a
// region -- eight reductions, the square root, sincos, the divisions -- is paid per warp instruction, not per region.
template <int G>
__global__ void __launch_bounds__(256) k_lsd_rect(const __grid_constant__ LineGeom g, LineBufs b, int bandRun) {
  const int oct = blockIdx.x, f = blockIdx.y, part = blockIdx.z, nparts = gridDim.z;
  if (oct >= g.noct) return;
  const LineOct& O = g.o[oct];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int sub = lane & (G - 1), grp = lane / G, gbase = lane - sub, ngrp = 32 / G;
  const int nreg = b.regCount[f * 2 + oct];
  const size_t pbase = (size_t)f * g.pxTotal + O.pxOff;
  // band-run path: the region table indexes the frame's band lists, unless the octave fell back to the serial kernel
  const unsigned* reg = (bandRun && !b.brFlags[((size_t)f * 2 + oct) * BR_FLAGS])
                            ? b.brList + (size_t)f * 2 * g.brListTotal
                            : b.reg + (size_t)f * g.regTotal + O.regOff;
  const double* mod = b.mod + pbase;
  const LineRegion* rtab = b.regTab + (size_t)f * g.segTotal + O.segOff;
  float4* segs = b.segs + (size_t)f * g.segTotal + O.segOff;
  const int W = O.sw;
  // all lanes of a warp make the same number of trips (the shuffles are warp-wide); a group without a region idles
  for (int r0 = (part * nw + wid) * ngrp; r0 < nreg; r0 += nparts * nw * ngrp) {
    const int r = r0 + grp;
    const bool have = r < nreg;
    LineRegion R;
    R.start = 0; R.size = 0; R.angle = 0.0;
    if (have) R = rtab[r];
    const unsigned* rp = reg + R.start;
    double x = 0, y = 0, sum = 0;
    for (int i = sub; i < R.size; i += G) {
      const unsigned p = rp[i];
      const int px = p & 0xffff, py = p >> 16;
      const double wgt = mod[py * W + px];
      x += (double)px * wgt;
      y += (double)py * wgt;
      sum += wgt;
    }
    x = group_sum_d<G>(x); y = group_sum_d<G>(y); sum = group_sum_d<G>(sum);
    {   // x / sum and y / sum in one division sequence (odd lanes take y)
      const double q = ((lane & 1) ? y : x) / sum;
      x = __shfl_sync(0xffffffffu, q, gbase);
      y = __shfl_sync(0xffffffffu, q, gbase + 1);
    }
    double Ixx = 0, Iyy = 0, Ixy = 0;
    for (int i = sub; i < R.size; i += G) {
      const unsigned p = rp[i];
      const int px = p & 0xffff, py = p >> 16;
      const double wgt = mod[py * W + px];
      const double dx = (double)px - x, dy = (double)py - y;
      Ixx += dy * dy * wgt;
      Iyy += dx * dx * wgt;
      Ixy -= dx * dy * wgt;
    }
    Ixx = group_sum_d<G>(Ixx); Iyy = group_sum_d<G>(Iyy); Ixy = group_sum_d<G>(Ixy);
    const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_dev((float)(lambda - Ixx), (float)Ixy)
                                           : (double)fast_atan2_dev((float)Ixy, (float)(lambda - Iyy));
    theta *= D2R;
    if (have) {
      double diff = theta - R.angle;
      while (diff <= -PI_D) diff += 2 * PI_D;
      while (diff > PI_D) diff -= 2 * PI_D;
      if (fabs(diff) > g.prec) theta += PI_D;
    } else {
      theta = 0.0;   // (an idle group: keep the argument of sincos finite)
    }
    double dx, dy;
    sincos(theta, &dy, &dx);
    double lmin = 0, lmax = 0;
    for (int i = sub; i < R.size; i += G) {
      const unsigned p = rp[i];
      const double rdx = (double)(p & 0xffff) - x, rdy = (double)(p >> 16) - y;
      const double l = __dadd_rn(__dmul_rn(rdx, dx), __dmul_rn(rdy, dy));
      lmax = l > lmax ? l : lmax;   // (l is never NaN: plain selects instead of the NaN-aware fmax / fmin sequences)
      lmin = l < lmin ? l : lmin;
    }
    lmin = group_min_d<G>(lmin);
    lmax = group_max_d<G>(lmax);
    if (have && sub < 4) {   // lanes 0..3 of the group finish x1, y1, x2, y2: one division sequence instead of four
      const double l = (sub & 2) ? lmax : lmin;
      const double c = (sub & 1) ? y : x, d = (sub & 1) ? dy : dx;
      double v = c + l * d;
      v += 0.5;
      if (g.lsdScale != 1.0) v /= g.lsdScale;
      reinterpret_cast<float*>(segs + r)[sub] = (float)v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// k_lsd_grow_refine: lsd_refine > 0 (LSD_REFINE_STD = 1, LSD_REFINE_ADV = 2), src/LSD/lsd.cpp:493-504,784-1134.  refine()
// gives pixels back to the "not used" state and re-grows regions with another tolerance, so the outcome of one seed
// feeds the next: the whole seed loop -- region_grow, region2rect, refine, reduce_region_radius, rect_improve / rect_nfa /
// nfa -- runs in one warp per (frame, octave), in the reference's order; the segments come straight out of this kernel
// (k_lsd_rect is skipped).  Lanes share the per-pixel work of a region (sums, extents, scan lines of the NFA rectangle).
// The f64 sums of a region are tree sums over lanes, i.e. not in the reference's left-to-right order: decisions and the
// float end points agree with the reference except where a comparison or a float rounding falls within ~1e-13 of a tie
// (same as k_lsd_rect).  The NFA uses CUDA's log / exp / pow / sinh / log10 (<= 2 ulp) where the reference uses libm.
// ---------------------------------------------------------------------------------------
struct LsdRectD { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };

__device__ __forceinline__ double angle_diff_signed_dev(double a, double b) {
  double diff = a - b;
  while (diff <= -PI_D) diff += 2 * PI_D;
  while (diff > PI_D) diff -= 2 * PI_D;
  return diff;
}

// region2rect + get_theta for the pixel list rp[0 .. size) (all lanes return the same rectangle)
__device__ void warp_region2rect(const unsigned* rp, int size, const double* __restrict__ mod, int W, double regAngle, double prec, double p,
                                 LsdRectD& rec, int lane) {
  double x = 0, y = 0, sum = 0;
  for (int i = lane; i < size; i += 32) {
    const unsigned q = __ldcg(rp + i);
    const int px = q & 0xffff, py = q >> 16;
    const double wgt = mod[py * W + px];
    x += (double)px * wgt;
    y += (double)py * wgt;
    sum += wgt;
  }
  x = warp_sum_d(x); y = warp_sum_d(y); sum = warp_sum_d(sum);
  x /= sum;
  y /= sum;
  double Ixx = 0, Iyy = 0, Ixy = 0;
  for (int i = lane; i < size; i += 32) {
    const unsigned q = __ldcg(rp + i);
    const int px = q & 0xffff, py = q >> 16;
    const double wgt = mod[py * W + px];
    const double dx = (double)px - x, dy = (double)py - y;
    Ixx += dy * dy * wgt;
    Iyy += dx * dx * wgt;
    Ixy -= dx * dy * wgt;
  }
  Ixx = warp_sum_d(Ixx); Iyy = warp_sum_d(Iyy); Ixy = warp_sum_d(Ixy);
  const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_dev((float)(lambda - Ixx), (float)Ixy)
                                         : (double)fast_atan2_dev((float)Ixy, (float)(lambda - Iyy));
  theta *= D2R;
  if (fabs(angle_diff_signed_dev(theta, regAngle)) > prec) theta += PI_D;
  double dx, dy;
  sincos(theta, &dy, &dx);
  double lmin = 0, lmax = 0, wmin = 0, wmax = 0;
  for (int i = lane; i < size; i += 32) {
    const unsigned q = __ldcg(rp + i);
    const double rdx = (double)(q & 0xffff) - x, rdy = (double)(q >> 16) - y;
    const double l = __dadd_rn(__dmul_rn(rdx, dx), __dmul_rn(rdy, dy));
    const double w = __dadd_rn(__dmul_rn(-rdx, dy), __dmul_rn(rdy, dx));
    lmax = fmax(lmax, l); lmin = fmin(lmin, l);
    wmax = fmax(wmax, w); wmin = fmin(wmin, w);
  }
  lmin = warp_min_d(lmin); lmax = warp_max_d(lmax);
  wmin = warp_min_d(wmin); wmax = warp_max_d(wmax);
  rec.x1 = x + lmin * dx; rec.y1 = y + lmin * dy;
  rec.x2 = x + lmax * dx; rec.y2 = y + lmax * dy;
  rec.width = wmax - wmin;
  rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
  if (rec.width < 1.0) rec.width = 1.0;
}

__device__ __forceinline__ double rect_density_dev(const LsdRectD& r, int size) {
  const double ddx = r.x2 - r.x1, ddy = r.y2 - r.y1;
  return (double)size / (sqrt(ddx * ddx + ddy * ddy) * r.width);
}

__device__ double log_gamma_dev(double x) {
  if (x > 15.0) return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
  const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
  double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
  double bsum = 0;
  for (int n = 0; n < 7; ++n) {
    a -= log(x + (double)n);
    bsum += q[n] * pow(x, (double)n);
  }
  return a + log(bsum);
}

__device__ double nfa_dev(int n, int k, double p, double LOG_NT) {
  if (n == 0 || k == 0) return -LOG_NT;
  if (n == k) return -LOG_NT - (double)n * log10(p);
  const double p_term = p / (1 - p);
  const double log1term = ((double)n + 1) - log_gamma_dev((double)k + 1) - log_gamma_dev((double)(n - k) + 1) + (double)k * log(p) +
                          (double)(n - k) * log(1.0 - p);
  double term = exp(log1term);
  {
    const double aa = fabs(term);
    const double abs_max = aa < 2.2250738585072014e-308 ? 2.2250738585072014e-308 : aa;
    if (term == 0.0 || (aa / abs_max) <= (100.0 * 2.220446049250313e-16)) {   // double_equal(term, 0)
      if (k > n * p) return -log1term / 2.30258509299404568402 - LOG_NT;
      return -LOG_NT;
    }
  }
  double bin_tail = term;
  for (int i = k + 1; i <= n; ++i) {
    const double bin_term = (double)(n - i + 1) / (double)i;
    const double mult_term = bin_term * p_term;
    term *= mult_term;
    bin_tail += term;
    if (bin_term < 1) {
      const double err = term * ((1 - pow(mult_term, (double)(n - i + 1))) / (1 - mult_term) - 1);
      if (err < 0.1 * fabs(-log10(bin_tail) - LOG_NT) * bin_tail) break;
    }
  }
  return -log10(bin_tail) - LOG_NT;
}

// rect_nfa: the scan conversion runs redundantly in every lane (a handful of scalar steps per row); the pixels of a row
// are shared among the lanes
__device__ double warp_rect_nfa(const LsdRectD& rec, const float* __restrict__ ang, int W, int H, double LOG_NT, int lane) {
  const double half_width = rec.width / 2.0;
  const double dyhw = rec.dy * half_width, dxhw = rec.dx * half_width;
  int ex[4] = {(int)(rec.x1 - dyhw), (int)(rec.x2 - dyhw), (int)(rec.x2 + dyhw), (int)(rec.x1 + dyhw)};
  int ey[4] = {(int)(rec.y1 + dxhw), (int)(rec.y2 + dxhw), (int)(rec.y2 - dxhw), (int)(rec.y1 - dxhw)};
  // sort by (x, y): 4 elements, insertion sort
#pragma unroll
  for (int i = 1; i < 4; i++)
#pragma unroll
    for (int j = i; j > 0; j--) {
      const bool lt = ex[j] == ex[j - 1] ? ey[j] < ey[j - 1] : ex[j] < ex[j - 1];
      if (lt) { int t = ex[j]; ex[j] = ex[j - 1]; ex[j - 1] = t; t = ey[j]; ey[j] = ey[j - 1]; ey[j - 1] = t; }
    }
  int iMin = 0, iMax = 0;
#pragma unroll
  for (int i = 1; i < 4; i++) {
    if (ey[iMin] > ey[i]) iMin = i;
    if (ey[iMax] < ey[i]) iMax = i;
  }
  unsigned taken = 1u << iMin;
  int iL = -1, iR = -1, iT = -1;
#pragma unroll
  for (int i = 0; i < 4; i++) if (!((taken >> i) & 1u) && (iL < 0 || ex[iL] > ex[i])) iL = i;
  taken |= 1u << iL;
#pragma unroll
  for (int i = 0; i < 4; i++) if (!((taken >> i) & 1u) && (iR < 0 || ex[iR] < ex[i])) iR = i;
  taken |= 1u << iR;
#pragma unroll
  for (int i = 0; i < 4; i++) if (!((taken >> i) & 1u) && (iT < 0 || ex[iT] > ex[i])) iT = i;
  const int mx = ex[iMin], my = ey[iMin], lx = ex[iL], ly = ey[iL], rx = ex[iR], ry = ey[iR], tx = ex[iT];
  // INTEGER quotients, and the tests against tailp's x, as in the vendored code (:1055-1063)
  const double flstep = (my != ly) ? (double)((mx - lx) / (my - ly)) : 0.0;
  const double slstep = (ly != tx) ? (double)((lx - tx) / (ly - tx)) : 0.0;
  const double frstep = (my != ry) ? (double)((mx - rx) / (my - ry)) : 0.0;
  const double srstep = (ry != tx) ? (double)((rx - tx) / (ry - tx)) : 0.0;
  double lstep = flstep, rstep = frstep;
  double left_x = mx, right_x = mx;
  int total = 0, alg = 0;
  const int yEnd = ey[iMax];
  for (int y = my; y <= yEnd; ++y) {
    if (y < 0 || y >= H) continue;
    const int xa = max((int)left_x, 0), xb = min((int)right_x, W - 1);
    for (int x = xa + lane; x <= xb; x += 32) {
      ++total;
      const float a = __ldg(ang + y * W + x);
      if (a != -1024.f && is_aligned_dev(__dmul_rn((double)a, D2R), rec.theta, rec.prec)) ++alg;
    }
    if (y >= ly) lstep = slstep;
    if (y >= ry) rstep = srstep;
    left_x += lstep;
    right_x += rstep;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    total += __shfl_xor_sync(0xffffffffu, total, o);
    alg += __shfl_xor_sync(0xffffffffu, alg, o);
  }
  return nfa_dev(total, alg, rec.p, LOG_NT);
}

__device__ double warp_rect_improve(LsdRectD& rec, const float* __restrict__ ang, int W, int H, double LOG_NT, double logEps, int lane) {
  const double delta = 0.5, delta_2 = delta / 2.0;
  double log_nfa = warp_rect_nfa(rec, ang, W, H, LOG_NT, lane);
  if (log_nfa > logEps) return log_nfa;
  LsdRectD r = rec;
  for (int n = 0; n < 5; ++n) {
    r.p /= 2;
    r.prec = r.p * PI_D;
    const double v = warp_rect_nfa(r, ang, W, H, LOG_NT, lane);
    if (v > log_nfa) { log_nfa = v; rec = r; }
  }
  if (log_nfa > logEps) return log_nfa;
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.width -= delta;
      const double v = warp_rect_nfa(r, ang, W, H, LOG_NT, lane);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  if (log_nfa > logEps) return log_nfa;
  for (int side = 0; side < 2; ++side) {
    r = rec;
    for (int n = 0; n < 5; ++n) {
      if ((r.width - delta) >= 0.5) {
        if (side == 0) { r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2; r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2; }
        else { r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2; r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2; }
        r.width -= delta;
        const double v = warp_rect_nfa(r, ang, W, H, LOG_NT, lane);
        if (v > log_nfa) { rec = r; log_nfa = v; }
      }
    }
    if (log_nfa > logEps) return log_nfa;
  }
  r = rec;
  for (int n = 0; n < 5; ++n) {
    if ((r.width - delta) >= 0.5) {
      r.p /= 2;
      r.prec = r.p * PI_D;
      const double v = warp_rect_nfa(r, ang, W, H, LOG_NT, lane);
      if (v > log_nfa) { rec = r; log_nfa = v; }
    }
  }
  return log_nfa;
}

// region_grow from (sx, sy) with tolerance prec into reg[0 ..): the building block of k_lsd_grow with the tolerance as
// a parameter.  exactOnly: the dot-product pre-test is skipped (tolerances near or beyond 90 degrees, NaN) and every
// available neighbour takes the reference's f64 test.  The seed's bit must be available on entry.  Returns the size;
// regAngleDeg = the float region angle in degrees (fresh: the seed's own angle).
__device__ int warp_region_grow(GrowBitmap& bm, unsigned* ring, unsigned* reg, int sx, int sy, int W, int H, const float2* __restrict__ rec,
                                const float* __restrict__ ang, const float2* __restrict__ seedcs, double prec, float kHi, float kLo,
                                bool exactOnly, int lane, double& regAngle) {
  const int e = lane >> 3, k8 = lane & 7;
  const int nidx = k8 < 4 ? k8 : k8 + 1;
  const int ndx = nidx % 3 - 1, ndy = nidx / 3 - 1;
  const unsigned laneBit = 1u << lane;
  const int sp = sy * W + sx;
  const float sang = __ldg(ang + sp);
  const float2 scs = __ldg(seedcs + sp);
  if (lane == 0) {
    bm.clear(sx, sy);
    ring[0] = (unsigned)sx | ((unsigned)sy << 16);
  }
  __syncwarp();
  const double seedAngle = __dmul_rn((double)sang, D2R);
  float sumdx = scs.x, sumdy = scs.y;
  bool fresh = true;
  int regSize = 1, flushed = 0;
  int i = 0, nb = 1;
  GrowBatch cur = grow_fetch(bm, ring, reg, regSize, 0, 1, e, ndx, ndy, W, H, rec);
  while (nb > 0) {
    const int ni = i + nb, nnb = min(4, regSize - ni);
    GrowBatch nxt = grow_fetch(bm, ring, reg, regSize, ni, nnb, e, ndx, ndy, W, H, rec);
    unsigned pm = cur.mask;
    if (pm) pm = __ballot_sync(0xffffffffu, (pm & laneBit) && grow_bit(bm, cur.bw, cur.bbit));
    float n2 = __fmaf_rn(sumdx, sumdx, sumdy * sumdy);
    while (pm) {
      const float dot = __fmaf_rn(sumdx, cur.rec.x, sumdy * cur.rec.y);
      const float d2 = dot * dot;
      const bool poss = (pm & laneBit) && (exactOnly || (dot > 0.f && d2 > kLo * n2));
      const unsigned possm = __ballot_sync(0xffffffffu, poss);
      if (!possm) break;
      const int l = __ffs(possm) - 1;
      const unsigned surem = __ballot_sync(0xffffffffu, poss && !exactOnly && d2 >= kHi * n2);
      pm &= ~((2u << l) - 1u);
      if (!((surem >> l) & 1u)) {
        const double ra = fresh ? seedAngle : __dmul_rn((double)fast_atan2_dev(sumdy, sumdx), D2R);
        const int lpk = __shfl_sync(0xffffffffu, cur.cpk, l);
        const float la = __ldg(ang + (lpk >> 16) * W + (lpk & 0xffff));
        if (!is_aligned_dev(__dmul_rn((double)la, D2R), ra, prec)) continue;
      }
      const float qc = __shfl_sync(0xffffffffu, cur.rec.x, l), qs = __shfl_sync(0xffffffffu, cur.rec.y, l);
      const int qpk = __shfl_sync(0xffffffffu, cur.cpk, l);
      pm &= ~__ballot_sync(0xffffffffu, cur.cpk == qpk);
      if (lane == l) {
        grow_clear(bm, cur.bw, cur.bbit);
        ring[regSize & (GROW_RQ - 1)] = (unsigned)qpk;
      }
      regSize++;
      fresh = false;
      sumdx = __fadd_rn(sumdx, qc);
      sumdy = __fadd_rn(sumdy, qs);
      n2 = __fmaf_rn(sumdx, sumdx, sumdy * sumdy);
    }
    __syncwarp();
    while (regSize - flushed >= 32) {
      __stcg(reg + flushed + lane, ring[(flushed + lane) & (GROW_RQ - 1)]);
      flushed += 32;
    }
    i = ni;
    if (nnb > 0) {
      cur = nxt;
      nb = nnb;
    } else {
      nb = min(4, regSize - i);
      if (nb > 0) cur = grow_fetch(bm, ring, reg, regSize, i, nb, e, ndx, ndy, W, H, rec);
    }
  }
  if (flushed + lane < regSize) __stcg(reg + flushed + lane, ring[(flushed + lane) & (GROW_RQ - 1)]);
  __syncwarp();
  regAngle = fresh ? seedAngle : __dmul_rn((double)fast_atan2_dev(sumdy, sumdx), D2R);
  return regSize;
}

// a pixel back to "not used" (src/LSD/lsd.cpp:799,851): several lanes may hit one word
__device__ __forceinline__ void grow_release(GrowBitmap& bm, int x, int y) {
  const unsigned bit = 1u << (x & 31);
  if (y - bm.top < GROW_K) atomicOr(&bm.sm[(y & (GROW_K - 1)) * bm.wpr + (x >> 5)], bit);
  else atomicOr(bm.gm + y * bm.wpr + (x >> 5), bit);
}

__global__ void __launch_bounds__(32) k_lsd_grow_refine(const __grid_constant__ LineGeom g, LineBufs b, int refine, double logEps,
                                                        double densityTh) {
  extern __shared__ unsigned smem_u[];
  const int oct = blockIdx.x, f = blockIdx.y, lane = threadIdx.x;
  if (oct >= g.noct) return;
  const LineOct& O = g.o[oct];
  const int W = O.sw, H = O.sh, wpr = O.wpr;
  unsigned* ring = smem_u;
  GrowBitmap bm;
  bm.sm = smem_u + GROW_RQ;
  bm.gm = b.bitmap + (size_t)f * g.bmTotal + O.bmOff;
  bm.wpr = wpr;
  bm.top = 0;
  for (int i = lane; i < min(GROW_K, H) * wpr; i += 32) bm.sm[i] = __ldcg(bm.gm + i);
  __syncwarp();
  const size_t pbase = (size_t)f * g.pxTotal + O.pxOff;
  const float2* __restrict__ rec = b.cs + pbase;
  const float* __restrict__ ang = b.ang + pbase;
  const float2* __restrict__ seedcs = b.seed + pbase;
  const double* __restrict__ mod = b.mod + pbase;
  unsigned* reg = b.reg + (size_t)f * g.regTotal + O.regOff;      // one region at a time
  float4* segs = b.segs + (size_t)f * g.segTotal + O.segOff;
  const double prec = g.prec, p = 22.5 / 180;
  const double LOG_NT = 5 * (log10((double)W) + log10((double)H)) / 2 + log10(11.0);
  int nseg = 0;
  bool overflow = false;

  for (int row = 0; row < H - 1; row++) {
    if (row > bm.top) {   // slide the shared window: rows [top, row) are exhausted
      // rows that leave the window go back to global memory first (refine may have released pixels in them -- they are
      // above `row`, hence exhausted: nothing to write), rows entering it are read
      const int r0 = bm.top + GROW_K, r1 = min(row + GROW_K, H);
      for (int r = r0; r < r1; r++)
        for (int wv = lane; wv < wpr; wv += 32) bm.sm[(r & (GROW_K - 1)) * wpr + wv] = __ldcg(bm.gm + r * wpr + wv);
      bm.top = row;
      __syncwarp();
    }
    for (int c0 = 0; c0 < wpr; c0 += 32) {
      while (true) {
        const int wi = c0 + lane;
        const int rowBase = (row & (GROW_K - 1)) * wpr;
        const unsigned word = wi < wpr ? bm.sm[rowBase + wi] : 0u;
        const unsigned nz = __ballot_sync(0xffffffffu, word != 0u);
        if (!nz) break;
        const int wl = __ffs(nz) - 1;
        const unsigned sw_ = __shfl_sync(0xffffffffu, word, wl);
        const int sx = (c0 + wl) * 32 + (__ffs(sw_) - 1), sy = row;
        double regAngle;
        int regSize = warp_region_grow(bm, ring, reg, sx, sy, W, H, rec, ang, seedcs, prec, g.alignHi2, g.alignLo2, false, lane, regAngle);
        if (regSize < O.minRegSize) continue;
        LsdRectD R;
        warp_region2rect(reg, regSize, mod, W, regAngle, prec, p, R, lane);
        // ---- refine (:784-829)
        bool keep = true;
        double density = rect_density_dev(R, regSize);
        if (density < densityTh) {
          const double xc = (double)sx, yc = (double)sy;
          const double angC = __dmul_rn((double)__ldg(ang + sy * W + sx), D2R);
          double sum = 0, ssum = 0;
          int n = 0;
          for (int i = lane; i < regSize; i += 32) {
            const unsigned q = __ldcg(reg + i);
            const int qx = q & 0xffff, qy = q >> 16;
            grow_release(bm, qx, qy);
            const double ddx = (double)qx - xc, ddy = (double)qy - yc;
            if (sqrt(ddx * ddx + ddy * ddy) < R.width) {
              const double d = angle_diff_signed_dev(__dmul_rn((double)__ldg(ang + qy * W + qx), D2R), angC);
              sum += d;
              ssum += d * d;
              ++n;
            }
          }
          sum = warp_sum_d(sum); ssum = warp_sum_d(ssum);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
          __syncwarp();
          const double mean = sum / (double)n;
          const double tau = 2.0 * sqrt((ssum - 2.0 * mean * sum) / (double)n + mean * mean);
          const bool exactOnly = !(tau + 2e-3 < 1.5);
          const double cH = cos(fmax(tau - 2e-3, 0.0)), cL = cos(tau + 2e-3);
          regSize = warp_region_grow(bm, ring, reg, sx, sy, W, H, rec, ang, seedcs, tau, (float)(cH * cH), (float)(cL * cL), exactOnly, lane,
                                     regAngle);
          if (regSize < 2) continue;
          warp_region2rect(reg, regSize, mod, W, regAngle, prec, p, R, lane);
          density = rect_density_dev(R, regSize);
          if (density < densityTh) {   // reduce_region_radius (:831-869)
            const double a1 = (R.x1 - xc) * (R.x1 - xc) + (R.y1 - yc) * (R.y1 - yc);
            const double a2 = (R.x2 - xc) * (R.x2 - xc) + (R.y2 - yc) * (R.y2 - yc);
            double radSq = a1 > a2 ? a1 : a2;
            while (density < densityTh) {
              radSq *= 0.75 * 0.75;
              int kept = 0;     // order-preserving compaction of the list (the reference swaps the last point in: only the
                                // order of the sums differs)
              for (int i0 = 0; i0 < regSize; i0 += 32) {
                const int i = i0 + lane;
                unsigned q = 0u;
                bool in = false;
                if (i < regSize) {
                  q = __ldcg(reg + i);
                  const double ddx = (double)(q & 0xffff) - xc, ddy = (double)(q >> 16) - yc;
                  in = !(ddx * ddx + ddy * ddy > radSq);
                  if (!in) grow_release(bm, q & 0xffff, q >> 16);
                }
                const unsigned m = __ballot_sync(0xffffffffu, in);
                __syncwarp();
                if (in) __stcg(reg + kept + __popc(m & ((1u << lane) - 1u)), q);
                kept += __popc(m);
                __syncwarp();
              }
              regSize = kept;
              if (regSize < 2) { keep = false; break; }
              warp_region2rect(reg, regSize, mod, W, regAngle, prec, p, R, lane);
              density = rect_density_dev(R, regSize);
            }
          }
        }
        if (!keep) continue;
        if (refine >= 2) {
          const double logNfa = warp_rect_improve(R, ang, W, H, LOG_NT, logEps, lane);
          if (logNfa <= logEps) continue;
        }
        if (nseg < O.segCap) {
          if (lane < 4) {
            double v = (lane & 2) ? ((lane & 1) ? R.y2 : R.x2) : ((lane & 1) ? R.y1 : R.x1);
            v += 0.5;
            if (g.lsdScale != 1.0) v /= g.lsdScale;
            reinterpret_cast<float*>(segs + nseg)[lane] = (float)v;
          }
          nseg++;
        } else {
          overflow = true;
        }
      }
    }
  }
  if (lane == 0) b.regCount[f * 2 + oct] = overflow ? -1 : nseg;
}

// ---------------------------------------------------------------------------------------
// k_line_assemble: one CTA per frame.
// ---------------------------------------------------------------------------------------
struct RawLine {
  float e[4];
  float length;
  bool valid;
};
__device__ __forceinline__ RawLine raw_line(float4 s, int ow, int oh, double minLength) {
  RawLine r;
  r.e[0] = s.x; r.e[1] = s.y; r.e[2] = s.z; r.e[3] = s.w;
  if (r.e[0] < 0) r.e[0] = 0;
  if (r.e[0] >= ow) r.e[0] = (float)ow - 1.0f;
  if (r.e[2] < 0) r.e[2] = 0;
  if (r.e[2] >= ow) r.e[2] = (float)ow - 1.0f;
  if (r.e[1] < 0) r.e[1] = 0;
  if (r.e[1] >= oh) r.e[1] = (float)oh - 1.0f;
  if (r.e[3] < 0) r.e[3] = 0;
  if (r.e[3] >= oh) r.e[3] = (float)oh - 1.0f;
  const float ddx = __fsub_rn(r.e[0], r.e[2]), ddy = __fsub_rn(r.e[1], r.e[3]);
  const double l = (double)(float)__dsqrt_rn(__dadd_rn(__dmul_rn((double)ddx, (double)ddx), __dmul_rn((double)ddy, (double)ddy)));
  r.length = (float)l;
  r.valid = l > minLength;
  return r;
}

// cv::LineIterator(img, pt1, pt2, 8).count (LSDDetector_custom.cpp:333-334).  An LSD endpoint in (w - 1.5, w)
// rounds to w, i.e. outside the image: OpenCV then clips the segment with cv::clipLine (integer
// Cohen-Sutherland on [0, w-1] x [0, h-1], int64 coordinates, double quotient truncated toward zero) before
// counting; a segment that misses the image counts 0.  Same steps as oracle_line.cpp:clip_line.
__device__ __forceinline__ long long clip_step(long long num, long long mul, long long den) {
  return (long long)__ddiv_rn(__dmul_rn((double)num, (double)mul), (double)den);
}
__device__ __forceinline__ int line_iterator_count(int w, int h, int ax, int ay, int bx, int by) {
  if ((unsigned)ax >= (unsigned)w || (unsigned)bx >= (unsigned)w || (unsigned)ay >= (unsigned)h || (unsigned)by >= (unsigned)h) {
    long long x1 = ax, y1 = ay, x2 = bx, y2 = by;
    const long long right = w - 1, bottom = h - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      long long a;
      if (c1 & 12) {
        a = c1 < 8 ? 0 : bottom;
        x1 += clip_step(a - y1, x2 - x1, y2 - y1);
        y1 = a;
        c1 = (x1 < 0) + (x1 > right) * 2;
      }
      if (c2 & 12) {
        a = c2 < 8 ? 0 : bottom;
        x2 += clip_step(a - y2, x2 - x1, y2 - y1);
        y2 = a;
        c2 = (x2 < 0) + (x2 > right) * 2;
      }
      if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        if (c1) {
          a = c1 == 1 ? 0 : right;
          y1 += clip_step(a - x1, y2 - y1, x2 - x1);
          x1 = a;
          c1 = 0;
        }
        if (c2) {
          a = c2 == 1 ? 0 : right;
          y2 += clip_step(a - x2, y2 - y1, x2 - x1);
          x2 = a;
          c2 = 0;
        }
      }
    }
    if ((c1 | c2) != 0) return 0;
    ax = (int)x1; ay = (int)y1; bx = (int)x2; by = (int)y2;
  }
  return max(abs(bx - ax), abs(by - ay)) + 1;
}

template <int NT>
__global__ void __launch_bounds__(NT) k_line_assemble(const __grid_constant__ LineGeom g, LineBufs b,
                                                      plvi_keyline* __restrict__ outKl, int* __restrict__ outCount) {
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  __shared__ int wsum[NT / 32], s_total, s_bad;
  float* resp = b.tmpResp + (size_t)f * g.segTotal;
  int* cls = b.tmpCls + (size_t)f * g.segTotal;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  // raw lines in detection order: octave 0 then octave 1; index space = [segOff, segOff + count)
  int total = 0;
  int cnt[2] = {0, 0};
  for (int o = 0; o < g.noct; o++) {
    cnt[o] = b.regCount[f * 2 + o];
    if (cnt[o] < 0) { if (tid == 0) s_bad = 1; cnt[o] = 0; }
    total += cnt[o];
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) outCount[f] = PLVI_ERR_CAPACITY;
    return;
  }
  // pass 1: validity + response; class ids = running index of valid lines
  const int chunk = (total + NT - 1) / NT;
  const int beg = min(tid * chunk, total), end = min(beg + chunk, total);
  auto locate = [&](int j, int& o, int& r) { o = (j < cnt[0]) ? 0 : 1; r = o ? j - cnt[0] : j; };
  int nv = 0;
  for (int j = beg; j < end; j++) {
    int o, r;
    locate(j, o, r);
    const LineOct& O = g.o[o];
    const RawLine rl = raw_line(b.segs[(size_t)f * g.segTotal + O.segOff + r], O.w, O.h, g.minLength);
    nv += rl.valid;
  }
  int incl = nv;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  if (tid == 0) {
    int a = 0;
    for (int w = 0; w < NT / 32; w++) { const int t = wsum[w]; wsum[w] = a; a += t; }
    s_total = a;
  }
  __syncthreads();
  int id = wsum[wid] + incl - nv;
  const int nvalid = s_total;
  for (int j = beg; j < end; j++) {
    int o, r;
    locate(j, o, r);
    const LineOct& O = g.o[o];
    const RawLine rl = raw_line(b.segs[(size_t)f * g.segTotal + O.segOff + r], O.w, O.h, g.minLength);
    if (rl.valid) {
      resp[id] = __fdiv_rn(rl.length, (float)max(O.w, O.h));
      cls[id] = j;  // raw index of the id-th valid line
      id++;
    }
  }
  __syncthreads();
  const bool select = (nvalid > g.nfeat) && g.nfeat != 0;
  const int nout = select ? g.nfeat : min(nvalid, g.keepCap);
  if (!select && nvalid > g.keepCap) {
    if (tid == 0) outCount[f] = PLVI_ERR_CAPACITY;
    return;
  }
  for (int v = tid; v < nvalid; v += NT) {
    int pos = v;
    if (select) {  // rank by (response desc, detection order asc)
      const float rv = resp[v];
      int rank = 0;
      for (int u = 0; u < nvalid; u++) {
        const float ru = resp[u];
        rank += (ru > rv) || (ru == rv && u < v);
      }
      pos = rank;
    }
    if (pos >= nout) continue;
    const int j = cls[v];
    int o, r;
    locate(j, o, r);
    const LineOct& O = g.o[o];
    const RawLine rl = raw_line(b.segs[(size_t)f * g.segTotal + O.segOff + r], O.w, O.h, g.minLength);
    const float octaveScale = o ? g.lineScale : 1.0f;   // pow(scale, octave), octave in {0,1}
    plvi_keyline k;
    k.startPointX = __fmul_rn(rl.e[0], octaveScale); k.startPointY = __fmul_rn(rl.e[1], octaveScale);
    k.endPointX = __fmul_rn(rl.e[2], octaveScale); k.endPointY = __fmul_rn(rl.e[3], octaveScale);
    k.sPointInOctaveX = rl.e[0]; k.sPointInOctaveY = rl.e[1];
    k.ePointInOctaveX = rl.e[2]; k.ePointInOctaveY = rl.e[3];
    k.lineLength = rl.length;
    const int ax = __float2int_rn(rl.e[0]), ay = __float2int_rn(rl.e[1]);
    const int bx = __float2int_rn(rl.e[2]), by = __float2int_rn(rl.e[3]);
    k.numOfPixels = line_iterator_count(O.w, O.h, ax, ay, bx, by);
    k.angle = glibc_atan2f(__fsub_rn(k.endPointY, k.startPointY), __fsub_rn(k.endPointX, k.startPointX));   // atan2(float, float)
    k.class_id = select ? pos : v;
    k.octave = o;
    k.size = __fmul_rn(__fsub_rn(k.endPointX, k.startPointX), __fsub_rn(k.endPointY, k.startPointY));
    k.response = resp[v];
    k.pt_x = __fdiv_rn(__fadd_rn(k.endPointX, k.startPointX), 2.f);
    k.pt_y = __fdiv_rn(__fadd_rn(k.endPointY, k.startPointY), 2.f);
    outKl[(size_t)f * g.keepCap + pos] = k;
  }
  if (tid == 0) outCount[f] = nout;
}

// ---------------------------------------------------------------------------------------
// LBD preprocessing
// ---------------------------------------------------------------------------------------
// Loads a (rows x 34 words) u8 tile whose first byte is global column gx0 (a multiple of 4) of
// rows gy0 .. gy0+rows-1, BORDER_REFLECT_101 outside the image.  sm pitch = 34 words.
__device__ __forceinline__ void load_tile34(const u8* __restrict__ s, int spitch, int w, int h, int gx0, int gy0,
                                            int rows, uint32_t* sm, int tid) {
  const bool aligned4 = ((spitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(s) & 3) == 0);
  for (int i = tid; i < rows * 34; i += 256) {
    const int r = i / 34, wq = i - r * 34;
    const int gy = reflect101_l(min(gy0 + r, h + 1), h);
    const int gx = gx0 + 4 * wq;
    const u8* row = s + (size_t)gy * spitch;
    uint32_t v;
    if (aligned4 && gx >= 0 && gx + 3 < w) {
      v = __ldg(reinterpret_cast<const uint32_t*>(row + gx));
    } else {
      v = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) v |= (uint32_t)__ldg(row + reflect101_l(min(gx + k, w + 1), w)) << (8 * k);
    }
    sm[i] = v;
  }
}

// k_gauss5: GaussianBlur(5x5, sigma 1) in OpenCV's 8.8 fixed point: kernel [14,62,104,62,14]/256 in
// both directions, no intermediate rounding, (v + 32768) >> 16.  All integer, so the passes commute:
// the vertical pass runs first on two 16-bit lanes per register (a column sum is <= 255 * 256), the
// horizontal pass combines the lane pairs with dp2a.  Tile = 128 x 32 output pixels.
__global__ void __launch_bounds__(256) k_gauss5(const u8* __restrict__ src, int spitch, size_t sfs, u8* __restrict__ dst,
                                                int dpitch, size_t dfs, int w, int h) {
  __shared__ __align__(16) uint32_t sin_[36 * 34];      // input rows y0-2 .. y0+33, bytes x0-4 .. x0+131
  __shared__ __align__(16) uint2 sv_[32 * 34];          // column sums: .x = even bytes of the word, .y = odd bytes
  const int tid = threadIdx.x, x0 = blockIdx.x * 128, y0 = blockIdx.y * 32;
  load_tile34(src + (size_t)blockIdx.z * sfs, spitch, w, h, x0 - 4, y0 - 2, 36, sin_, tid);
  __syncthreads();
  for (int i = tid; i < 32 * 34; i += 256) {
    const int r = i / 34, wq = i - r * 34;
    const uint32_t* q = sin_ + r * 34 + wq;
    const uint32_t w0 = q[0], w1 = q[34], w2 = q[68], w3 = q[102], w4 = q[136];
    const uint32_t M = 0x00ff00ffu;
    const uint32_t e = 14u * ((w0 & M) + (w4 & M)) + 62u * ((w1 & M) + (w3 & M)) + 104u * (w2 & M);
    const uint32_t o = 14u * (((w0 >> 8) & M) + ((w4 >> 8) & M)) + 62u * (((w1 >> 8) & M) + ((w3 >> 8) & M)) +
                       104u * ((w2 >> 8) & M);
    sv_[i] = make_uint2(e, o);
  }
  __syncthreads();
  u8* d = dst + (size_t)blockIdx.z * dfs;
  // weights as dp2a operand bytes (b0 multiplies the low 16-bit lane, b1 the high one)
  const uint32_t K_0_14 = 14u << 8, K_0_62 = 62u << 8, K_104_14 = 104u | (14u << 8), K_62_0 = 62u, K_62_62 = 62u | (62u << 8),
                 K_14_104 = 14u | (104u << 8), K_14_0 = 14u;
  for (int i = tid; i < 32 * 32; i += 256) {
    const int r = i >> 5, j = (i & 31) + 1;             // tile word j holds output columns x0 + 4(j-1) ..
    const int gy = y0 + r, gx = x0 + 4 * (j - 1);
    if (gy >= h || gx >= w) continue;
    const uint2 A = sv_[r * 34 + j - 1], B = sv_[r * 34 + j], Cn = sv_[r * 34 + j + 1];
    // V[4j-2] = A.x.hi, V[4j-1] = A.y.hi, V[4j] = B.x.lo, V[4j+1] = B.y.lo, V[4j+2] = B.x.hi, V[4j+3] = B.y.hi,
    // V[4j+4] = Cn.x.lo, V[4j+5] = Cn.y.lo
    const uint32_t o0 = __dp2a_lo(A.x, K_0_14, __dp2a_lo(A.y, K_0_62, __dp2a_lo(B.x, K_104_14, __dp2a_lo(B.y, K_62_0, 32768u))));
    const uint32_t o1 = __dp2a_lo(A.y, K_0_14, __dp2a_lo(B.x, K_62_62, __dp2a_lo(B.y, K_104_14, 32768u)));
    const uint32_t o2 = __dp2a_lo(B.x, K_14_104, __dp2a_lo(B.y, K_62_62, __dp2a_lo(Cn.x, K_14_0, 32768u)));
    const uint32_t o3 = __dp2a_lo(B.y, K_14_104, __dp2a_lo(B.x, K_0_62, __dp2a_lo(Cn.x, K_62_0, __dp2a_lo(Cn.y, K_14_0, 32768u))));
    const uint32_t out = (o0 >> 16) | ((o1 >> 16) << 8) | ((o2 >> 16) << 16) | ((o3 >> 16) << 24);
    u8* dp = d + (size_t)gy * dpitch + gx;
    if (gx + 3 < w) {
      *reinterpret_cast<uint32_t*>(dp) = out;
    } else {
      for (int k = 0; gx + k < w; k++) dp[k] = (u8)(out >> (8 * k));
    }
  }
}

// k_pyrdown: cv::pyrDown ([1,4,6,4,1]^2, (v + 128) >> 8, BORDER_REFLECT_101).  Tile = 64 x 16 output
// pixels; horizontal pass with dp4a into 16-bit sums (<= 4080), vertical pass on two lanes per
// register (<= 65408, no carry between the lanes).
__global__ void __launch_bounds__(256) k_pyrdown(const u8* __restrict__ src, int spitch, size_t sfs, int w, int h,
                                                 u8* __restrict__ dst, int dpitch, size_t dfs, int dw, int dh) {
  __shared__ __align__(16) uint32_t sin_[36 * 34];      // source rows 2*Y0-2 .. 2*Y0+33, bytes 2*X0-4 .. 2*X0+131
  __shared__ __align__(16) uint32_t sh_[36 * 32];       // horizontal sums, two 16-bit values per word
  const int tid = threadIdx.x, X0 = blockIdx.x * 64, Y0 = blockIdx.y * 16;
  load_tile34(src + (size_t)blockIdx.z * sfs, spitch, w, h, 2 * X0 - 4, 2 * Y0 - 2, 36, sin_, tid);
  __syncthreads();
  const uint32_t K = 1u | (4u << 8) | (6u << 16) | (4u << 24);
  for (int i = tid; i < 36 * 16; i += 256) {
    const int r = i >> 4, q = i & 15;                   // outputs X0 + 4q .. 4q+3 of source row r
    const uint32_t* p = sin_ + r * 34 + 2 * q;          // tile bytes 8q .. 8q+15; output j uses bytes 8q+2+2j .. +4
    const uint32_t w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3];
    const uint32_t h0 = __dp4a(__funnelshift_r(w0, w1, 16), K, (w1 >> 16) & 0xffu);
    const uint32_t h1 = __dp4a(w1, K, w2 & 0xffu);
    const uint32_t h2 = __dp4a(__funnelshift_r(w1, w2, 16), K, (w2 >> 16) & 0xffu);
    const uint32_t h3 = __dp4a(w2, K, w3 & 0xffu);
    *reinterpret_cast<uint2*>(sh_ + r * 32 + 2 * q) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
  }
  __syncthreads();
  u8* d = dst + (size_t)blockIdx.z * dfs;
  {
    const int Y = tid >> 4, q = tid & 15;               // 16 rows x 16 groups of 4 outputs
    const int gy = Y0 + Y, gx = X0 + 4 * q;
    if (gy < dh && gx < dw) {
      const uint32_t* p = sh_ + (2 * Y) * 32 + 2 * q;
      uint32_t acc[2];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const uint32_t v0 = p[k], v1 = p[32 + k], v2 = p[64 + k], v3 = p[96 + k], v4 = p[128 + k];
        acc[k] = ((v0 + v4 + 4u * (v1 + v3) + 6u * v2 + 0x00800080u) >> 8) & 0x00ff00ffu;
      }
      const uint32_t out = (acc[0] & 0xffu) | ((acc[0] >> 8) & 0xff00u) | ((acc[1] & 0xffu) << 16) | ((acc[1] >> 16) << 24);
      u8* dp = d + (size_t)gy * dpitch + gx;
      if (gx + 3 < dw) {
        *reinterpret_cast<uint32_t*>(dp) = out;
      } else {
        for (int k = 0; gx + k < dw; k++) dp[k] = (u8)(out >> (8 * k));
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_sobel(const u8* __restrict__ src, int spitch, size_t sfs, int w, int h,
                                               short2* __restrict__ dst, size_t dfs) {
  // thread = 4 horizontally adjacent pixels: one aligned word per row plus the two bytes next to
  // it.  dx = S[k+1] - S[k-1] with the column sums S = p0 + 2 p1 + p2, dy = D[k-1] + 2 D[k] + D[k+1]
  // with D = p2 - p0; the four in-word columns are handled as two 16-bit lanes per register.
  const int x4 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x4 >= w || y >= h) return;
  const u8* s = src + (size_t)blockIdx.z * sfs;
  const u8* rows[3] = {s + (size_t)reflect101_l(y - 1, h) * spitch, s + (size_t)y * spitch,
                       s + (size_t)reflect101_l(y + 1, h) * spitch};
  short2* out = dst + (size_t)blockIdx.z * dfs + (size_t)y * w + x4;
  const bool fast = x4 + 3 < w && ((spitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(s) & 3) == 0);
  if (fast) {
    const int xl = x4 > 0 ? x4 - 1 : 1, xr = x4 + 4 < w ? x4 + 4 : 2 * (w - 1) - (x4 + 4);
    uint32_t b[3];
    int l[3], r[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      b[i] = __ldg(reinterpret_cast<const uint32_t*>(rows[i] + x4));
      l[i] = __ldg(rows[i] + xl);
      r[i] = __ldg(rows[i] + xr);
    }
    const uint32_t M = 0x00ff00ffu;
    const uint32_t e0 = b[0] & M, e1 = b[1] & M, e2 = b[2] & M;
    const uint32_t o0 = (b[0] >> 8) & M, o1 = (b[1] >> 8) & M, o2 = (b[2] >> 8) & M;
    const uint32_t SE = e0 + 2u * e1 + e2, SO = o0 + 2u * o1 + o2;   // lanes: S[0] | S[2] << 16, S[1] | S[3] << 16
    const uint32_t DE = e2 + (M - e0), DO = o2 + (M - o0);           // lanes: D + 255
    const int S0 = SE & 0xffff, S2 = SE >> 16, S1 = SO & 0xffff, S3 = SO >> 16;
    const int D0 = (int)(DE & 0xffff) - 255, D2 = (int)(DE >> 16) - 255, D1 = (int)(DO & 0xffff) - 255, D3 = (int)(DO >> 16) - 255;
    const int SL = l[0] + 2 * l[1] + l[2], SR = r[0] + 2 * r[1] + r[2];
    const int DL = l[2] - l[0], DR = r[2] - r[0];
    const int dx0 = S1 - SL, dx1 = S2 - S0, dx2 = S3 - S1, dx3 = SR - S2;
    const int dy0 = DL + 2 * D0 + D1, dy1 = D0 + 2 * D1 + D2, dy2 = D1 + 2 * D2 + D3, dy3 = D2 + 2 * D3 + DR;
    uint4 v;
    v.x = (uint32_t)(dx0 & 0xffff) | ((uint32_t)dy0 << 16);
    v.y = (uint32_t)(dx1 & 0xffff) | ((uint32_t)dy1 << 16);
    v.z = (uint32_t)(dx2 & 0xffff) | ((uint32_t)dy2 << 16);
    v.w = (uint32_t)(dx3 & 0xffff) | ((uint32_t)dy3 << 16);
    if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      *reinterpret_cast<uint4*>(out) = v;
    } else {
      uint32_t* o32 = reinterpret_cast<uint32_t*>(out);
      o32[0] = v.x; o32[1] = v.y; o32[2] = v.z; o32[3] = v.w;
    }
    return;
  }
  int p[3][6];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int i = 0; i < 6; i++) p[r][i] = __ldg(rows[r] + reflect101_l(min(x4 - 1 + i, w + 1), w));
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (x4 + k >= w) break;
    const int dx = (p[0][k + 2] - p[0][k]) + 2 * (p[1][k + 2] - p[1][k]) + (p[2][k + 2] - p[2][k]);
    const int dy = (p[2][k] - p[0][k]) + 2 * (p[2][k + 1] - p[0][k + 1]) + (p[2][k + 2] - p[0][k + 2]);
    out[k] = make_short2((short)dx, (short)dy);
  }
}

// ---------------------------------------------------------------------------------------
// k_lbd: CTA of 64 threads per line; thread = one of the 63 rows of the support region
// (the reference accumulates sample coordinates and row sums serially along the line in
// float, so a row is one sequential chain); then 8 lanes fold rows into bands in row order.
// ---------------------------------------------------------------------------------------
__constant__ unsigned char c_comb[32][2] = {
    {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7},
    {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8}, {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

// k_lbd_rows: the 63 support-region rows of every line (thread = row; serial float chains kept).
__global__ void __launch_bounds__(64, 16) k_lbd_rows(const __grid_constant__ LineGeom g, LineBufs b,
                                                     const plvi_keyline* __restrict__ kls,
                                                     const int* __restrict__ counts) {
  const int li = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
  const int n = counts[f];
  if (li >= n || tid >= 63) return;
  const plvi_keyline kl = kls[(size_t)f * g.keepCap + li];
  const LineOct& O = g.o[kl.octave];
  const short2* grad = b.grad + (size_t)f * g.lbdTotal + O.lbdOff;
  const int realWidth = O.lw, imageWidth = O.lw - 1, imageHeight = O.lh - 1;
  const int L = kl.numOfPixels;
  const int halfWidth = (L - 1) / 2;
  const float midX = __fmul_rn(0.5f, __fadd_rn(kl.sPointInOctaveX, kl.ePointInOctaveX));
  const float midY = __fmul_rn(0.5f, __fadd_rn(kl.sPointInOctaveY, kl.ePointInOctaveY));
  float dL0, dL1;
  glibc_sincosf(kl.angle, dL1, dL0);   // cos(direction), sin(direction) on a float: host libm cosf / sinf
  const float dO0 = -dL1, dO1 = dL0;
  float sX = __fadd_rn(__fadd_rn(__fmul_rn(-dL0, (float)halfWidth), __fmul_rn(dL1, 31.f)), midX);
  float sY = __fadd_rn(__fsub_rn(__fmul_rn(-dL1, (float)halfWidth), __fmul_rn(dL0, 31.f)), midY);
  for (int r = 0; r < tid; r++) {
    sX = __fsub_rn(sX, dL1);
    sY = __fadd_rn(sY, dL0);
  }
  float pL = 0.f, nL = 0.f, pO = 0.f, nO = 0.f;
  for (int wID = 0; wID < L; wID++) {
    int tc = (int)(short)(int)roundf(sX);
    const int xCor = tc < 0 ? 0 : (tc > imageWidth ? imageWidth : tc);
    tc = (int)(short)(int)roundf(sY);
    const int yCor = tc < 0 ? 0 : (tc > imageHeight ? imageHeight : tc);
    const short2 gv = __ldg(grad + yCor * realWidth + xCor);
    const float gDL = __fadd_rn(__fmul_rn((float)gv.x, dL0), __fmul_rn((float)gv.y, dL1));
    const float gDO = __fadd_rn(__fmul_rn((float)gv.x, dO0), __fmul_rn((float)gv.y, dO1));
    if (gDL > 0) pL = __fadd_rn(pL, gDL); else nL = __fsub_rn(nL, gDL);
    if (gDO > 0) pO = __fadd_rn(pO, gDO); else nO = __fsub_rn(nO, gDO);
    sX = __fadd_rn(sX, dL0);
    sY = __fadd_rn(sY, dL1);
  }
  const float c = (float)b.lbdG[tid];
  pL = __fmul_rn(c, pL); nL = __fmul_rn(c, nL);
  pO = __fmul_rn(c, pO); nO = __fmul_rn(c, nO);
  float* rows = b.lbdRows + ((size_t)f * g.keepCap + li) * 512;   // [8][64]
  rows[0 * 64 + tid] = pL; rows[1 * 64 + tid] = nL; rows[2 * 64 + tid] = __fmul_rn(pL, pL); rows[3 * 64 + tid] = __fmul_rn(nL, nL);
  rows[4 * 64 + tid] = pO; rows[5 * 64 + tid] = nO; rows[6 * 64 + tid] = __fmul_rn(pO, pO); rows[7 * 64 + tid] = __fmul_rn(nO, nO);
}

// k_lbd_fold: eight lanes per line (four lines per warp, 16 per block).  Each lane folds the 63 rows of one statistic
// into the 9 bands in row order, the group's first lane forms / normalises the 72 floats (a serial chain in the
// reference's order: with a whole warp per line 31 lanes idled through it), the eight lanes pack 4 bytes each.
#define LBD_FOLD_LPB 16
__global__ void __launch_bounds__(128) k_lbd_fold(const __grid_constant__ LineGeom g, LineBufs b,
                                                  const plvi_keyline* __restrict__ kls,
                                                  const int* __restrict__ counts, uint8_t* __restrict__ desc,
                                                  double* __restrict__ lineEq) {
  const int lane = threadIdx.x & 7, wid = threadIdx.x >> 3;
  const int li = blockIdx.x * LBD_FOLD_LPB + wid, f = blockIdx.y;
  if (li >= counts[f]) return;
  __shared__ float sband[LBD_FOLD_LPB][8][9];
  __shared__ float sdes[LBD_FOLD_LPB][72];
  const float* rows = b.lbdRows + ((size_t)f * g.keepCap + li) * 512;
  float* des = sdes[wid];
  {
    const int q = lane;
    const bool sq = (q == 2 || q == 3 || q == 6 || q == 7);
    float band[9];
#pragma unroll
    for (int k = 0; k < 9; k++) band[k] = 0.f;
#pragma unroll
    for (int bandID = 0; bandID < 9; bandID++) {
#pragma unroll
      for (int m = 0; m < 7; m++) {
        const float rs = rows[q * 64 + bandID * 7 + m];
        float cf = (float)b.lbdL[m + 7];
        band[bandID] = __fadd_rn(band[bandID], sq ? __fmul_rn(__fmul_rn(cf, cf), rs) : __fmul_rn(cf, rs));
        if (bandID - 1 >= 0) {
          cf = (float)b.lbdL[m + 14];
          band[bandID - 1] = __fadd_rn(band[bandID - 1], sq ? __fmul_rn(__fmul_rn(cf, cf), rs) : __fmul_rn(cf, rs));
        }
        if (bandID + 1 < 9) {
          cf = (float)b.lbdL[m];
          band[bandID + 1] = __fadd_rn(band[bandID + 1], sq ? __fmul_rn(__fmul_rn(cf, cf), rs) : __fmul_rn(cf, rs));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) sband[wid][q][k] = band[k];
  }
  __syncwarp();
  if (lane == 0) {
    const plvi_keyline kl = kls[(size_t)f * g.keepCap + li];
    const float invN2 = (float)(1.0 / 14.0), invN3 = (float)(1.0 / 21.0);
    for (int k = 0; k < 9; k++) {
      const float invN = (k == 0 || k == 8) ? invN2 : invN3;
      const int d = k * 8;
      float t;
      t = __fmul_rn(sband[wid][0][k], invN); des[d] = t;
      des[d + 4] = __fsqrt_rn(__fsub_rn(__fmul_rn(sband[wid][2][k], invN), __fmul_rn(t, t)));
      t = __fmul_rn(sband[wid][1][k], invN); des[d + 1] = t;
      des[d + 5] = __fsqrt_rn(__fsub_rn(__fmul_rn(sband[wid][3][k], invN), __fmul_rn(t, t)));
      t = __fmul_rn(sband[wid][4][k], invN); des[d + 2] = t;
      des[d + 6] = __fsqrt_rn(__fsub_rn(__fmul_rn(sband[wid][6][k], invN), __fmul_rn(t, t)));
      t = __fmul_rn(sband[wid][5][k], invN); des[d + 3] = t;
      des[d + 7] = __fsqrt_rn(__fsub_rn(__fmul_rn(sband[wid][7][k], invN), __fmul_rn(t, t)));
    }
    float tM = 0.f, tS = 0.f;
    for (int k = 0; k < 9; k++) {
      for (int j = 0; j < 4; j++) tM = __fadd_rn(tM, __fmul_rn(des[8 * k + j], des[8 * k + j]));
      for (int j = 4; j < 8; j++) tS = __fadd_rn(tS, __fmul_rn(des[8 * k + j], des[8 * k + j]));
    }
    tM = __fdiv_rn(1.f, __fsqrt_rn(tM));
    tS = __fdiv_rn(1.f, __fsqrt_rn(tS));
    for (int k = 0; k < 9; k++) {
      for (int j = 0; j < 4; j++) des[8 * k + j] = __fmul_rn(des[8 * k + j], tM);
      for (int j = 4; j < 8; j++) des[8 * k + j] = __fmul_rn(des[8 * k + j], tS);
    }
    for (int i = 0; i < 72; i++)
      if ((double)des[i] > 0.4) des[i] = 0.4f;
    float t = 0.f;
    for (int i = 0; i < 72; i++) t = __fadd_rn(t, __fmul_rn(des[i], des[i]));
    t = __fdiv_rn(1.f, __fsqrt_rn(t));
    for (int i = 0; i < 72; i++) des[i] = __fmul_rn(des[i], t);
    // line equation (f64), src/LineExtractor.cc:106-116
    const double sx = kl.startPointX, sy = kl.startPointY, ex = kl.endPointX, ey = kl.endPointY;
    const double l0 = __dsub_rn(sy, ey), l1 = __dsub_rn(ex, sx), l2 = __dsub_rn(__dmul_rn(sx, ey), __dmul_rn(sy, ex));
    const double nrm = __dsqrt_rn(__dadd_rn(__dmul_rn(l0, l0), __dmul_rn(l1, l1)));
    double* eq = lineEq + ((size_t)f * g.keepCap + li) * 3;
    eq[0] = __ddiv_rn(l0, nrm); eq[1] = __ddiv_rn(l1, nrm); eq[2] = __ddiv_rn(l2, nrm);
  }
  __syncwarp();
  {
    unsigned word = 0u;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int byte = lane * 4 + j;
      const float* f1 = des + 8 * c_comb[byte][0];
      const float* f2 = des + 8 * c_comb[byte][1];
      unsigned r = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) r |= (unsigned)(f1[i] > f2[i]) << i;
      word |= r << (8 * j);
    }
    uint8_t* out = desc + ((size_t)f * g.keepCap + li) * 32 + lane * 4;
    if ((reinterpret_cast<uintptr_t>(out) & 3u) == 0) *reinterpret_cast<unsigned*>(out) = word;
    else { out[0] = (uint8_t)word; out[1] = (uint8_t)(word >> 8); out[2] = (uint8_t)(word >> 16); out[3] = (uint8_t)(word >> 24); }
  }
}

// ---------------------------------------------------------------------------------------
// launch sequence
// ---------------------------------------------------------------------------------------
// shared memory of k_lsd_pre for octave O: row-filtered window + column-filtered window (f64) + u8 window
size_t lsd_pre_smem(const LineGeom& g, const LineOct& O) {
  if (g.hk == 0) return (size_t)(g.preTH + 1) * (PRE_TW + 2) * sizeof(double);
  const size_t rows = (size_t)O.preRH + 2 * g.hk;
  return ((rows + 1) * O.preSW + (size_t)O.preRH * O.preSW) * sizeof(double) + rows * ((O.preSW + 2 * g.hk + 11) & ~3);
}

int launch_line_pipeline(const LineGeom& g, const LinePtrs& p, const LineBufs& b, int n, plvi_keyline* dKl,
                         uint8_t* dDesc, double* dEq, int* dCounts, cudaStream_t st, LineAux aux, int* launches,
                         StageProf* prof) {
  int nl = 0;
  StageProf nop;
  if (!prof) prof = &nop;
  prof->begin(st);
  // LSD pyramid level 1 (2x bilinear; LSDDetectorC::ComputePyramid)
  if (g.noct > 1) {
    const LineOct& d = g.o[1];
    launch_resize_u8(p.img[0], p.ipitch[0], p.ifs[0], g.o[0].w, g.o[0].h, const_cast<u8*>(p.img[1]), p.ipitch[1],
                     p.ifs[1], d.w, d.h, b.rsTab, b.rsTab + d.w, n, st);
    nl++;
    prof->mark("k_resize", st);
  }
  for (int o = 0; o < g.noct; o++) {
    const LineOct& O = g.o[o];
    const dim3 pgrid(O.wpr, (O.sh + g.preTH - 1) / g.preTH, n);
    const size_t psm = lsd_pre_smem(g, O);
    switch (g.hk) {
      case 0: k_lsd_pre<0><<<pgrid, 256, psm, st>>>(g, o, p.img[o], p.ipitch[o], p.ifs[o], b.tabs, b); break;
      case 3: k_lsd_pre<3><<<pgrid, 256, psm, st>>>(g, o, p.img[o], p.ipitch[o], p.ifs[o], b.tabs, b); break;
      case 4: k_lsd_pre<4><<<pgrid, 256, psm, st>>>(g, o, p.img[o], p.ipitch[o], p.ifs[o], b.tabs, b); break;
      case 5: k_lsd_pre<5><<<pgrid, 256, psm, st>>>(g, o, p.img[o], p.ipitch[o], p.ifs[o], b.tabs, b); break;
      case 6: k_lsd_pre<6><<<pgrid, 256, psm, st>>>(g, o, p.img[o], p.ipitch[o], p.ifs[o], b.tabs, b); break;
      case 7: k_lsd_pre<7><<<pgrid, 256, psm, st>>>(g, o, p.img[o], p.ipitch[o], p.ifs[o], b.tabs, b); break;
      default: k_lsd_pre<8><<<pgrid, 256, psm, st>>>(g, o, p.img[o], p.ipitch[o], p.ifs[o], b.tabs, b); break;
    }
    prof->mark("k_lsd_pre", st);
    nl += 1;
  }
  // From here on the main stream holds the latency-bound region growing (low issue-slot use): a caller may hold
  // other issue-bound work (the ORB pipeline) back until this point (plvi_line_stage_event).
  // plvi_line_stage_counter: k_lsd_spec counts its blocks as they start; the other schedules let a waiter pass at once
  {
    const bool specPath = g.refine == 0 && !(b.brMax > 0 && n <= b.brUse) && b.useSpec;
    PLVI_CUDA_TRY(cudaMemsetAsync(b.specStart, specPath ? 0 : 0x3f, sizeof(int), st));
  }
  // (on the speculation path the event is recorded behind k_lsd_spec_init: released right after k_lsd_pre, the ORB
  // pyramid flooded the SMs and the two short set-up kernels of the speculation took 4.2 ms instead of 0.3 -- on the
  // critical path of the step)
  auto record_stage = [&]() -> int {
    if (aux.stage && (!prof->on || StageProf::timeline())) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(st, &cs);
      PLVI_CUDA_TRY(cudaEventRecordWithFlags(aux.stage, st, cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault));
    }
    return PLVI_OK;
  };
  const bool stageLate = g.refine == 0 && !(b.brMax > 0 && n <= b.brUse) && b.useSpec;
  if (!stageLate) { const int rc = record_stage(); if (rc != PLVI_OK) return rc; }
  // The LBD pyramid + Sobel only depend on the input frame: they run on the auxiliary stream while
  // the latency-bound region growing occupies the main one (serially when profiling, for clean times).
  const bool fork = aux.stream != nullptr && (!prof->on || StageProf::timeline());
  cudaStream_t ls = fork ? aux.stream : st;
  if (fork) {
    PLVI_CUDA_TRY(cudaEventRecord(aux.fork, st));
    PLVI_CUDA_TRY(cudaStreamWaitEvent(aux.stream, aux.fork, 0));
  }
  {
    const LineOct& O0 = g.o[0];
    k_gauss5<<<dim3((O0.lw + 127) / 128, (O0.lh + 31) / 32, n), 256, 0, ls>>>(p.img[0], p.ipitch[0], p.ifs[0], b.lbdImg0,
                                                                             O0.lpitch, (size_t)O0.lpitch * O0.lh, O0.lw, O0.lh);
    prof->mark("k_gauss5", st);
    k_sobel<<<dim3((O0.lw + 255) / 256, (O0.lh + 3) / 4, n), 256, 0, ls>>>(b.lbdImg0, O0.lpitch, (size_t)O0.lpitch * O0.lh, O0.lw,
                                                                        O0.lh, b.grad + O0.lbdOff, g.lbdTotal);
    prof->mark("k_sobel", st);
    nl += 2;
    if (g.noct > 1) {
      const LineOct& O1 = g.o[1];
      k_pyrdown<<<dim3((O1.lw + 63) / 64, (O1.lh + 15) / 16, n), 256, 0, ls>>>(b.lbdImg0, O0.lpitch, (size_t)O0.lpitch * O0.lh,
                                                                            O0.lw, O0.lh, b.lbdImg1, O1.lpitch,
                                                                            (size_t)O1.lpitch * O1.lh, O1.lw, O1.lh);
      prof->mark("k_pyrdown", st);
      k_sobel<<<dim3((O1.lw + 255) / 256, (O1.lh + 3) / 4, n), 256, 0, ls>>>(b.lbdImg1, O1.lpitch, (size_t)O1.lpitch * O1.lh,
                                                                          O1.lw, O1.lh, b.grad + O1.lbdOff, g.lbdTotal);
      prof->mark("k_sobel", st);
      nl += 2;
    }
  }
  if (fork) PLVI_CUDA_TRY(cudaEventRecord(aux.join, aux.stream));
  const size_t growSmem = ((size_t)g.o[0].wpr * GROW_K + GROW_RQ) * sizeof(unsigned);
  const bool bandRun = b.brMax > 0 && n <= b.brUse && g.refine == 0;
  if (g.refine > 0) {
    // lsd_refine 1 / 2: refine() couples consecutive seeds through released pixels -- one serial warp per (frame, octave)
    k_lsd_grow_refine<<<dim3(g.noct, n), 32, growSmem, st>>>(g, b, g.refine, 1.0, 0.6);   // log_eps, density_th: src/LineExtractor.cc:62-63
    prof->mark("k_lsd_grow_refine", st);
    nl += 1;
  } else if (bandRun) {
    // small batch: many bands per frame, rounds of (compose inputs, run the bands whose input changed)
    PLVI_CUDA_TRY(cudaMemsetAsync(b.brFlags, 0, (size_t)n * 2 * BR_FLAGS * sizeof(int), st));
    const int maxWords = g.o[0].sh * g.o[0].wpr;
    const bool bigWindow = (long long)n * g.brBandsPerFrame <= 148 * 5;
    const size_t runSmem = ((size_t)g.o[0].wpr * (bigWindow ? BR_K_BIG : BR_K_SMALL) * 2 + GROW_RQ) * sizeof(unsigned);
    const dim3 cgrid((maxWords + 255) / 256, g.noct, n);
    for (int r = 1; r <= b.brRounds; r++) {
      k_lsd_band_compose<<<cgrid, 256, 0, st>>>(g, b, r, 0);
      if (bigWindow) k_lsd_band_run<BR_K_BIG><<<dim3(g.brBandsPerFrame, n), 32, runSmem, st>>>(g, b);
      else k_lsd_band_run<BR_K_SMALL><<<dim3(g.brBandsPerFrame, n), 32, runSmem, st>>>(g, b);
      nl += 2;
    }
    prof->mark("k_lsd_band_rounds", st);
    k_lsd_band_compose<<<cgrid, 256, 0, st>>>(g, b, b.brRounds + 1, 1);
    k_lsd_band_gather<<<dim3(g.noct, n), 256, 0, st>>>(g, b);
    k_lsd_grow<<<dim3(g.noct, n), 32, growSmem, st>>>(g, b, 1);
    prof->mark("k_lsd_band_finish", st);
    nl += 3;
  } else if (b.useSpec) {
    // equal-load bands shorten the longest chain while the batch is latency-bound (measured on B200: 512 frames +7 %,
    // 1024 +2 %); from ~4096 frames on the kernel is bound by its memory instructions and equal rows are 1.5 % ahead
    LineBufs bs = b;
    bs.eqLoad = (b.eqLoad == 1 ? (n <= 2048) : (b.eqLoad == 2)) && g.o[0].sh <= 1024;
    k_lsd_band_split<<<dim3(g.noct, n), 256, 0, st>>>(g, bs);
    k_lsd_spec_init<<<dim3(g.tasksPerFrame, n), 256, 0, st>>>(g, b);
    nl += 1;
    prof->mark("k_lsd_spec_init", st);
    { const int rc = record_stage(); if (rc != PLVI_OK) return rc; }
    {
      const dim3 sgrid(g.tasksPerFrame, (n + 32 * GROW_WPB - 1) / (32 * GROW_WPB));
      // large batches: the 64-register build (room for the ORB kernels beside it); smaller ones are bound by the length
      // of the chain and take the natural allocation (512 frames: 29 vs 35 ms)
      if (bs.eqLoad) k_lsd_spec<true, SPEC_MINB_SMALL><<<sgrid, 32 * GROW_WPB, 0, st>>>(g, b, n);
      else if (n > 2048) k_lsd_spec<false, SPEC_MINB><<<sgrid, 32 * GROW_WPB, 0, st>>>(g, b, n);
      else k_lsd_spec<false, SPEC_MINB_SMALL><<<sgrid, 32 * GROW_WPB, 0, st>>>(g, b, n);
    }
    prof->mark("k_lsd_spec", st);
    const size_t commitSmem = growSmem + ((size_t)g.o[0].wpr + 1) * GROW_K * sizeof(unsigned);   // + phantom window and its row masks
    k_lsd_commit<<<dim3(g.noct * ((n + COMMIT_WPB - 1) / COMMIT_WPB)), 32 * COMMIT_WPB, commitSmem * COMMIT_WPB, st>>>(
        g, b, n, (int)(commitSmem / sizeof(unsigned)));
    prof->mark("k_lsd_commit", st);
    nl += 2;
  } else {
    k_lsd_grow<<<dim3(g.noct, n), 32, growSmem, st>>>(g, b, 0);
    prof->mark("k_lsd_grow", st);
  }
  if (g.refine == 0) {
    k_lsd_rect<RECT_G><<<dim3(g.noct, n, bandRun ? 8 : 2), 256, 0, st>>>(g, b, bandRun ? 1 : 0);
    prof->mark("k_lsd_rect", st);
  }
  k_line_assemble<512><<<n, 512, 0, st>>>(g, b, dKl, dCounts);
  prof->mark("k_line_assemble", st);
  nl += 3;
  if (fork) PLVI_CUDA_TRY(cudaStreamWaitEvent(st, aux.join, 0));
  k_lbd_rows<<<dim3(g.keepCap, n), 64, 0, st>>>(g, b, dKl, dCounts);
  prof->mark("k_lbd_rows", st);
  k_lbd_fold<<<dim3((g.keepCap + LBD_FOLD_LPB - 1) / LBD_FOLD_LPB, n), 128, 0, st>>>(g, b, dKl, dCounts, dDesc, dEq);
  prof->mark("k_lbd_fold", st);
  nl += 2;
  PLVI_CUDA_TRY(cudaGetLastError());
  if (launches) *launches = nl;
  return PLVI_OK;
}

int line_kernel_attrs(const LineGeom& g) {
  const size_t growSmem = ((size_t)g.o[0].wpr * GROW_K + GROW_RQ) * sizeof(unsigned);
  const size_t commitSmem = (growSmem + ((size_t)g.o[0].wpr + 1) * GROW_K * sizeof(unsigned)) * COMMIT_WPB;
  if (commitSmem > 200 * 1024) { set_error("image too large for the LSD shared-memory bitmap"); return PLVI_ERR_CAPACITY; }
  {
    size_t psm = 0;
    for (int o = 0; o < g.noct; o++) psm = std::max(psm, lsd_pre_smem(g, g.o[o]));
    if (psm > 200 * 1024) { set_error("lsd_scale too small for the shared-memory tile of k_lsd_pre"); return PLVI_ERR_CAPACITY; }
    if (psm > 40 * 1024) {
      const int v = (int)psm;
      PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_pre<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
      PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_pre<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
      PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_pre<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
      PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_pre<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
      PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_pre<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
      PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_pre<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
      PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_pre<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
    }
  }
  // the growth kernels want as many resident warps as registers allow: give shared memory the large carve-out
  PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_commit, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  {
    const size_t runSmem = ((size_t)g.o[0].wpr * BR_K_BIG * 2 + GROW_RQ) * sizeof(unsigned);
    if (runSmem > 220 * 1024) { set_error("image too wide for the band-run shared-memory window"); return PLVI_ERR_CAPACITY; }
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_band_run<BR_K_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)runSmem));
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_band_run<BR_K_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)commitSmem));
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_band_run<BR_K_SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  if (commitSmem > 40 * 1024) {
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_grow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)commitSmem));
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_grow_refine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)commitSmem));
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_lsd_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)commitSmem));

  }
  return PLVI_OK;
}

}  // namespace plvi
