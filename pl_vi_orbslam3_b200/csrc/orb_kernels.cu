// ORB extraction kernels for sm_100a.
//
//   k_resize     ComputePyramid            src/ORBextractor.cc:1152-1177 (cv::resize INTER_LINEAR, u8)
//   k_fast       grid FAST + per-cell th   src/ORBextractor.cc:763-855   (cv::FAST 9/16 + 3x3 NMS)
//   k_octree     DistributeOctTree         src/ORBextractor.cc:537-761
//   k_blur7      GaussianBlur 7x7 s=2      src/ORBextractor.cc:1114-1115
//   k_layout     output row assignment     src/ORBextractor.cc:1104-1144 (lapping split)
//   k_orient_desc IC_Angle + rBRIEF        src/ORBextractor.cc:75-145
//
// All stages are integer/byte streaming or gather work (no dense contraction): the
// design rules are coalesced loads, shared-memory tiles and enough CTAs per launch to
// fill 148 SMs -- every launch covers all frames of the batch (and all levels where the
// stage allows), so grids are thousands of CTAs.
#include "orb_pattern_table.h"
#include "plvi_internal.cuh"
#include "line_internal.cuh"

namespace plvi {

__device__ const signed char d_pattern[1024] = PLVI_ORB_PATTERN_VALUES;
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// ---------------------------------------------------------------------------------
// block-wide helpers
// ---------------------------------------------------------------------------------
// In-place exclusive scan of arr[0..n) (shared memory); returns the total.  Must be
// called by all NT threads; wtmp needs 33 ints.
template <int NT>
__device__ int block_excl_scan(int* arr, int n, int* wtmp) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int chunk = (n + NT - 1) / NT;
  const int beg = min(tid * chunk, n), end = min(beg + chunk, n);
  int sum = 0;
  for (int i = beg; i < end; i++) sum += arr[i];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wtmp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int v = lane < NT / 32 ? wtmp[lane] : 0;
    int iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += t;
    }
    wtmp[lane] = iv - v;
    if (lane == 31) wtmp[32] = iv;
  }
  __syncthreads();
  int base = wtmp[wid] + incl - sum;
  const int total = wtmp[32];
  for (int i = beg; i < end; i++) {
    int t = arr[i];
    arr[i] = base;
    base += t;
  }
  __syncthreads();
  return total;
}

// ---------------------------------------------------------------------------------
// k_resize: one pyramid level from the previous one.  Fixed-point bilinear exactly as
// OpenCV's 8-bit INTER_LINEAR path: coefficient rows (ofs, a0 | a1 << 16) are computed
// on the host in float32; H = S[s]*a0 + S[s+1]*a1;
// dst = (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2.
// Thread = 4 horizontally adjacent dst pixels (one aligned 32-bit store).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize(const u8* __restrict__ src, int spitch, size_t sfs,
                                                int sw, int sh, u8* __restrict__ dst, int dpitch,
                                                size_t dfs, int dw, int dh,
                                                const int2* __restrict__ xtab,
                                                const int2* __restrict__ ytab) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x4 >= dw || y >= dh) return;
  const int2 yt = __ldg(&ytab[y]);
  const int b0 = (short)(yt.y & 0xffff), b1 = yt.y >> 16;
  const u8* r0 = src + (size_t)blockIdx.z * sfs + (size_t)yt.x * spitch;
  const u8* r1 = src + (size_t)blockIdx.z * sfs + (size_t)min(yt.x + 1, sh - 1) * spitch;
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int x = x4 + i;
    if (x < dw) {
      const int2 xt = __ldg(&xtab[x]);
      const int a0 = (short)(xt.y & 0xffff), a1 = xt.y >> 16;
      const int s0 = xt.x, s1 = min(xt.x + 1, sw - 1);
      const int h0 = __ldg(r0 + s0) * a0 + __ldg(r0 + s1) * a1;
      const int h1 = __ldg(r1 + s0) * a0 + __ldg(r1 + s1) * a1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      out |= (uint32_t)v << (8 * i);
    }
  }
  *reinterpret_cast<uint32_t*>(dst + (size_t)blockIdx.z * dfs + (size_t)y * dpitch + x4) = out;
}

// k_resize_tma: the same arithmetic on a source window staged in shared memory by the TMA engine.  One CTA
// produces a RS_TW x RS_TH tile of the destination level; its source rows (<= RS_SRC_H rows of <= RS_SRC_W bytes,
// 16-byte aligned) are fetched with one bulk asynchronous copy per row (cp.async.bulk, completion counted on an
// mbarrier), so no thread issues byte-granular global loads and the two dependent global round trips per pixel of
// k_resize (table entry -> source bytes) become one bulk wait plus shared-memory reads.
#define RS_TW 256
#define RS_TH 8
#define RS_SRC_W 576
#define RS_SRC_H 20
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) k_resize_tma(const u8* __restrict__ src, int spitch, size_t sfs, int sw, int sh,
                                                    u8* __restrict__ dst, int dpitch, size_t dfs, int dw, int dh,
                                                    const int2* __restrict__ xtab, const int2* __restrict__ ytab) {
  __shared__ __align__(128) u8 tile[RS_SRC_H * RS_SRC_W];
  __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * RS_TW, y0 = blockIdx.y * RS_TH;
  const int ys0 = __ldg(&ytab[y0]).x;
  const int ys1 = min(__ldg(&ytab[min(y0 + RS_TH - 1, dh - 1)]).x + 1, sh - 1);
  const int xs0 = __ldg(&xtab[x0]).x & ~15;
  const int xe = min(__ldg(&xtab[min(x0 + RS_TW - 1, dw - 1)]).x + 1, sw - 1);
  const int nrows = ys1 - ys0 + 1;
  const int nbytes = (xe + 1 - xs0 + 15) & ~15;
  const uint32_t barAddr = smem_addr_u32(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barAddr), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(nrows * nbytes) : "memory");
    const u8* g = src + (size_t)blockIdx.z * sfs + (size_t)ys0 * spitch + xs0;
    for (int r = 0; r < nrows; r++)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr_u32(tile + r * RS_SRC_W)),
                   "l"(g + (size_t)r * spitch), "r"(nbytes), "r"(barAddr)
                   : "memory");
  }
  // everybody waits for the transaction count of phase 0
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "RS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra RS_DONE;\n"
      "bra RS_WAIT;\n"
      "RS_DONE:\n"
      "}" ::"r"(barAddr),
      "r"(0)
      : "memory");
  const int x4 = x0 + (tid & 63) * 4;
  if (x4 >= dw) return;
  int s0[4], s1[4], a0[4], a1[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int2 xt = __ldg(&xtab[min(x4 + i, dw - 1)]);
    a0[i] = (short)(xt.y & 0xffff); a1[i] = xt.y >> 16;
    s0[i] = xt.x - xs0; s1[i] = min(xt.x + 1, sw - 1) - xs0;
  }
#pragma unroll
  for (int k = 0; k < RS_TH / 4; k++) {
    const int y = y0 + (tid >> 6) + 4 * k;
    if (y >= dh) break;
    const int2 yt = __ldg(&ytab[y]);
    const int b0 = (short)(yt.y & 0xffff), b1 = yt.y >> 16;
    const u8* r0 = tile + (yt.x - ys0) * RS_SRC_W;
    const u8* r1 = tile + (min(yt.x + 1, sh - 1) - ys0) * RS_SRC_W;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int h0 = r0[s0[i]] * a0[i] + r0[s1[i]] * a1[i];
      const int h1 = r1[s0[i]] * a0[i] + r1[s1[i]] * a1[i];
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      if (x4 + i < dw) out |= (uint32_t)v << (8 * i);
    }
    *reinterpret_cast<uint32_t*>(dst + (size_t)blockIdx.z * dfs + (size_t)y * dpitch + x4) = out;
  }
}

// ---------------------------------------------------------------------------------
// k_fast.  One CTA = up to 4 horizontally adjacent FAST cells of one cell row of one
// level of one frame.  The reference runs cv::FAST on each cell's ROI
// [ini, ini+cell+6) separately: scores exist only >= 3 px inside the ROI, so the
// detection strips of the cells tile the level without overlap and the 3x3 NMS never
// sees a score from a neighbouring cell.  A cell keeps its th=iniTh survivors, or, if
// it has none, its th=minTh survivors: two phases over the tile, the second one confined to
// the cells the first one left empty (the score is threshold independent).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int fast_arc_score(const int (&d)[16]) {
  // max over the 16 arcs of 9 contiguous ring pixels of min(d) and of min(-d)
  int bp = -256, bn = -256;
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    int mn = min(d[(k + 1) & 15], d[(k + 2) & 15]);
    int mx = max(d[(k + 1) & 15], d[(k + 2) & 15]);
#pragma unroll
    for (int j = 3; j <= 8; j++) {
      mn = min(mn, d[(k + j) & 15]);
      mx = max(mx, d[(k + j) & 15]);
    }
    bp = max(bp, max(min(mn, d[k]), min(mn, d[(k + 9) & 15])));
    bn = max(bn, max(-max(mx, d[k]), -max(mx, d[(k + 9) & 15])));
  }
  return max(bp, bn) - 1;
}

// Shared-memory tile layout: ROI column c is stored at byte column c + 1, so that the first
// detection column (c = 3) is word aligned and groups of 4 detection pixels are one 32-bit word.
#define FAST_PAD 1

__global__ void __launch_bounds__(256, 5) k_fast(const __grid_constant__ OrbGeom g,
                                              const __grid_constant__ OrbPtrs p,
                                              const FastTile* __restrict__ tiles,
                                              uint32_t* __restrict__ cand,
                                              int* __restrict__ candCount, int tilePitch,
                                              int tileRows, int listCap, int survCap) {
  extern __shared__ __align__(16) u8 smem[];
  const FastTile t = tiles[blockIdx.x];
  const int f = blockIdx.y;
  const OrbLevel& L = g.lv[t.level];
  const int maxBX = L.w - kEdge, maxBY = L.h - kEdge;
  const int iniY = kEdge + t.cellRow * L.hCell;
  const int iniX = kEdge + t.cellCol0 * L.wCell;
  const int maxY = min(iniY + L.hCell + 6, maxBY);
  const int maxX = min(iniX + t.ncells * L.wCell + 6, maxBX);
  const int rw = maxX - iniX, rh = maxY - iniY;
  if (rw < 7 || rh < 7) return;

  u8* simg = smem;                                   // [tileRows][tilePitch]
  u8* ssc = smem + tileRows * tilePitch;             // scores, same shape
  unsigned short* list1 = reinterpret_cast<unsigned short*>(ssc + tileRows * tilePitch);  // [listCap] dy << 8 | dx
  unsigned short* list2 = list1 + listCap;                                               // [listCap] corners
  uint32_t* kept = reinterpret_cast<uint32_t*>(list2 + listCap);                         // [survCap] NMS survivors of the cells' deciding phase
  __shared__ int s_n1, s_n2, s_nkept, s_base;
  __shared__ int s_cellFlag[8];
  const int tid = threadIdx.x;
  if (tid == 0) { s_n1 = 0; s_n2 = 0; s_nkept = 0; }
  if (tid < 8) s_cellFlag[tid] = 0;

  // ---- tile load: smem word k of a row holds ROI bytes 4k-1 .. 4k+2
  const int ipitch = p.ipitch[t.level];
  const u8* gbase = p.img[t.level] + (size_t)f * p.ifs[t.level];
  const int nwords = (rw + FAST_PAD + 3) >> 2;
  const bool aligned4 = ((ipitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(gbase) & 3) == 0);
  {
    // 7 row groups x 36 word columns (a tile row is 32 .. 36 words): 9 of 10 threads load, where 64 columns x 4 rows
    // left half of them idle
    const int ty = (tid * 1821) >> 16, tx = tid - ty * 36;   // tid / 36, tid % 36 for tid < 256
    if (ty < 7)
      for (int wv = tx; wv < nwords; wv += 36) {
        const int gx = iniX + 4 * wv - FAST_PAD;          // global x of the word's first byte (>= 15)
        for (int r = ty; r < rh; r += 7) {
          const u8* grow = gbase + (size_t)(iniY + r) * ipitch;
          uint32_t w;
          if (aligned4) {
            const int ax = gx & ~3, sh = (gx & 3) * 8;
            const uint32_t lo = __ldg(reinterpret_cast<const uint32_t*>(grow + ax));
            const uint32_t hi = __ldg(reinterpret_cast<const uint32_t*>(grow + min(ax + 4, (L.w - 1) & ~3)));
            w = __funnelshift_r(lo, hi, sh);
          } else {
            w = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) w |= (uint32_t)__ldg(grow + min(gx + k, L.w - 1)) << (8 * k);
          }
          reinterpret_cast<uint32_t*>(simg + r * tilePitch)[wv] = w;
        }
      }
  }
  for (int i = tid; i < (tileRows * tilePitch) >> 2; i += 256) reinterpret_cast<uint32_t*>(ssc)[i] = 0;
  __syncthreads();

  const int minTh = g.minTh, iniTh = g.iniTh;
  const int dw = rw - 6, dh = rh - 6;
  const int wCell = L.wCell;
  const float invWCell = 1.f / (float)wCell;
  // cell of detection column dx: (dx + 0.5) / wCell is at least 0.5 / wCell away from an integer, far beyond float rounding
  auto cell_of = [&](int dx) { return __float2int_rz(((float)dx + 0.5f) * invWCell); };
  // Two phases, as the reference does per cell (src/ORBextractor.cc:809-817): FAST at iniTh; the cells that are left
  // without a corner run FAST again at minTh.  The score does not depend on the threshold and a corner's 3x3 NMS only
  // loses to scores >= its own, so phase 0 is exactly cv::FAST(iniTh) + NMS, and phase 1 touches nothing but the
  // pixels of the empty cells (scores of neighbouring cells are masked in pass D).  Most of the work -- segment tests of
  // the pre-test's false positives, the arc scores -- is proportional to the corners at the threshold in use: on
  // textured frames 9 of 10 cells are settled at iniTh with a third of the corners minTh would raise.
  unsigned fb = 0u;   // phase 1: cells (bits) that run again at minTh
  for (int phase = 0; phase < 2; phase++) {
    const int th = phase == 0 ? iniTh : minTh;
    if (phase == 1) {
      for (int cidx = 0; cidx < t.ncells; cidx++) fb |= s_cellFlag[cidx] ? 0u : (1u << cidx);
      if (fb == 0u || iniTh == minTh) break;
      __syncthreads();   // every thread has read the flags and the list counters of phase 0
      if (tid == 0) { s_n1 = 0; s_n2 = 0; }
      __syncthreads();
    }
    // ---- pass A: 4 pixels per thread.  Any 9-arc of the 16-pixel ring covers two neighbouring compass
    // pixels (ring 0/4/8/12), so a corner needs two neighbouring compass pixels that are both brighter
    // than c + th or both darker than c - th.
    {
      const int dwords = (dw + 3) >> 2;
      const uint32_t th4 = (uint32_t)th * 0x01010101u;
      // warp = rows wid, wid + 8, ..., lane = word column (a detection row is 31 .. 33 words: the second trip is rare and
      // short).  kmask: pixels of the word that count -- inside the detection area and, in phase 1, inside an empty cell
      const int lane = tid & 31, wid = tid >> 5;
      for (int wv = lane; wv < dwords; wv += 32) {
        const int dx0 = wv * 4;
        unsigned kmask = dw - dx0 >= 4 ? 0xfu : ((1u << max(dw - dx0, 0)) - 1u);
        if (phase == 1) {
          unsigned fm = 0u;
#pragma unroll
          for (int k = 0; k < 4; k++) fm |= ((fb >> cell_of(min(dx0 + k, dw - 1))) & 1u) << k;
          kmask &= fm;
        }
        if (!kmask) continue;
        for (int dy = wid; dy < dh; dy += 8) {
          const uint32_t* row = reinterpret_cast<const uint32_t*>(simg + (dy + 3) * tilePitch) + 1 + wv;
          const uint32_t c = row[0], lft = row[-1], rgt = row[1];
          const uint32_t up = row[-3 * (tilePitch >> 2)], dn = row[3 * (tilePitch >> 2)];
          const uint32_t l3 = __funnelshift_r(lft, c, 8), r3 = __funnelshift_r(c, rgt, 24);   // columns -3 / +3
          const uint32_t hi4 = __vaddus4(c, th4), lo4 = __vsubus4(c, th4);                    // saturated: never passed if clipped
          const uint32_t bU = __vsetgtu4(up, hi4), bD = __vsetgtu4(dn, hi4), bL = __vsetgtu4(l3, hi4), bR = __vsetgtu4(r3, hi4);
          const uint32_t kU = __vsetltu4(up, lo4), kD = __vsetltu4(dn, lo4), kL = __vsetltu4(l3, lo4), kR = __vsetltu4(r3, lo4);
          const uint32_t cnt = ((bU | bD) & (bL | bR)) | ((kU | kD) & (kL | kR));             // 0x01 per candidate byte
          // the four flag bytes gathered into a nibble: byte k (bit 8k) times 2^(21 - 7k) lands on bit 21 + k, and no
          // two of the sixteen partial products share a bit
          unsigned m = ((cnt * 0x00204081u) >> 21) & kmask;
          if (!m) continue;
          int o = atomicAdd(&s_n1, __popc(m));
          const unsigned base = (unsigned)((dy << 8) | dx0);
          do {
            list1[o++] = (unsigned short)(base + (unsigned)(__ffs(m) - 1));
            m &= m - 1u;
          } while (m);
        }
      }
    }
    __syncthreads();
    // ---- pass B: exact segment test for the listed pixels; corners go to a second list
    const int n1 = s_n1;
    for (int i = tid; i < n1; i += 256) {
      const unsigned short pos = list1[i];
      const int dy = pos >> 8, dx = pos & 0xff;
      const u8* q = simg + (dy + 3) * tilePitch + 3 + FAST_PAD + dx;
      const int c = q[0];
      const int hi = c + th, lo = c - th;
      uint32_t mb = 0, md = 0;       // one bit per ring pixel: brighter than c+th / darker than c-th
#define FAST_RING(OFF)                                              \
      {                                                             \
        const int v = q[OFF];                                       \
        mb = __funnelshift_l((uint32_t)(hi - v), mb, 1);            \
        md = __funnelshift_l((uint32_t)(v - lo), md, 1);            \
      }
      FAST_RING(3 * tilePitch) FAST_RING(3 * tilePitch + 1) FAST_RING(2 * tilePitch + 2) FAST_RING(tilePitch + 3)
      FAST_RING(3) FAST_RING(-tilePitch + 3) FAST_RING(-2 * tilePitch + 2) FAST_RING(-3 * tilePitch + 1)
      FAST_RING(-3 * tilePitch) FAST_RING(-3 * tilePitch - 1) FAST_RING(-2 * tilePitch - 2) FAST_RING(-tilePitch - 3)
      FAST_RING(-3) FAST_RING(tilePitch - 3) FAST_RING(2 * tilePitch - 2) FAST_RING(3 * tilePitch - 1)
#undef FAST_RING
      mb |= mb << 16;
      md |= md << 16;
      mb &= mb >> 1; mb &= mb >> 2; mb &= mb >> 4; mb &= mb >> 1;
      md &= md >> 1; md &= md >> 2; md &= md >> 4; md &= md >> 1;
      if (((mb | md) & 0xffffu) == 0) continue;
      list2[atomicAdd(&s_n2, 1)] = pos;
    }
    __syncthreads();
    // ---- pass C: score of every corner (all lanes busy)
    const int n2 = s_n2;
    for (int i = tid; i < n2; i += 256) {
      const int dy = list2[i] >> 8, dx = list2[i] & 0xff;
      const u8* q = simg + (dy + 3) * tilePitch + 3 + FAST_PAD + dx;
      const int c = q[0];
      int d[16];
      d[0] = c - q[3 * tilePitch];        d[1] = c - q[3 * tilePitch + 1];   d[2] = c - q[2 * tilePitch + 2];
      d[3] = c - q[tilePitch + 3];        d[4] = c - q[3];                   d[5] = c - q[-tilePitch + 3];
      d[6] = c - q[-2 * tilePitch + 2];   d[7] = c - q[-3 * tilePitch + 1];  d[8] = c - q[-3 * tilePitch];
      d[9] = c - q[-3 * tilePitch - 1];   d[10] = c - q[-2 * tilePitch - 2]; d[11] = c - q[-tilePitch - 3];
      d[12] = c - q[-3];                  d[13] = c - q[tilePitch - 3];      d[14] = c - q[2 * tilePitch - 2];
      d[15] = c - q[3 * tilePitch - 1];
      ssc[(dy + 3) * tilePitch + 3 + FAST_PAD + dx] = (u8)fast_arc_score(d);
    }
    __syncthreads();
    // ---- pass D: 3x3 non-maximum suppression confined to the cell; every survivor is kept (phase 0: its cell is
    // settled; phase 1: its cell had no corner at iniTh)
    for (int i = tid; i < n2; i += 256) {
      const int dy = list2[i] >> 8, dx = list2[i] & 0xff;
      const u8* r1 = ssc + (dy + 3) * tilePitch + 3 + FAST_PAD + dx;
      const int sc = r1[0];
      const int cell = cell_of(dx), xin = dx - cell * wCell;
      // scores outside the detection area are 0 in the tile; only the cell's vertical borders need masking
      const int mL = xin > 0 ? 0xff : 0, mR = (xin < wCell - 1) ? 0xff : 0;
      const u8* r0 = r1 - tilePitch;
      const u8* r2 = r1 + tilePitch;
      const int m = max(max(max(r0[-1] & mL, r1[-1] & mL), max(r2[-1] & mL, r0[0])),
                        max(max(r2[0], r0[1] & mR), max(r1[1] & mR, r2[1] & mR)));
      if (m >= sc) continue;
      const int idx = atomicAdd(&s_nkept, 1);
      if (idx < survCap) kept[idx] = pack_xys(iniX + 3 + dx - kEdge, iniY + 3 + dy - kEdge, sc);
      if (phase == 0) s_cellFlag[cell] = 1;
    }
    __syncthreads();
  }
  const int nk = min(s_nkept, survCap);
  if (nk == 0) return;
  if (tid == 0) s_base = atomicAdd(&candCount[f * g.nlevels + t.level], nk);
  __syncthreads();
  const int base = s_base;
  uint32_t* out = cand + (size_t)f * g.candTotal + L.candOff;
  for (int i = tid; i < nk; i += 256)
    if (base + i < L.candCap) out[base + i] = kept[i];
}

// ---------------------------------------------------------------------------------
// k_octree: one CTA per (level, frame).  The reference's list-of-nodes algorithm is
// re-expressed as passes over (a) the keys -- each key knows the list position of its
// node -- and (b) the node list, which is rebuilt per pass as
//     reversed(children in creation order) ++ (surviving nodes in old order)
// which is exactly what push_front + erase produce.  Phase 1 splits every multi-key
// node per pass; phase 2 splits the multi-key nodes largest-first (ties: later created
// first, see oracle/oracle_orb.cpp) and stops as soon as the list holds N nodes.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int oct_quadrant(uint32_t key, ushort4 gm) {
  const int mx = gm.x + ((gm.z - gm.x + 1) >> 1);
  const int my = gm.y + ((gm.w - gm.y + 1) >> 1);
  return (unpack_x(key) < mx ? 0 : 1) + (unpack_y(key) < my ? 0 : 2);
}

template <int NT>
__global__ void __launch_bounds__(NT) k_octree(const __grid_constant__ OrbGeom g,
                                               const uint32_t* __restrict__ cand,
                                               const int* __restrict__ candCount,
                                               uint16_t* __restrict__ knodeAll,
                                               uint32_t* __restrict__ lvlKp,
                                               int* __restrict__ lvlCount) {
  extern __shared__ __align__(16) u8 smem[];
  const int lvl = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
  const OrbLevel& L = g.lv[lvl];
  const int M = g.maxNodes;
  const int n = min(candCount[f * g.nlevels + lvl], L.candCap);
  const int N = L.quota;
  if (n == 0) {
    if (tid == 0) lvlCount[f * g.nlevels + lvl] = 0;
    return;
  }
  const uint32_t* keys = cand + (size_t)f * g.candTotal + L.candOff;
  uint16_t* knode = knodeAll + (size_t)f * g.candTotal + L.candOff;

  // shared carve-up
  unsigned long long* best = reinterpret_cast<unsigned long long*>(smem);  // [M] (8B aligned first)
  ushort4* geomA = reinterpret_cast<ushort4*>(best + M);
  ushort4* geomB = geomA + M;
  int* cntA = reinterpret_cast<int*>(geomB + M);
  int* cntB = cntA + M;
  int* seqA = cntB + M;
  int* seqB = seqA + M;
  int* cc = seqB + M;          // [4M] child key counts
  int* order = cc + 4 * M;     // [M] node at processing rank r
  int* cstart = order + M;     // [M] first child creation index per rank
  int* keep = cstart + M;      // [M] survivor rank per position
  int* srank = keep + M;       // [M] processing rank per position or -1
  int* tmp = srank + M;        // [M] scratch
  unsigned short* cpos = reinterpret_cast<unsigned short*>(tmp + M);  // [4M] new position of child
  __shared__ int wtmp[33];
  __shared__ int s_rstar, s_nexp;

  // ---- root nodes (src/ORBextractor.cc:541-583)
  const int nIni = L.nIni;
  const float hX = L.hX;
  const int H = L.h - 2 * kEdge;
  for (int i = tid; i < nIni; i += NT) {
    geomA[i] = make_ushort4((unsigned short)(int)__fmul_rn(hX, (float)i), 0,
                            (unsigned short)(int)__fmul_rn(hX, (float)(i + 1)), (unsigned short)H);
    cntA[i] = 0;
    seqA[i] = i;
  }
  __syncthreads();
  for (int k = tid; k < n; k += NT) {
    int idx = (int)__fdiv_rn((float)unpack_x(keys[k]), hX);
    idx = min(idx, nIni - 1);
    atomicAdd(&cntA[idx], 1);
    knode[k] = (uint16_t)idx;
  }
  __syncthreads();
  for (int i = tid; i < nIni; i += NT) keep[i] = cntA[i] > 0;
  __syncthreads();
  int Lsz = block_excl_scan<NT>(keep, nIni, wtmp);
  for (int i = tid; i < nIni; i += NT)
    if (cntA[i] > 0) {
      geomB[keep[i]] = geomA[i];
      cntB[keep[i]] = cntA[i];
      seqB[keep[i]] = seqA[i];
    }
  for (int k = tid; k < n; k += NT) knode[k] = (uint16_t)keep[knode[k]];
  __syncthreads();
  ushort4 *geom = geomB, *geomN = geomA;
  int *cnt = cntB, *cntN = cntA, *seq = seqB, *seqN = seqA;
  int seqBase = nIni;
  int phase = 1;

  while (true) {
    const int prevL = Lsz;
    for (int i = tid; i < 4 * Lsz; i += NT) cc[i] = 0;
    if (tid == 0) { s_rstar = 0x7fffffff; s_nexp = 0; }
    __syncthreads();
    for (int k = tid; k < n; k += NT) {
      const int nd = knode[k];
      if (cnt[nd] > 1) atomicAdd(&cc[nd * 4 + oct_quadrant(keys[k], geom[nd])], 1);
    }
    __syncthreads();
    // multi-key nodes in list order -> order[0..Mc)
    for (int i = tid; i < Lsz; i += NT) tmp[i] = cnt[i] > 1;
    __syncthreads();
    const int Mc = block_excl_scan<NT>(tmp, Lsz, wtmp);
    for (int i = tid; i < Lsz; i += NT) {
      srank[i] = -1;
      if (cnt[i] > 1) order[tmp[i]] = i;
    }
    __syncthreads();
    int nSplit = Mc;
    if (phase == 2 && Mc > 0) {
      // processing order: (size, creation seq) descending  (src/ORBextractor.cc:679-683)
      for (int i = tid; i < Mc; i += NT) {
        const int a = order[i], ca = cnt[a], sa = seq[a];
        int r = 0;
        for (int j = 0; j < Mc; j++) {
          const int b = order[j], cb = cnt[b];
          r += (cb > ca) || (cb == ca && seq[b] > sa);
        }
        cstart[r] = a;  // sorted order, staged
      }
      __syncthreads();
      for (int i = tid; i < Mc; i += NT) {
        const int a = cstart[i];
        order[i] = a;
        const int* c4 = cc + a * 4;
        tmp[i] = (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0) - 1;
      }
      __syncthreads();
      // inclusive growth of the list; stop right after the split that reaches N (:728-729)
      for (int i = tid; i < Mc; i += NT) keep[i] = tmp[i];
      __syncthreads();
      block_excl_scan<NT>(keep, Mc, wtmp);
      for (int i = tid; i < Mc; i += NT)
        if (Lsz + keep[i] + tmp[i] >= N) atomicMin(&s_rstar, i);
      __syncthreads();
      nSplit = min(Mc, s_rstar == 0x7fffffff ? Mc : s_rstar + 1);
    }
    for (int r = tid; r < nSplit; r += NT) {
      const int a = order[r];
      srank[a] = r;
      const int* c4 = cc + a * 4;
      cstart[r] = (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0);
    }
    __syncthreads();
    const int C = block_excl_scan<NT>(cstart, nSplit, wtmp);
    for (int i = tid; i < Lsz; i += NT) keep[i] = srank[i] < 0;
    __syncthreads();
    const int K = block_excl_scan<NT>(keep, Lsz, wtmp);
    const int Lnew = C + K;
    // build the new list (positions double as node ids)
    int myExp = 0;
    for (int r = tid; r < nSplit; r += NT) {
      const int a = order[r];
      const ushort4 gm = geom[a];
      const unsigned short mx = gm.x + ((gm.z - gm.x + 1) >> 1);
      const unsigned short my = gm.y + ((gm.w - gm.y + 1) >> 1);
      int ci = cstart[r];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int c = cc[a * 4 + q];
        if (c == 0) continue;
        const int np = C - 1 - ci;
        ushort4 ch;
        ch.x = (q & 1) ? mx : gm.x;
        ch.z = (q & 1) ? gm.z : mx;
        ch.y = (q & 2) ? my : gm.y;
        ch.w = (q & 2) ? gm.w : my;
        geomN[np] = ch;
        cntN[np] = c;
        seqN[np] = seqBase + ci;
        cpos[a * 4 + q] = (unsigned short)np;
        myExp += c > 1;
        ci++;
      }
    }
    for (int i = tid; i < Lsz; i += NT)
      if (srank[i] < 0) {
        const int np = C + keep[i];
        geomN[np] = geom[i];
        cntN[np] = cnt[i];
        seqN[np] = seq[i];
      }
    if (myExp) atomicAdd(&s_nexp, myExp);
    __syncthreads();
    for (int k = tid; k < n; k += NT) {
      const int nd = knode[k];
      knode[k] = srank[nd] >= 0 ? cpos[nd * 4 + oct_quadrant(keys[k], geom[nd])]
                                : (uint16_t)(C + keep[nd]);
    }
    __syncthreads();
    const int nToExpand = s_nexp;
    { ushort4* t = geom; geom = geomN; geomN = t; }
    { int* t = cnt; cnt = cntN; cntN = t; }
    { int* t = seq; seq = seqN; seqN = t; }
    Lsz = Lnew;
    seqBase += C;
    if (Lsz >= N || Lsz == prevL) break;
    if (phase == 1 && Lsz + 3 * nToExpand > N) phase = 2;
    __syncthreads();
  }

  // ---- best key per node: max response, first in the reference's emission order
  // (cell row, cell col, y, x) wins ties (src/ORBextractor.cc:742-758)
  for (int i = tid; i < Lsz; i += NT) best[i] = 0ull;
  __syncthreads();
  for (int k = tid; k < n; k += NT) {
    const uint32_t key = keys[k];
    const int x = unpack_x(key), y = unpack_y(key);
    const unsigned cr = 255u - (unsigned)((y - 3) / L.hCell), ccol = 255u - (unsigned)((x - 3) / L.wCell);
    const unsigned long long v = ((unsigned long long)unpack_s(key) << 40) |
                                 ((unsigned long long)cr << 32) | ((unsigned long long)ccol << 24) |
                                 ((unsigned long long)(4095 - y) << 12) | (unsigned long long)(4095 - x);
    atomicMax(&best[knode[k]], v);
  }
  __syncthreads();
  const int nOut = min(Lsz, L.kpCap);
  uint32_t* out = lvlKp + (size_t)f * g.kpTotal + L.kpOff;
  for (int i = tid; i < nOut; i += NT) {
    const unsigned long long v = best[i];
    const int x = 4095 - (int)(v & 0xFFF), y = 4095 - (int)((v >> 12) & 0xFFF);
    out[i] = pack_xys(x + kEdge, y + kEdge, (int)(v >> 40));
  }
  if (tid == 0) lvlCount[f * g.nlevels + lvl] = nOut;
}

// ---------------------------------------------------------------------------------
// k_blur7: GaussianBlur(7x7, sigma 2, REFLECT_101) of every level, fixed point
// {18,34,48,56,48,34,18}/256 per axis, (v + 32768) >> 16.  Tile 128x32 outputs.
// ---------------------------------------------------------------------------------
#define BLUR_TW 128
#define BLUR_TH 32
// BORDER_REFLECT_101 for -len < p < 2 * len - 1 (one reflection: the halos are 3 pixels, levels >= 30)
__device__ __forceinline__ int reflect101(int p, int len) {
  p = p < 0 ? -p : p;
  return p >= len ? 2 * (len - 1) - p : p;
}

// Tile = 128 x 32 outputs.  The input tile (38 rows x 136 bytes, starting 4 bytes left of the
// tile so that words stay aligned) is loaded as 32-bit words; the horizontal pass builds the
// two 4-byte tap windows of each pixel with funnel shifts and uses two u8 dot products
// (IDP.4A); the vertical pass works on 32-bit row sums.
__global__ void __launch_bounds__(256) k_blur7(const __grid_constant__ OrbGeom g,
                                               const __grid_constant__ OrbPtrs p,
                                               const BlurTile* __restrict__ tiles) {
  __shared__ __align__(16) uint32_t sin_[(BLUR_TH + 6) * 36];          // 34 words used per row
  __shared__ __align__(16) uint32_t sh_[(BLUR_TH + 6) * BLUR_TW];      // horizontal sums (<= 65280)
  const BlurTile t = tiles[blockIdx.x];
  const int f = blockIdx.y, tid = threadIdx.x;
  const OrbLevel& L = g.lv[t.level];
  const int x0 = t.tx * BLUR_TW, y0 = t.ty * BLUR_TH;
  const int ipitch = p.ipitch[t.level];
  const u8* src = p.img[t.level] + (size_t)f * p.ifs[t.level];
  const bool aligned4 = ((ipitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0);
  // Input tile: 7 row groups x 36 word columns (34 used): the column's place relative to the image border is decided
  // once per thread, and a thread's six row loads are in flight together (one load per loop trip with a division and
  // the border tests around it was 30 % of the kernel's instructions and most of its stalls).
  {
    const int ty = (tid * 1821) >> 16, wq = tid - ty * 36;   // tid / 36, tid % 36 for tid < 256
    if (ty < 7 && wq < 34) {
      const int gx = x0 - 4 + 4 * wq;
      // 0: aligned word inside the row, 1: touches the border (bytes, reflected), 2: beyond the reflected border
      const int cls = (aligned4 && gx >= 0 && gx + 3 < L.w) ? 0 : (gx > L.w + 2 ? 2 : 1);
      uint32_t w[6];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const int r = ty + 7 * k, yy = y0 + r - 3;
        w[k] = 0;   // beyond the reflected border: read by no output pixel
        if (r < BLUR_TH + 6 && cls != 2 && yy <= L.h + 2) {
          const u8* row = src + (size_t)reflect101(yy, L.h) * ipitch;
          if (cls == 0) {
            w[k] = __ldg(reinterpret_cast<const uint32_t*>(row + gx));
          } else {
#pragma unroll
            for (int kk = 0; kk < 4; kk++) w[k] |= (uint32_t)__ldg(row + reflect101(min(gx + kk, L.w + 2), L.w)) << (8 * kk);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const int r = ty + 7 * k;
        if (r < BLUR_TH + 6) sin_[r * 36 + wq] = w[k];
      }
    }
  }
  __syncthreads();
  const uint32_t K0 = 18u | (34u << 8) | (48u << 16) | (56u << 24), K1 = 48u | (34u << 8) | (18u << 16);
  for (int i = tid; i < (BLUR_TH + 6) * (BLUR_TW / 4); i += 256) {
    const int r = i >> 5, j = i & 31;               // output pixels 4j .. 4j+3 of row r
    const uint32_t* q = sin_ + r * 36 + j;          // words holding tile bytes 4j-4 .. 4j+7
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
    uint4 o;
    o.x = __dp4a(__funnelshift_r(w0, w1, 8), K0, __dp4a(__funnelshift_r(w1, w2, 8), K1, 0u));
    o.y = __dp4a(__funnelshift_r(w0, w1, 16), K0, __dp4a(__funnelshift_r(w1, w2, 16), K1, 0u));
    o.z = __dp4a(__funnelshift_r(w0, w1, 24), K0, __dp4a(__funnelshift_r(w1, w2, 24), K1, 0u));
    o.w = __dp4a(w1, K0, __dp4a(w2, K1, 0u));
    *reinterpret_cast<uint4*>(sh_ + r * BLUR_TW + 4 * j) = o;
  }
  __syncthreads();
  u8* dst = p.blur[t.level] + (size_t)f * p.bfs[t.level];
  for (int i = tid; i < (BLUR_TH / 2) * (BLUR_TW / 4); i += 256) {
    const int r = (i >> 5) * 2, c4 = (i & 31) * 4;  // output rows r, r+1; columns c4 .. c4+3
    const int gx = x0 + c4;
    if (gx >= L.w) continue;
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = *reinterpret_cast<const uint4*>(sh_ + (r + k) * BLUR_TW + c4);
    uint32_t out0 = 0, out1 = 0;
#define BLUR_COL(FIELD, SHIFT)                                                                             \
    {                                                                                                      \
      const uint32_t a = 18u * (v[0].FIELD + v[6].FIELD) + 34u * (v[1].FIELD + v[5].FIELD) +               \
                         48u * (v[2].FIELD + v[4].FIELD) + 56u * v[3].FIELD;                               \
      const uint32_t bq = 18u * (v[1].FIELD + v[7].FIELD) + 34u * (v[2].FIELD + v[6].FIELD) +              \
                          48u * (v[3].FIELD + v[5].FIELD) + 56u * v[4].FIELD;                              \
      out0 |= ((a + 32768u) >> 16) << SHIFT;                                                               \
      out1 |= ((bq + 32768u) >> 16) << SHIFT;                                                              \
    }
    BLUR_COL(x, 0) BLUR_COL(y, 8) BLUR_COL(z, 16) BLUR_COL(w, 24)
#undef BLUR_COL
    if (y0 + r < L.h) *reinterpret_cast<uint32_t*>(dst + (size_t)(y0 + r) * L.pitch + gx) = out0;
    if (y0 + r + 1 < L.h) *reinterpret_cast<uint32_t*>(dst + (size_t)(y0 + r + 1) * L.pitch + gx) = out1;
  }
}

// ---------------------------------------------------------------------------------
// k_layout: output row of every staged keypoint.  Keypoints are visited level by level
// in octree list order; those with lap0 <= x_scaled <= lap1 fill the output from the
// back, the others from the front (src/ORBextractor.cc:1104-1144).
// ---------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k_layout(const __grid_constant__ OrbGeom g,
                                               const uint32_t* __restrict__ lvlKp,
                                               const int* __restrict__ lvlCount, int lap0,
                                               int lap1, int* __restrict__ slot,
                                               int* __restrict__ counts, int* __restrict__ mono) {
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  __shared__ int wv[NT / 32], wl[NT / 32], tot[2];
  const int total = g.kpTotal;
  const int chunk = (total + NT - 1) / NT;
  const int beg = min(tid * chunk, total), end = min(beg + chunk, total);
  const uint32_t* kp = lvlKp + (size_t)f * total;
  const int* lc = lvlCount + f * g.nlevels;
  auto classify = [&](int s, bool& valid, bool& lap) {
    int l = 0;
    while (l + 1 < g.nlevels && s >= g.lv[l + 1].kpOff) l++;
    valid = (s - g.lv[l].kpOff) < lc[l];
    lap = false;
    if (valid) {
      float x = (float)unpack_x(kp[s]);
      if (l) x = __fmul_rn(x, g.lv[l].scale);
      lap = x >= (float)lap0 && x <= (float)lap1;
    }
  };
  int nv = 0, nl = 0;
  for (int s = beg; s < end; s++) {
    bool v, l;
    classify(s, v, l);
    nv += v;
    nl += l;
  }
  int iv = nv, il = nl;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int a = __shfl_up_sync(0xffffffffu, iv, o), b = __shfl_up_sync(0xffffffffu, il, o);
    if (lane >= o) { iv += a; il += b; }
  }
  if (lane == 31) { wv[wid] = iv; wl[wid] = il; }
  __syncthreads();
  if (tid == 0) {
    int a = 0, b = 0;
    for (int w = 0; w < NT / 32; w++) {
      int ta = wv[w], tb = wl[w];
      wv[w] = a; wl[w] = b;
      a += ta; b += tb;
    }
    tot[0] = a; tot[1] = b;
  }
  __syncthreads();
  int bv = wv[wid] + iv - nv, bl = wl[wid] + il - nl;
  const int n = tot[0];
  for (int s = beg; s < end; s++) {
    bool v, l;
    classify(s, v, l);
    int o = -1;
    if (v) o = l ? (n - 1 - bl) : (bv - bl);
    slot[(size_t)f * total + s] = o;
    bv += v;
    bl += l;
  }
  if (tid == 0) {
    counts[f] = n;
    mono[f] = n - tot[1];
  }
}

// ---------------------------------------------------------------------------------
// k_orient_desc: warp per keypoint.  IC_Angle on the unblurred level (lane = column
// offset u in [-15,15]), cv::fastAtan2 polynomial, then 256 rotated pair tests on the
// blurred level (lane = descriptor byte).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float dev_fast_atan2(float y, float x) {
  const float sc = 57.29577951308232f;
  const float p1 = __fmul_rn(0.9997878412794807f, sc), p3 = __fmul_rn(-0.3258083974640975f, sc);
  const float p5 = __fmul_rn(0.1555786518463281f, sc), p7 = __fmul_rn(-0.04432655554792128f, sc);
  const float eps = 2.220446049250313e-16f;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

#define OD_WARPS 4
#ifndef OD_MINB
#define OD_MINB 5   // 96 registers: 5.4 ms per 4096 frames (128 registers: slower step, 80 / 72 with spills: 5.7 / 6.2 ms)
#endif
__global__ void __launch_bounds__(OD_WARPS * 32, OD_MINB) k_orient_desc(
    const __grid_constant__ OrbGeom g, const __grid_constant__ OrbPtrs p,
    const uint32_t* __restrict__ lvlKp, const int* __restrict__ lvlCount,
    const int* __restrict__ slot, int kpPerCta, plvi_keypoint* __restrict__ kps,
    uint8_t* __restrict__ desc, int cap) {
  const int f = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // this lane's 8 pair tests
  float px[16], py[16];   // kept as floats: the rotation below would convert them for every keypoint
#pragma unroll
  for (int j = 0; j < 16; j++) {
    px[j] = (float)d_pattern[lane * 32 + j * 2];
    py[j] = (float)d_pattern[lane * 32 + j * 2 + 1];
  }
  const int sBeg = blockIdx.x * kpPerCta, sEnd = min(sBeg + kpPerCta, g.kpTotal);
  for (int s = sBeg + wid; s < sEnd; s += OD_WARPS) {
    int l = 0;
    while (l + 1 < g.nlevels && s >= g.lv[l + 1].kpOff) l++;
    if (s - g.lv[l].kpOff >= lvlCount[f * g.nlevels + l]) continue;
    const uint32_t pk = lvlKp[(size_t)f * g.kpTotal + s];
    const int x = unpack_x(pk), y = unpack_y(pk);
    const int ipitch = p.ipitch[l];
    const u8* c0 = p.img[l] + (size_t)f * p.ifs[l] + (size_t)y * ipitch + x;
    const int u = lane - 15;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
      const int au = abs(u);
      // fully unrolled: the row bounds become constants and the 31 loads of the patch column are in flight together
      // (the kernel waited on them four at a time: half of its stall samples)
      int vals[31];
#pragma unroll
      for (int v = -15; v <= 15; v++) vals[v + 15] = (au <= c_umax[v < 0 ? -v : v]) ? (int)__ldg(c0 + v * ipitch + u) : 0;
      int colsum = 0;
#pragma unroll
      for (int v = -15; v <= 15; v++) {
        colsum += vals[v + 15];
        m01 += v * vals[v + 15];
      }
      m10 = u * colsum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m10 += __shfl_xor_sync(0xffffffffu, m10, o);
      m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = dev_fast_atan2((float)m01, (float)m10);
    float a = 0.f, b = 0.f;
    if (lane == 0) {
      const float rad = __fmul_rn(angle, 0.017453292519943295f);
      glibc_sincosf(rad, b, a);   // the reference's cos(float) / sin(float): host libm cosf / sinf
    }
    a = __shfl_sync(0xffffffffu, a, 0);
    b = __shfl_sync(0xffffffffu, b, 0);
    const int bp = g.lv[l].pitch;
    const u8* cb = p.blur[l] + (size_t)f * p.bfs[l] + (size_t)y * bp + x;
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float x0 = px[2 * j], y0 = py[2 * j];
      const float x1 = px[2 * j + 1], y1 = py[2 * j + 1];
      const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
      const int q0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
      const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
      const int q1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
      const int t0 = __ldg(cb + r0 * bp + q0), t1 = __ldg(cb + r1 * bp + q1);
      val |= (t0 < t1) << j;
    }
    const int o = slot[(size_t)f * g.kpTotal + s];
    desc[((size_t)f * cap + o) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
      plvi_keypoint k;
      const float sc = g.lv[l].scale;
      k.x = l ? __fmul_rn((float)x, sc) : (float)x;
      k.y = l ? __fmul_rn((float)y, sc) : (float)y;
      k.size = (float)g.lv[l].sizeField;
      k.angle = angle;
      k.response = (float)unpack_s(pk);
      k.octave = l;
      k.class_id = -1;
      kps[(size_t)f * cap + o] = k;
    }
  }
}

void launch_resize_u8(const u8* src, int spitch, size_t sfs, int sw, int sh, u8* dst, int dpitch, size_t dfs,
                      int dw, int dh, const int2* xtab, const int2* ytab, int n, cudaStream_t st) {
  // staged variant: 16-byte aligned rows and a source window that fits the shared tile (scale factors up to 2)
  const bool aligned = ((uintptr_t)src & 15) == 0 && (spitch & 15) == 0 && (sfs & 15) == 0 && spitch >= ((sw + 15) & ~15);
  const long spanW = ((long)RS_TW * sw + dw - 1) / dw + 3 + 32, spanH = ((long)RS_TH * sh + dh - 1) / dh + 3;
  static const bool direct = [] { const char* e = getenv("PLVI_RESIZE_DIRECT"); return e && e[0] == '1'; }();
  if (!direct && aligned && spanW <= RS_SRC_W && spanH <= RS_SRC_H) {
    dim3 grd((dw + RS_TW - 1) / RS_TW, (dh + RS_TH - 1) / RS_TH, n);
    k_resize_tma<<<grd, 256, 0, st>>>(src, spitch, sfs, sw, sh, dst, dpitch, dfs, dw, dh, xtab, ytab);
    return;
  }
  dim3 blk(64, 4), grd((dw + 255) / 256, (dh + 3) / 4, n);
  k_resize<<<grd, blk, 0, st>>>(src, spitch, sfs, sw, sh, dst, dpitch, dfs, dw, dh, xtab, ytab);
}

// ---------------------------------------------------------------------------------
// host-side launch sequence
// ---------------------------------------------------------------------------------
#define OCT_NT 256

static size_t octree_smem_bytes(int M) {
  // best 8 | geom 2x8 | cnt 2x4 | seq 2x4 | cc 16 | order,cstart,keep,srank,tmp 5x4 | cpos 8
  return (size_t)M * (8 + 16 + 8 + 8 + 16 + 20 + 8) + 64;
}

static void fast_tile_dims(const OrbGeom& g, int* tilePitch, int* tileRows, int* listCap, int* survCap) {
  int maxW = 0, maxH = 0, maxSurv = 0, maxList = 0;
  for (int l = 0; l < g.nlevels; l++) {
    const OrbLevel& L = g.lv[l];
    maxW = max(maxW, 4 * L.wCell + 6);
    maxH = max(maxH, L.hCell + 6);
    maxSurv = max(maxSurv, 4 * ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2));
    maxList = max(maxList, 4 * L.wCell * L.hCell);
  }
  *tilePitch = (maxW + 1 + 4 + 15) & ~15;   // +1 pad column, +4: word reads one word past the last detection word
  *tileRows = maxH;
  *listCap = (maxList + 7) & ~7;
  *survCap = maxSurv;
}

int orb_kernel_attrs(const OrbGeom& g, int* fastSmem, int* octSmem) {
  int tp, tr, lc, sc;
  fast_tile_dims(g, &tp, &tr, &lc, &sc);
  *fastSmem = 2 * tp * tr + 2 * lc * (int)sizeof(unsigned short) + sc * (int)sizeof(uint32_t);
  *octSmem = (int)octree_smem_bytes(g.maxNodes);
  if (*octSmem > 48 * 1024)
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_octree<OCT_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, *octSmem));
  if (*fastSmem > 48 * 1024)
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, *fastSmem));
  return PLVI_OK;
}

int launch_orb_pipeline(const OrbGeom& g, const OrbPtrs& p, const OrbScratch& s, int n, int lap0,
                        int lap1, plvi_keypoint* d_kps, uint8_t* d_desc, int* d_counts,
                        int* d_mono, int cap, cudaStream_t st, int* launches, StageProf* prof, cudaEvent_t waitAfterPyramid, int stages) {
  // stages: bit 0 = the pyramid, bit 1 = everything after it (plvi_orb_pyramid_device runs the pyramid ahead)
  int nl = 0;
  StageProf nop;
  if (!prof) prof = &nop;
  prof->begin(st);
  if (stages & 1) {
    // pyramid (chained: level l from level l-1)
    for (int l = 1; l < g.nlevels; l++) {
      const OrbLevel& d = g.lv[l];
      const OrbLevel& sl = g.lv[l - 1];
      launch_resize_u8(p.img[l - 1], p.ipitch[l - 1], p.ifs[l - 1], sl.w, sl.h, const_cast<u8*>(p.img[l]), p.ipitch[l],
                       p.ifs[l], d.w, d.h, s.rsTab + d.rsOff, s.rsTab + d.rsOff + d.w, n, st);
      nl++;
    }
    prof->mark("k_resize", st);
  }
  if (!(stages & 2)) {
    PLVI_CUDA_TRY(cudaGetLastError());
    if (launches) *launches = nl;
    return PLVI_OK;
  }
  PLVI_CUDA_TRY(cudaMemsetAsync(s.candCount, 0, sizeof(int) * (size_t)n * g.nlevels, st));
  if (waitAfterPyramid) {   // plvi_orb_wait_event_after_pyramid: the rest of the sequence follows the caller's event
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    PLVI_CUDA_TRY(cudaStreamWaitEvent(st, waitAfterPyramid, cs == cudaStreamCaptureStatusActive ? cudaEventWaitExternal : cudaEventWaitDefault));
    prof->mark("(wait)", st);
  }
  // FAST tile geometry recomputed here must match capi's smem sizing
  {
    int tilePitch, maxH, listCap, maxSurv;
    fast_tile_dims(g, &tilePitch, &maxH, &listCap, &maxSurv);
    k_fast<<<dim3(s.nFastTiles, n), 256, s.fastSmem, st>>>(g, p, s.fastTiles, s.cand, s.candCount,
                                                          tilePitch, maxH, listCap, maxSurv);
    nl++;
  }
  prof->mark("k_fast", st);
  k_octree<OCT_NT><<<dim3(g.nlevels, n), OCT_NT, s.octSmem, st>>>(g, s.cand, s.candCount, s.knode,
                                                               s.lvlKp, s.lvlCount);
  nl++;
  prof->mark("k_octree", st);
  k_blur7<<<dim3(s.nBlurTiles, n), 256, 0, st>>>(g, p, s.blurTiles);
  nl++;
  prof->mark("k_blur7", st);
  k_layout<256><<<n, 256, 0, st>>>(g, s.lvlKp, s.lvlCount, lap0, lap1, s.slot, d_counts, d_mono);
  nl++;
  prof->mark("k_layout", st);
  const int kpPerCta = 32;
  k_orient_desc<<<dim3((g.kpTotal + kpPerCta - 1) / kpPerCta, n), OD_WARPS * 32, 0, st>>>(
      g, p, s.lvlKp, s.lvlCount, s.slot, kpPerCta, d_kps, d_desc, cap);
  nl++;
  prof->mark("k_orient_desc", st);
  PLVI_CUDA_TRY(cudaGetLastError());
  if (launches) *launches = nl;
  return PLVI_OK;
}

}  // namespace plvi
