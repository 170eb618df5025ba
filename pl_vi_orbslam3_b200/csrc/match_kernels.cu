// Hamming descriptor search kernels for sm_100a + their C ABI.
//
//   k_hamming_pairs   ORBmatcher::DescriptorDistance / LineMatcher::distance /
//                     LineMatcher::DescriptorDistance      src/ORBmatcher.cc:2350-2366,
//                                                          src/LineMatcher.cpp:173-189,487-499
//   k_search          ORBmatcher::SearchByProjection(F,F)  src/ORBmatcher.cc:1962-2178
//                     ORBmatcher::SearchByProjection(F,MP) src/ORBmatcher.cc:44-214
//                     ORBmatcher::SearchForInitialization  src/ORBmatcher.cc:706-820
//                     + Frame::AssignFeaturesToGrid / GetFeaturesInArea src/Frame.cc:644-675,1006-1087
//                     + ComputeThreeMaxima                 src/ORBmatcher.cc:2304-2345
//   k_line_match      LineMatcher::matchNNR / match        src/LineMatcher.cpp:41-111
//
// The reference searches are sequential over the queries: a query skips train features
// that an earlier query claimed (or, for initialisation, matched at a smaller distance).
// k_search keeps that order exactly but takes the Hamming work out of the serial chain:
// one CTA per frame pair; its warps evaluate W consecutive queries in parallel against
// the state as of the chunk start, then warp 0 commits them in query order and
// re-evaluates the rare query whose best / second-best candidate was taken by an earlier
// query of the same chunk.  "best" and "second best" are the two smallest candidates
// under the key (distance, position in GetFeaturesInArea order), which is what the
// reference's `<` update chain produces.
#include <cstring>
#include <vector>

#include "plvi_internal.cuh"

namespace plvi {

#define GRID_COLS 64
#define GRID_ROWS 48
#define GRID_CELLS (GRID_COLS * GRID_ROWS)
#define HISTO_LENGTH 30
#define SEARCH_WARPS 8
#ifndef SEARCH_PER_WARP
#define SEARCH_PER_WARP 2
#endif
#define SEARCH_CHUNK (SEARCH_PER_WARP * (SEARCH_WARPS - 1))   // queries per chunk of k_search: two per evaluating warp (four: 1.80 vs 1.70 ms per 1024 frames)

__device__ __forceinline__ int hamming256_regs(const uint32_t (&q)[8], const uint8_t* __restrict__ d) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(d));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(d) + 1);
  return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
         __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
}

__global__ void k_hamming_pairs(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n,
                                int shift25, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* pa = reinterpret_cast<const uint32_t*>(a + (size_t)i * 32);
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(b + (size_t)i * 32);
  int d = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int c = __popc(__ldg(pa + k) ^ __ldg(pb + k));
    d += shift25 ? (c >> 1) : c;   // the >>25 variant sums floor(popcount/2) per word
  }
  out[i] = d;
}

// top-2 by (key) with payload
struct Top2 {
  uint32_t k0, k1;  // keys: dist << 23 | cell ordinal << 11 | rank in cell
  int i0, i1;
};
__device__ __forceinline__ void top2_insert(Top2& t, uint32_t k, int i) {
  if (k < t.k0) { t.k1 = t.k0; t.i1 = t.i0; t.k0 = k; t.i0 = i; }
  else if (k < t.k1) { t.k1 = k; t.i1 = i; }
}
__device__ __forceinline__ void top2_warp_merge(Top2& t) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t ok0 = __shfl_xor_sync(0xffffffffu, t.k0, o), ok1 = __shfl_xor_sync(0xffffffffu, t.k1, o);
    const int oi0 = __shfl_xor_sync(0xffffffffu, t.i0, o), oi1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
    top2_insert(t, ok0, oi0);
    top2_insert(t, ok1, oi1);
  }
}

struct SearchArgs {
  int mode;  // 0 frame-frame, 1 map points, 2 initialisation, 3 bag-of-words groups
  const int* items; int istride;   // mode 3: frame features grouped by vocabulary node
  const plvi_keypoint* keys; const uint8_t* desc; const uint8_t* blocked; const int* tcount; int tstride;
  plvi_grid grid;
  plvi_query* q; const uint8_t* qdesc; const int* qcount; int qstride;
  int th; float nnratio; int checkOri;
  int* matchTrain; int* matchQuery; int* nmatches;
  int* matchedDist;  // scratch [P][tstride] (init mode)
  // rectified-stereo side information (plvi_matcher_set_stereo), or nullptr: mvuRight of the train frame and the
  // right-image coordinate of every query's projection
  const float* uright; const float* qur;
};

struct QRes {
  int best, bestDist, second, secondDist;  // indices or -1; dist INT_MAX when absent
};
// what the in-order commit needs besides the result, fetched by the evaluating warps (eight at a time) so that the
// serial commit reads shared memory only: the query record and the candidates' key points were two dependent L2 round
// trips per committed query
struct QAux {
  int flags;            // query flags (bit 0: skip)
  float qangle;         // query angle
  float bangle;         // angle of the best candidate's key point
  int l1, l2;           // octaves of the best / second candidate (-1: none)
};

__device__ __forceinline__ bool eligible(const SearchArgs& a, int i2, int d, const uint8_t* blk,
                                         const unsigned short* mdist) {
  if (a.mode == 2) return !(mdist[i2] <= d);
  return !blk[i2];
}

// Evaluate one query with a full warp.  cellStart/items describe mGrid in smem.
__device__ QRes eval_query(const SearchArgs& a, const plvi_keypoint* __restrict__ keys,
                           const uint8_t* __restrict__ desc, const plvi_query& q,
                           const uint8_t* __restrict__ qd, const int* cellStart,
                           const unsigned short* items, const uint8_t* ioct, const uint8_t* blk, const unsigned short* mdist,
                           const float* __restrict__ uright, float qur) {
  const int lane = threadIdx.x & 31;
  QRes r = {-1, 0x7fffffff, -1, 0x7fffffff};
  const float x = q.u, y = q.v, rad = q.radius;
  const plvi_grid& g = a.grid;
  const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, g.min_x), rad), g.inv_w)));
  const int cx1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, g.min_x), rad), g.inv_w)));
  const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, g.min_y), rad), g.inv_h)));
  const int cy1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, g.min_y), rad), g.inv_h)));
  if (cx0 >= GRID_COLS || cx1 < 0 || cy0 >= GRID_ROWS || cy1 < 0) return r;
  const bool checkLevels = (q.min_level > 0) || (q.max_level >= 0);
  uint32_t qw[8];
#pragma unroll
  for (int k = 0; k < 8; k++) qw[k] = __ldg(reinterpret_cast<const uint32_t*>(qd) + k);
  const int ny = cy1 - cy0 + 1, nc = (cx1 - cx0 + 1) * ny;
  Top2 t = {0xffffffffu, 0xffffffffu, -1, -1};
  for (int c = lane; c < nc; c += 32) {
    const int ix = cx0 + c / ny, iy = cy0 + c % ny;
    const int cell = ix * GRID_ROWS + iy;
    const int s = cellStart[cell], e = cellStart[cell + 1];
    for (int k = s; k < e; k++) {
      const int i2 = items[k];
      if (checkLevels) {   // from the shared-memory copy of the octaves: most candidates of a wide window end here
        const int oc = ioct[k];
        if (oc < q.min_level) continue;
        if (q.max_level >= 0 && oc > q.max_level) continue;
      }
      if (a.mode != 2 && blk[i2]) continue;   // (the initialisation mode decides by the distance: eligible())
      const float kx = keys[i2].x, ky = keys[i2].y;
      if (!(fabsf(__fsub_rn(kx, x)) < rad && fabsf(__fsub_rn(ky, y)) < rad)) continue;
      if (uright) {   // stereo observation: |ur - mvuRight[i2]| must stay inside the window (src/ORBmatcher.cc:91-96, 2041-2047)
        const float ur2 = __ldg(uright + i2);
        if (ur2 > 0.f && fabsf(__fsub_rn(qur, ur2)) > rad) continue;
      }
      const int d = hamming256_regs(qw, desc + (size_t)i2 * 32);
      if (!eligible(a, i2, d, blk, mdist)) continue;
      top2_insert(t, ((uint32_t)d << 23) | ((uint32_t)c << 11) | (uint32_t)min(k - s, 2047), i2);
    }
  }
  top2_warp_merge(t);
  if (t.i0 >= 0) { r.best = t.i0; r.bestDist = (int)(t.k0 >> 23); }
  if (t.i1 >= 0) { r.second = t.i1; r.secondDist = (int)(t.k1 >> 23); }
  return r;
}

// Mode 3 (SearchByBoW): candidates are the frame features of the query's vocabulary node,
// items[min_level .. max_level), in vIndicesF order.
__device__ QRes eval_query_bow(const uint8_t* __restrict__ desc, const plvi_query& q, const uint8_t* __restrict__ qd,
                               const int* __restrict__ items, const uint8_t* blk) {
  const int lane = threadIdx.x & 31;
  QRes r = {-1, 0x7fffffff, -1, 0x7fffffff};
  uint32_t qw[8];
#pragma unroll
  for (int k = 0; k < 8; k++) qw[k] = __ldg(reinterpret_cast<const uint32_t*>(qd) + k);
  Top2 t = {0xffffffffu, 0xffffffffu, -1, -1};
  for (int k = q.min_level + lane; k < q.max_level; k += 32) {
    const int i2 = __ldg(items + k);
    if (blk[i2]) continue;
    const int d = hamming256_regs(qw, desc + (size_t)i2 * 32);
    top2_insert(t, ((uint32_t)d << 23) | (uint32_t)min(k - q.min_level, (1 << 23) - 1), i2);
  }
  top2_warp_merge(t);
  if (t.i0 >= 0) { r.best = t.i0; r.bestDist = (int)(t.k0 >> 23); }
  if (t.i1 >= 0) { r.second = t.i1; r.secondDist = (int)(t.k1 >> 23); }
  return r;
}

__device__ __forceinline__ int rot_bin(float a1, float a2) {
  float rot = __fsub_rn(a1, a2);
  if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
  int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_search(const SearchArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int NT = SEARCH_WARPS * 32;
  const int n = min(a.tcount[pair], a.tstride), nq = min(a.qcount[pair], a.qstride);
  const plvi_keypoint* keys = a.keys + (size_t)pair * a.tstride;
  const uint8_t* desc = a.desc + (size_t)pair * a.tstride * 32;
  const float* uright = (a.uright && a.qur && a.mode < 2) ? a.uright + (size_t)pair * a.tstride : nullptr;
  const float* qur = uright ? a.qur + (size_t)pair * a.qstride : nullptr;
  plvi_query* q = a.q + (size_t)pair * a.qstride;
  const uint8_t* qdesc = a.qdesc + (size_t)pair * a.qstride * 32;
  int* owner = a.matchTrain + (size_t)pair * a.tstride;   // train -> query (mvpMapPoints / vnMatches21)
  int* m12 = a.matchQuery + (size_t)pair * a.qstride;     // query -> train

  int* cellStart = reinterpret_cast<int*>(smem);                 // [GRID_CELLS + 1]
  int* cursor = cellStart + GRID_CELLS + 1;                      // [GRID_CELLS]
  unsigned short* items = reinterpret_cast<unsigned short*>(cursor + GRID_CELLS);  // [tstride]
  uint8_t* blk = reinterpret_cast<uint8_t*>(items + a.tstride);  // [tstride]
  signed char* qbin = reinterpret_cast<signed char*>(blk + a.tstride);  // [qstride] rot bin of a commit
  // initialisation mode: smallest distance a train feature has been claimed with (distances are <= 256; in shared
  // memory because every candidate of every query tests it)
  uint8_t* ioct = reinterpret_cast<uint8_t*>(qbin + a.qstride);          // [tstride] octave of items[k]: the level test of a candidate without its key point
  unsigned short* mdist = a.mode == 2 ? reinterpret_cast<unsigned short*>(smem + (((GRID_CELLS * 2 + 1) * sizeof(int) + (size_t)a.tstride * 4 + a.qstride + 1) & ~(size_t)1)) : nullptr;
  unsigned short* sown = mdist ? mdist + a.tstride : nullptr;   // ... and the query that holds it (copy of owner[])
  __shared__ QRes res[2][SEARCH_CHUNK];
  __shared__ QAux aux[2][SEARCH_CHUNK];
  __shared__ int hist[HISTO_LENGTH];
  __shared__ int s_nm, s_keep[3];
  __shared__ int wtmp[33];

  // ---- AssignFeaturesToGrid: CSR of the 64x48 grid, cell lists in keypoint order
  for (int i = tid; i < GRID_CELLS; i += NT) cursor[i] = 0;
  if (tid < HISTO_LENGTH) hist[tid] = 0;
  if (tid == 0) s_nm = 0;
  __syncthreads();
  for (int i = tid; i < n; i += NT) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(keys[i].x, a.grid.min_x), a.grid.inv_w));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(keys[i].y, a.grid.min_y), a.grid.inv_h));
    if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) atomicAdd(&cursor[px * GRID_ROWS + py], 1);
    owner[i] = -1;
    blk[i] = a.blocked ? a.blocked[(size_t)pair * a.tstride + i] : 0;
    if (mdist) { mdist[i] = 0x7fffu; sown[i] = 0xffffu; }
  }
  for (int i = tid; i < nq; i += NT) { m12[i] = -1; qbin[i] = -1; }
  __syncthreads();
  {  // exclusive scan of the cell counts (chunked, NT threads)
    const int chunk = (GRID_CELLS + NT - 1) / NT;
    const int beg = min(tid * chunk, GRID_CELLS), end = min(beg + chunk, GRID_CELLS);
    int sum = 0;
    for (int i = beg; i < end; i++) sum += cursor[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wtmp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int v = lane < SEARCH_WARPS ? wtmp[lane] : 0;
      int iv = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, iv, o);
        if (lane >= o) iv += t;
      }
      wtmp[lane] = iv - v;
    }
    __syncthreads();
    int base = wtmp[wid] + incl - sum;
    for (int i = beg; i < end; i++) {
      cellStart[i] = base;
      base += cursor[i];
      cursor[i] = 0;
    }
    if (end == GRID_CELLS && beg < end) cellStart[GRID_CELLS] = base;
  }
  __syncthreads();
  for (int i = tid; i < n; i += NT) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(keys[i].x, a.grid.min_x), a.grid.inv_w));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(keys[i].y, a.grid.min_y), a.grid.inv_h));
    if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) {
      const int c = px * GRID_ROWS + py;
      items[cellStart[c] + atomicAdd(&cursor[c], 1)] = (unsigned short)i;
    }
  }
  __syncthreads();
  for (int c = tid; c < GRID_CELLS; c += NT) {  // restore insertion (index) order inside each cell
    const int s = cellStart[c], e = cellStart[c + 1];
    for (int i = s + 1; i < e; i++) {
      const unsigned short v = items[i];
      int j = i - 1;
      while (j >= s && items[j] > v) { items[j + 1] = items[j]; j--; }
      items[j + 1] = v;
    }
  }
  __syncthreads();
  for (int k = tid; k < cellStart[GRID_CELLS]; k += NT) ioct[k] = (uint8_t)min(max(keys[items[k]].octave, 0), 255);
  __syncthreads();

  // ---- queries.  Warps 1 .. 7 evaluate chunk c + 1 (SEARCH_CHUNK queries, two per warp) while warp 0 commits chunk c in
  // order: one barrier per chunk, and the serial commit no longer waits for the evaluation.  An evaluation may be
  // arbitrarily early: a feature, once blocked (or, in the initialisation mode, claimed with a distance), never becomes
  // eligible again, so the set an evaluation saw is a superset of the eligible set at commit time whatever mixture of
  // old and new flags it read, and its two smallest candidates are still the two smallest if both are still eligible --
  // which is what the commit checks (re-evaluating otherwise).
  auto evaluate = [&](int qi, QRes& r, QAux& x) {
    r = QRes{-1, 0x7fffffff, -1, 0x7fffffff};
    x = QAux{1, 0.f, 0.f, -1, -1};
    if (qi < nq) {
      const plvi_query qv = q[qi];
      x.flags = qv.flags;
      x.qangle = qv.angle;
      if (!(qv.flags & 1))
        r = a.mode == 3 ? eval_query_bow(desc, qv, qdesc + (size_t)qi * 32, a.items + (size_t)pair * a.istride, blk)
                        : eval_query(a, keys, desc, qv, qdesc + (size_t)qi * 32, cellStart, items, ioct, blk, mdist, uright, qur ? qur[qi] : 0.f);
    }
    if (r.best >= 0) { x.bangle = keys[r.best].angle; x.l1 = keys[r.best].octave; }
    if (r.second >= 0) x.l2 = keys[r.second].octave;
  };
  const int nchunk = (nq + SEARCH_CHUNK - 1) / SEARCH_CHUNK;
  for (int c = -1; c < nchunk; c++) {
    if (wid > 0 && c + 1 < nchunk) {
      const int base = (c + 1) * SEARCH_CHUNK, buf = (c + 1) & 1;
#pragma unroll 1
      for (int k = 0; k < SEARCH_CHUNK / (SEARCH_WARPS - 1); k++) {
        const int slot = (wid - 1) + (SEARCH_WARPS - 1) * k;
        QRes r; QAux x;
        evaluate(base + slot, r, x);
        if (lane == 0) { res[buf][slot] = r; aux[buf][slot] = x; }
      }
    }
    if (wid == 0 && c >= 0) {
      const int q0 = c * SEARCH_CHUNK, buf = c & 1;
      for (int j = 0; j < SEARCH_CHUNK && q0 + j < nq; j++) {
        const int qj = q0 + j;
        QAux ax = aux[buf][j];
        if (ax.flags & 1) continue;
        QRes rr = res[buf][j];
        // stale if an earlier commit of this chunk removed the best or the second best
        bool stale = false;
        if (a.mode == 2) {
          stale = (rr.best >= 0 && mdist[rr.best] <= rr.bestDist) ||
                  (rr.second >= 0 && mdist[rr.second] <= rr.secondDist);
        } else {
          stale = (rr.best >= 0 && blk[rr.best]) || (a.mode != 0 && rr.second >= 0 && blk[rr.second]);
        }
        if (stale) {
          const plvi_query qq = q[qj];
          rr = a.mode == 3 ? eval_query_bow(desc, qq, qdesc + (size_t)qj * 32, a.items + (size_t)pair * a.istride, blk)
                           : eval_query(a, keys, desc, qq, qdesc + (size_t)qj * 32, cellStart, items, ioct, blk, mdist, uright, qur ? qur[qj] : 0.f);
          ax.l1 = ax.l2 = -1;
          if (rr.best >= 0) { ax.bangle = keys[rr.best].angle; ax.l1 = keys[rr.best].octave; }
          if (rr.second >= 0) ax.l2 = keys[rr.second].octave;
        }
        if (lane == 0 && rr.best >= 0) {
          if (a.mode == 0) {
            if (rr.bestDist <= a.th) {
              owner[rr.best] = qj;
              m12[qj] = rr.best;
              if (!(ax.flags & 2)) blk[rr.best] = 1;
              s_nm++;
              if (a.checkOri) {
                const int b = rot_bin(ax.qangle, ax.bangle);
                qbin[qj] = (signed char)b;
                hist[b]++;
              }
            }
          } else if (a.mode == 1) {
            if (rr.bestDist <= a.th) {
              const int l1 = ax.l1, l2 = rr.second >= 0 ? ax.l2 : -1;
              const int d2 = rr.second >= 0 ? rr.secondDist : 256;
              const bool reject = (l1 == l2) && ((float)rr.bestDist > __fmul_rn(a.nnratio, (float)d2));
              if (!reject && (l1 != l2 || (float)rr.bestDist <= __fmul_rn(a.nnratio, (float)d2))) {
                owner[rr.best] = qj;
                m12[qj] = rr.best;
                if (!(ax.flags & 2)) blk[rr.best] = 1;
                s_nm++;
              }
            }
          } else if (a.mode == 3) {
            const int d2 = rr.second >= 0 ? rr.secondDist : 256;
            if (rr.bestDist <= a.th && (float)rr.bestDist < __fmul_rn(a.nnratio, (float)d2)) {
              owner[rr.best] = qj;
              m12[qj] = rr.best;
              blk[rr.best] = 1;       // vpMapPointMatches[realIdxF] != NULL blocks, whatever the observations
              s_nm++;
              if (a.checkOri) {
                const int b = rot_bin(ax.qangle, ax.bangle);
                qbin[qj] = (signed char)b;
                hist[b]++;
              }
            }
          } else {
            if (rr.bestDist <= a.th &&
                (float)rr.bestDist < __fmul_rn((float)rr.secondDist, a.nnratio)) {
              const int prev = sown[rr.best];   // the query that held this feature (0xffff: none)
              if (prev != 0xffff) { m12[prev] = -1; s_nm--; }
              m12[qj] = rr.best;
              owner[rr.best] = qj;
              sown[rr.best] = (unsigned short)qj;
              mdist[rr.best] = rr.bestDist;
              s_nm++;
              if (a.checkOri) {
                const int b = rot_bin(ax.qangle, ax.bangle);
                qbin[qj] = (signed char)b;
                hist[b]++;
              }
            }
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }

  // ---- rotation consistency (ComputeThreeMaxima)
  if (a.checkOri && a.mode != 1) {
    if (tid == 0) {
      int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
      for (int i = 0; i < HISTO_LENGTH; i++) {
        const int s = hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
        else if (s > max3) { max3 = s; i3 = i; }
      }
      if ((float)max2 < 0.1f * (float)max1) { i2 = -1; i3 = -1; }
      else if ((float)max3 < 0.1f * (float)max1) { i3 = -1; }
      s_keep[0] = i1; s_keep[1] = i2; s_keep[2] = i3;
    }
    __syncthreads();
    int dec = 0;
    for (int i = tid; i < nq; i += NT) {
      const int b = qbin[i];
      if (b < 0 || b == s_keep[0] || b == s_keep[1] || b == s_keep[2]) continue;
      if (a.mode == 0 || a.mode == 3) {
        // the reference nulls mvpMapPoints[bestIdx2] of every culled assignment
        // (check_orientation == 2 reports such a feature as -2, so that a caller holding pointers can tell "assigned,
        // then nulled" from "never touched")
        const int i2 = m12[i];
        owner[i2] = a.checkOri == 2 ? -2 : -1;
        m12[i] = -1;
        dec++;
      } else if (m12[i] >= 0) {
        owner[m12[i]] = -1;
        m12[i] = -1;
        dec++;
      }
    }
    if (dec) atomicSub(&s_nm, dec);
    __syncthreads();
  }
  if (a.mode == 2) {  // "Update prev matched"
    for (int i = tid; i < nq; i += NT)
      if (m12[i] >= 0) { q[i].u = keys[m12[i]].x; q[i].v = keys[m12[i]].y; }
  }
  if (tid == 0) a.nmatches[pair] = s_nm;
}

// Identity-pose stand-in for the host-side projection of SearchByProjection(Frame,Frame)
// (src/ORBmatcher.cc:1992-2023): every keypoint of the query frame "projects" onto its own
// position; radius = th * mvScaleFactors[octave]; levels octave-1 .. octave+1.
__global__ void k_queries_from_keypoints(const plvi_keypoint* __restrict__ kps, const int* __restrict__ counts,
                                         int stride, float th, float scaleFactor, plvi_query* __restrict__ q) {
  const int pair = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= counts[pair]) return;
  const plvi_keypoint k = kps[(size_t)pair * stride + i];
  float sf = 1.f;
  for (int l = 0; l < k.octave; l++) sf = __fmul_rn(sf, scaleFactor);
  plvi_query o;
  o.u = k.x; o.v = k.y;
  o.radius = __fmul_rn(th, sf);
  o.min_level = k.octave - 1;
  o.max_level = k.octave + 1;
  o.angle = k.angle;
  o.flags = 0;
  q[(size_t)pair * stride + i] = o;
}

// C3 pairs stored as consecutive frames (query frame 2p, searched frame 2p + 1) under a known image-to-image affine
// map: the host-side projection of SearchByProjection(CurrentFrame, LastFrame) (src/ORBmatcher.cc:1992-2023) with the
// warp in the place of the pose -- u = a0 x + a1 y + a2, v = a3 x + a4 y + a5, points that land outside the image
// bounds are dropped (:2007-2010), radius = th * mvScaleFactors[octave], levels octave-1 .. octave+1 -- and the
// query set of SearchForInitialization (src/ORBmatcher.cc:717-727: level-0 keypoints, window around their own
// position).  Also compacts the per-frame counts into per-pair arrays.
__global__ void k_pair_queries(const plvi_keypoint* __restrict__ kps, const int* __restrict__ counts, int stride, int ostride, float th,
                               float scaleFactor, float a0, float a1, float a2, float a3, float a4, float a5, float minX,
                               float maxX, float minY, float maxY, float initWindow, plvi_query* __restrict__ qProj,
                               plvi_query* __restrict__ qInit, int* __restrict__ qcount, int* __restrict__ tcount) {
  const int pair = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(counts[2 * pair], stride);
  if (i == 0) { qcount[pair] = n; tcount[pair] = min(counts[2 * pair + 1], stride); }
  if (i >= n) return;
  const plvi_keypoint k = kps[(size_t)(2 * pair) * stride + i];
  float sf = 1.f;
  for (int l = 0; l < k.octave; l++) sf = __fmul_rn(sf, scaleFactor);
  plvi_query o;
  o.u = __fadd_rn(__fadd_rn(__fmul_rn(a0, k.x), __fmul_rn(a1, k.y)), a2);
  o.v = __fadd_rn(__fadd_rn(__fmul_rn(a3, k.x), __fmul_rn(a4, k.y)), a5);
  o.radius = __fmul_rn(th, sf);
  o.min_level = k.octave - 1;
  o.max_level = k.octave + 1;
  o.angle = k.angle;
  o.flags = (o.u < minX || o.u > maxX || o.v < minY || o.v > maxY) ? 1 : 0;
  qProj[(size_t)pair * ostride + i] = o;
  if (qInit) {
    o.u = k.x; o.v = k.y;
    o.radius = initWindow;
    o.min_level = o.max_level = k.octave;
    o.flags = k.octave > 0 ? 1 : 0;
    qInit[(size_t)pair * ostride + i] = o;
  }
}

__global__ void k_gather_i32(const int* __restrict__ src, int n, int first, int step, int* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[first + i * step];
}

// ---- the disparity / overlap / depth filter after the stereo line search: Frame::ComputeStereoMatches_Lines
// (src/Frame.cc:1453-1500), lineSegmentOverlapStereo (:1502-1533), filterLineSegmentDisparity (:1535-1546).  Thread per left
// line, CTA per stereo pair; doubles with explicit round-to-nearest operations (the reference's Eigen expressions
// without contraction); std::min / std::max as the reference's comparisons (NaN behaviour included).
__device__ __forceinline__ double std_min_d(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double std_max_d(double a, double b) { return (a < b) ? b : a; }

__device__ double line_overlap_stereo(double spl_obs, double epl_obs, double spl_proj, double epl_proj) {
  double overlap = 1.0;
  const double lineHorizTh = (double)0.1f;
  if (fabs(__dsub_rn(epl_obs, spl_obs)) > lineHorizTh) {
    const double sln = std_min_d(spl_obs, epl_obs), eln = std_max_d(spl_obs, epl_obs);
    const double spn = std_min_d(spl_proj, epl_proj), epn = std_max_d(spl_proj, epl_proj);
    const double length = __dsub_rn(eln, spn);
    if ((epn < sln) || (spn > eln)) overlap = 0.0;
    else if ((epn > eln) && (spn < sln)) overlap = __dsub_rn(eln, sln);
    else overlap = __dsub_rn(std_min_d(eln, epn), std_max_d(sln, spn));
    if (length > (double)0.01f) overlap = __ddiv_rn(overlap, length);
    else overlap = 0.0;
    if (overlap > 1.0) overlap = 1.0;
  }
  return overlap;
}

__global__ void __launch_bounds__(128) k_line_stereo_depth(const float4* __restrict__ seg1, const int* __restrict__ n1, int stride1,
                                                           const float4* __restrict__ seg2, const int* __restrict__ n2, int stride2,
                                                           const int* __restrict__ matches12,
                                                           const float4* __restrict__ seg1un, float mbf, float2* __restrict__ disparity,
                                                           float2* __restrict__ depth, double* __restrict__ le, int* __restrict__ ndepth) {
  const int pair = blockIdx.x;
  const int n = min(n1[pair], stride1), nr = min(n2[pair], stride2);
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int mine = 0;
  for (int i1 = threadIdx.x; i1 < n; i1 += blockDim.x) {
    const size_t o = (size_t)pair * stride1 + i1;
    float2 dsp = make_float2(-1.f, -1.f), dep = make_float2(-1.f, -1.f);
    if (le && nr == 0) {   // the reference returns before mvle_l is filled (src/Frame.cc:1419-1420)
      le[3 * o] = 0.0; le[3 * o + 1] = 0.0; le[3 * o + 2] = 0.0;
    } else if (le) {   // mvle_l: the normalised image line through the undistorted end points
      const float4 u = seg1un[o];
      const double a0 = u.x, a1 = u.y, b0 = u.z, b1 = u.w;
      const double c0 = __dsub_rn(__dmul_rn(a1, 1.0), __dmul_rn(1.0, b1)), c1 = __dsub_rn(__dmul_rn(1.0, b0), __dmul_rn(a0, 1.0));
      const double c2 = __dsub_rn(__dmul_rn(a0, b1), __dmul_rn(a1, b0));
      const double nrm = __dsqrt_rn(__dadd_rn(__dmul_rn(c0, c0), __dmul_rn(c1, c1)));
      le[3 * o] = __ddiv_rn(c0, nrm); le[3 * o + 1] = __ddiv_rn(c1, nrm); le[3 * o + 2] = __ddiv_rn(c2, nrm);
    }
    const int i2 = matches12[o];
    if (i2 >= 0 && i2 < nr) {
      const float4 L = seg1[o], R = seg2[(size_t)pair * stride2 + i2];
      const double xl1 = L.x, yl1 = L.y, xl2 = L.z, yl2 = L.w;
      double xr1 = R.x, yr1 = R.y, xr2 = R.z, yr2 = R.w;
      const double overlap = line_overlap_stereo(yl1, yl2, yr1, yr2);
      // the comma initialisers overwrite sp_r first: the second expression already reads the new sp_r (:1468-1469)
      xr1 = __ddiv_rn(__dadd_rn(__dmul_rn(xr1, __dsub_rn(yl1, yr2)), __dmul_rn(xr2, __dsub_rn(yr1, yl1))), __dsub_rn(yr1, yr2));
      yr1 = yl1;
      xr2 = __ddiv_rn(__dadd_rn(__dmul_rn(xr1, __dsub_rn(yl2, yr2)), __dmul_rn(xr2, __dsub_rn(yr1, yl2))), __dsub_rn(yr1, yr2));
      yr2 = yl2;
      double ds = __dsub_rn(xl1, xr1), de = __dsub_rn(xl2, xr2);
      if (__ddiv_rn(std_min_d(ds, de), std_max_d(ds, de)) < (double)0.7f) { ds = -1.0; de = -1.0; }
      if (ds >= 1.0 && de >= 1.0 && fabs(__dsub_rn(yl1, yl2)) > (double)0.1f && fabs(__dsub_rn(yr1, yr2)) > (double)0.1f &&
          overlap > (double)0.75f) {
        dsp = make_float2((float)ds, (float)de);
        dep = make_float2(__fdiv_rn(mbf, (float)ds), __fdiv_rn(mbf, (float)de));
        mine++;
      }
    }
    disparity[o] = dsp;
    depth[o] = dep;
  }
  if (mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0) ndepth[pair] = s_cnt;
}


// ---- lines: knn-2 + ratio in both directions + mutual check, one CTA per pair ---------
__device__ void nnr_rows(const uint8_t* sa, int na, const uint8_t* sb, int nb, float nnr, int* out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = wid; i < na; i += nw) {
    uint32_t qw[8];
#pragma unroll
    for (int k = 0; k < 8; k++) qw[k] = reinterpret_cast<const uint32_t*>(sa + i * 32)[k];
    Top2 t = {0xffffffffu, 0xffffffffu, -1, -1};
    for (int j = lane; j < nb; j += 32) {
      const uint32_t* pb = reinterpret_cast<const uint32_t*>(sb + j * 32);
      int d = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) d += __popc(qw[k] ^ pb[k]);
      top2_insert(t, ((uint32_t)d << 16) | (uint32_t)j, j);
    }
    top2_warp_merge(t);
    if (lane == 0) {
      int m = -1;
      if (nb >= 2 && (float)(t.k0 >> 16) < __fmul_rn((float)(t.k1 >> 16), nnr)) m = t.i0;
      out[i] = m;
    }
  }
}

__global__ void __launch_bounds__(256) k_line_match(const uint8_t* __restrict__ d1, const int* __restrict__ n1p,
                                                    int stride1, const uint8_t* __restrict__ d2,
                                                    const int* __restrict__ n2p, int stride2, float nnr,
                                                    int mutual, int* __restrict__ m12out,
                                                    int* __restrict__ nmatches) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int pair = blockIdx.x, tid = threadIdx.x;
  const int n1 = min(n1p[pair], stride1), n2 = min(n2p[pair], stride2);
  uint8_t* s1 = smem;
  uint8_t* s2 = s1 + (size_t)stride1 * 32;
  int* m12 = reinterpret_cast<int*>(s2 + (size_t)stride2 * 32);
  int* m21 = m12 + stride1;
  __shared__ int s_cnt;
  if (tid == 0) s_cnt = 0;
  const uint4* g1 = reinterpret_cast<const uint4*>(d1 + (size_t)pair * stride1 * 32);
  const uint4* g2 = reinterpret_cast<const uint4*>(d2 + (size_t)pair * stride2 * 32);
  for (int i = tid; i < n1 * 2; i += blockDim.x) reinterpret_cast<uint4*>(s1)[i] = __ldg(g1 + i);
  for (int i = tid; i < n2 * 2; i += blockDim.x) reinterpret_cast<uint4*>(s2)[i] = __ldg(g2 + i);
  __syncthreads();
  nnr_rows(s1, n1, s2, n2, nnr, m12);
  if (mutual) nnr_rows(s2, n2, s1, n1, nnr, m21);
  __syncthreads();
  int cnt = 0;
  for (int i = tid; i < n1; i += blockDim.x) {
    int m = m12[i];
    if (mutual && m >= 0 && m21[m] != i) m = -1;
    m12out[(size_t)pair * stride1 + i] = m;
    cnt += m >= 0;
  }
  if (cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (tid == 0) nmatches[pair] = s_cnt;
}

}  // namespace plvi

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
using namespace plvi;

// ---------------------------------------------------------------------------------
// k_line_match_mad: LineMatcher::SerachForInitialize / SearchForTriangulation(KF, KF)
// (src/LineMatcher.cpp:113-171): kNN-2 of every row of desc1 in desc2, then the robust
// threshold of Frame/KeyFrame::lineDescriptorMAD (src/Frame.cc:1089-1113):
//   nn_mad   = 1.4826 * median(|d0 - median(d0)|)
//   nn12_mad = 1.4826 * median(|(d1 - d0) - median(d1 - d0)|),  "median" = element n/2 of the sorted list
// a query is matched to its nearest neighbour iff d1 - d0 > factor * nn12_mad.  All distances
// are integers in [0, 256], so the medians come from 257-bin histograms.
// ---------------------------------------------------------------------------------
__device__ int hist_select(const int* hist, int rank) {   // smallest v with #{x <= v} > rank
  int cum = 0;
  for (int v = 0; v <= 256; v++) {
    cum += hist[v];
    if (cum > rank) return v;
  }
  return 256;
}

__global__ void __launch_bounds__(256) k_line_match_mad(const uint8_t* __restrict__ d1, const int* __restrict__ n1p, int stride1,
                                                        const uint8_t* __restrict__ d2, const int* __restrict__ n2p, int stride2,
                                                        const uint8_t* __restrict__ mask1, const uint8_t* __restrict__ mask2,
                                                        double factor, int* __restrict__ m12out, int* __restrict__ nmatches,
                                                        double* __restrict__ madOut) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int n1 = min(n1p[pair], stride1), n2 = min(n2p[pair], stride2);
  uint8_t* s1 = smem;
  uint8_t* s2 = s1 + (size_t)stride1 * 32;
  int* dist0 = reinterpret_cast<int*>(s2 + (size_t)stride2 * 32);   // [stride1]
  int* dist1 = dist0 + stride1;
  int* train = dist1 + stride1;
  __shared__ int hist[257];
  __shared__ int s_med, s_cnt;
  __shared__ double s_mad[2];
  int* out = m12out + (size_t)pair * stride1;
  if (n1 < 1 || n2 < 2) {   // knnMatch(k = 2) needs two train rows (the reference reads lmatches[i][1] unconditionally)
    for (int i = tid; i < n1; i += blockDim.x) out[i] = -1;
    if (tid == 0) { nmatches[pair] = 0; madOut[2 * pair] = 0.0; madOut[2 * pair + 1] = 0.0; }
    return;
  }
  const uint4* g1 = reinterpret_cast<const uint4*>(d1 + (size_t)pair * stride1 * 32);
  const uint4* g2 = reinterpret_cast<const uint4*>(d2 + (size_t)pair * stride2 * 32);
  for (int i = tid; i < n1 * 2; i += blockDim.x) reinterpret_cast<uint4*>(s1)[i] = __ldg(g1 + i);
  for (int i = tid; i < n2 * 2; i += blockDim.x) reinterpret_cast<uint4*>(s2)[i] = __ldg(g2 + i);
  __syncthreads();
  for (int i = wid; i < n1; i += nw) {
    uint32_t qw[8];
#pragma unroll
    for (int k = 0; k < 8; k++) qw[k] = reinterpret_cast<const uint32_t*>(s1 + i * 32)[k];
    Top2 t = {0xffffffffu, 0xffffffffu, -1, -1};
    for (int j = lane; j < n2; j += 32) {
      const uint32_t* pb = reinterpret_cast<const uint32_t*>(s2 + j * 32);
      int d = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) d += __popc(qw[k] ^ pb[k]);
      top2_insert(t, ((uint32_t)d << 16) | (uint32_t)j, j);
    }
    top2_warp_merge(t);
    if (lane == 0) { dist0[i] = (int)(t.k0 >> 16); dist1[i] = (int)(t.k1 >> 16); train[i] = t.i0; }
  }
  __syncthreads();
  const int mid = n1 / 2;
  // pass 0: d0, pass 1: d1 - d0
  for (int pass = 0; pass < 2; pass++) {
    for (int v = tid; v < 257; v += blockDim.x) hist[v] = 0;
    __syncthreads();
    for (int i = tid; i < n1; i += blockDim.x) atomicAdd(&hist[pass == 0 ? dist0[i] : dist1[i] - dist0[i]], 1);
    __syncthreads();
    // d0 is sorted ascending, d1 - d0 DESCENDING (conpare_descriptor_by_NN12_dist, include/LineMatcher.h:63-68)
    if (tid == 0) s_med = hist_select(hist, pass == 0 ? mid : n1 - 1 - mid);
    __syncthreads();
    const int med = s_med;
    for (int v = tid; v < 257; v += blockDim.x) hist[v] = 0;
    __syncthreads();
    for (int i = tid; i < n1; i += blockDim.x) atomicAdd(&hist[abs((pass == 0 ? dist0[i] : dist1[i] - dist0[i]) - med)], 1);
    __syncthreads();
    if (tid == 0) s_mad[pass] = __dmul_rn(1.4826, (double)(float)hist_select(hist, mid));
    __syncthreads();
  }
  const double th = __dmul_rn(s_mad[1], factor);
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  int cnt = 0;
  for (int i = tid; i < n1; i += blockDim.x) {
    int m = train[i];
    if ((mask1 && mask1[(size_t)pair * stride1 + i]) || (mask2 && mask2[(size_t)pair * stride2 + m])) m = -1;
    else if (!((double)(dist1[i] - dist0[i]) > th)) m = -1;
    out[i] = m;
    cnt += m >= 0;
  }
  if (cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (tid == 0) { nmatches[pair] = s_cnt; madOut[2 * pair] = s_mad[0]; madOut[2 * pair + 1] = s_mad[1]; }
}

// ---------------------------------------------------------------------------------
// k_distinctive: MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:330-402): of the N observed
// descriptors of a map point keep the one with the least median distance to the others; median =
// element int(0.5 * (N - 1)) of the sorted row (self distance 0 included), first index wins ties.
// One warp per map point: lane = row, 257-bin u16 histogram per lane in shared memory.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_distinctive(const uint8_t* __restrict__ desc, const int* __restrict__ counts, int stride,
                                                    int* __restrict__ bestIdx, uint8_t* __restrict__ bestDesc) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int mp = blockIdx.x, lane = threadIdx.x;
  const int n = min(counts[mp], stride);
  uint8_t* sd = smem;                                                   // [stride][32]
  unsigned short* hist = reinterpret_cast<unsigned short*>(sd + (size_t)stride * 32) + lane * 258;   // [32][258]
  if (n <= 0) {
    if (lane == 0) bestIdx[mp] = -1;
    return;
  }
  const uint4* g = reinterpret_cast<const uint4*>(desc + (size_t)mp * stride * 32);
  for (int i = lane; i < n * 2; i += 32) reinterpret_cast<uint4*>(sd)[i] = __ldg(g + i);
  __syncwarp();
  const int mid = (int)(0.5 * (n - 1));
  unsigned bestKey = 0xffffffffu;    // median << 16 | row
  for (int i = lane; i < n; i += 32) {
    for (int v = 0; v < 257; v++) hist[v] = 0;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(sd + i * 32);
    for (int j = 0; j < n; j++) {
      const uint32_t* bq = reinterpret_cast<const uint32_t*>(sd + j * 32);
      int d = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) d += __popc(a[k] ^ bq[k]);
      hist[d]++;
    }
    int cum = 0, med = 256;
    for (int v = 0; v <= 256; v++) {
      cum += hist[v];
      if (cum > mid) { med = v; break; }
    }
    bestKey = min(bestKey, ((unsigned)med << 16) | (unsigned)i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bestKey = min(bestKey, __shfl_xor_sync(0xffffffffu, bestKey, o));
  const int b = (int)(bestKey & 0xffffu);
  if (lane == 0) bestIdx[mp] = b;
  if (bestDesc && lane < 8) reinterpret_cast<uint32_t*>(bestDesc + (size_t)mp * 32)[lane] = reinterpret_cast<const uint32_t*>(sd + b * 32)[lane];
}

// ---------------------------------------------------------------------------------
// k_search_triangulation: descriptor + epipolar part of ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12,
// vMatchedPairs, bOnlyStereo = false, bCoarse) (src/ORBmatcher.cc:965-1206), monocular pinhole path with
// Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:135-157).  Nothing is claimed in this function
// (vbMatched2 is never set), so the queries are independent: one warp per query, lanes over the
// candidates of the vocabulary node.  The reference's scan keeps the candidate with the smallest
// distance among those passing the geometric tests, the LAST one on ties (dist > bestDist is skipped,
// equality replaces).  float arithmetic without contraction.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_search_triangulation(const plvi_keypoint* __restrict__ keysAll, const uint8_t* __restrict__ descAll,
                                                              const uint8_t* __restrict__ blockedAll, const int* __restrict__ tcount,
                                                              int tstride, const int* __restrict__ itemsAll, int istride,
                                                              const plvi_query* __restrict__ qAll, const uint8_t* __restrict__ qdescAll,
                                                              const int* __restrict__ qcount, int qstride,
                                                              const plvi_epipolar* __restrict__ geomAll, int thLow, int checkOri,
                                                              int* __restrict__ m12All, int* __restrict__ nmatches) {
  __shared__ int hist[HISTO_LENGTH];
  __shared__ int s_keep[3], s_nm;
  __shared__ plvi_epipolar G;
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nq = min(qcount[pair], qstride);
  const plvi_keypoint* keys = keysAll + (size_t)pair * tstride;
  const uint8_t* desc = descAll + (size_t)pair * tstride * 32;
  const uint8_t* blk = blockedAll ? blockedAll + (size_t)pair * tstride : nullptr;
  const int* items = itemsAll + (size_t)pair * istride;
  const plvi_query* q = qAll + (size_t)pair * qstride;
  const uint8_t* qdesc = qdescAll + (size_t)pair * qstride * 32;
  int* m12 = m12All + (size_t)pair * qstride;
  (void)tcount;
  if (tid < HISTO_LENGTH) hist[tid] = 0;
  if (tid == 0) { s_nm = 0; G = geomAll[pair]; }
  __syncthreads();
  for (int qi = wid; qi < nq; qi += 8) {
    const plvi_query Q = q[qi];
    unsigned best = 0xffffffffu;
    if (!(Q.flags & 1)) {
      uint32_t qw[8];
#pragma unroll
      for (int k = 0; k < 8; k++) qw[k] = __ldg(reinterpret_cast<const uint32_t*>(qdesc + (size_t)qi * 32) + k);
      // epipolar line of kp1 in image 2: l = x1' F12 = [a b c]
      const float a = __fadd_rn(__fadd_rn(__fmul_rn(Q.u, G.F12[0]), __fmul_rn(Q.v, G.F12[3])), G.F12[6]);
      const float b = __fadd_rn(__fadd_rn(__fmul_rn(Q.u, G.F12[1]), __fmul_rn(Q.v, G.F12[4])), G.F12[7]);
      const float c = __fadd_rn(__fadd_rn(__fmul_rn(Q.u, G.F12[2]), __fmul_rn(Q.v, G.F12[5])), G.F12[8]);
      const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
      for (int k = Q.min_level + lane; k < Q.max_level; k += 32) {
        const int i2 = __ldg(items + k);
        const unsigned bv = blk ? blk[i2] : 0u;   // bit0: has a map point; bit1: bStereo2 (mvuRight[idx2] >= 0)
        if (bv & 1u) continue;
        const int d = hamming256_regs(qw, desc + (size_t)i2 * 32);
        if (d > thLow) continue;
        const plvi_keypoint kp2 = keys[i2];
        if (G.check_epipole && !(Q.flags & 4) && !(bv & 2u)) {   // only when neither feature is a stereo observation
          const float ex = __fsub_rn(G.ep_x, kp2.x), ey = __fsub_rn(G.ep_y, kp2.y);
          if (__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)) < __fmul_rn(100.f, G.scale_factors[kp2.octave])) continue;
        }
        if (!G.coarse) {
          if (den == 0.f) continue;
          const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, kp2.x), __fmul_rn(b, kp2.y)), c);
          const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
          if (!((double)dsqr < __dmul_rn(3.84, (double)G.level_sigma2[kp2.octave]))) continue;
        }
        // smallest distance, last candidate on ties
        best = min(best, ((unsigned)d << 20) | (unsigned)(0xfffff - min(k - Q.min_level, 0xfffff)));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) {
      int m = -1;
      if (best != 0xffffffffu) {
        m = items[Q.min_level + (0xfffff - (int)(best & 0xfffffu))];
        atomicAdd(&s_nm, 1);
        if (checkOri) atomicAdd(&hist[rot_bin(Q.angle, keys[m].angle)], 1);
      }
      m12[qi] = m;
    }
  }
  __syncthreads();
  if (checkOri) {
    if (tid == 0) {
      int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
      for (int i = 0; i < HISTO_LENGTH; i++) {
        const int sN = hist[i];
        if (sN > max1) { max3 = max2; max2 = max1; max1 = sN; i3 = i2; i2 = i1; i1 = i; }
        else if (sN > max2) { max3 = max2; max2 = sN; i3 = i2; i2 = i; }
        else if (sN > max3) { max3 = sN; i3 = i; }
      }
      if ((float)max2 < 0.1f * (float)max1) { i2 = -1; i3 = -1; }
      else if ((float)max3 < 0.1f * (float)max1) { i3 = -1; }
      s_keep[0] = i1; s_keep[1] = i2; s_keep[2] = i3;
    }
    __syncthreads();
    int dec = 0;
    for (int i = tid; i < nq; i += 256) {
      const int m = m12[i];
      if (m < 0) continue;
      const int bb = rot_bin(q[i].angle, keys[m].angle);
      if (bb == s_keep[0] || bb == s_keep[1] || bb == s_keep[2]) continue;
      m12[i] = -1;
      dec++;
    }
    if (dec) atomicSub(&s_nm, dec);
    __syncthreads();
  }
  if (tid == 0) nmatches[pair] = s_nm;
}

// ---------------------------------------------------------------------------------------
// k_search_in_radius: the per-map-point search of ORBmatcher::Fuse (both overloads), SearchBySim3 (both directions)
// and SearchByProjection(KeyFrame*, Scw, ...) (src/ORBmatcher.cc:1399-1610, 1612-1734, 1736-1960, 473-596).
// Queries never claim anything, so they are independent: CTA per keyframe, the 64x48 grid as CSR in shared memory
// (KeyFrame::AssignFeaturesToGrid order), one warp per query, lanes over the cells of GetFeaturesInArea.  The best
// candidate is the minimum of (distance, cell position, position inside the cell) = the first smallest distance in
// the reference's iteration order.
// ---------------------------------------------------------------------------------------
struct RadiusArgs {
  const plvi_keypoint* keys; const uint8_t* desc; const int* tcount; int tstride;
  plvi_grid grid;
  const plvi_query* q; const uint8_t* qdesc; const int* qcount; int qstride;
  float invSigma2[16];
  double chi2;
  int th;
  int* bestIdx; int* bestDist; int* nfound;
  const float* uright; const float* qur;   // plvi_matcher_set_stereo: mvuRight of the keyframe, ur of every query (or nullptr)
};

__global__ void __launch_bounds__(SEARCH_WARPS * 32) k_search_in_radius(const RadiusArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int NT = SEARCH_WARPS * 32;
  const int n = min(a.tcount[pair], a.tstride), nq = min(a.qcount[pair], a.qstride);
  const float* uright = (a.uright && a.qur) ? a.uright + (size_t)pair * a.tstride : nullptr;
  const float* qurAll = uright ? a.qur + (size_t)pair * a.qstride : nullptr;
  const plvi_keypoint* keys = a.keys + (size_t)pair * a.tstride;
  const uint8_t* desc = a.desc + (size_t)pair * a.tstride * 32;
  const plvi_query* q = a.q + (size_t)pair * a.qstride;
  const uint8_t* qdesc = a.qdesc + (size_t)pair * a.qstride * 32;
  int* cellStart = reinterpret_cast<int*>(smem);                 // [GRID_CELLS + 1]
  int* cursor = cellStart + GRID_CELLS + 1;                      // [GRID_CELLS]
  unsigned short* items = reinterpret_cast<unsigned short*>(cursor + GRID_CELLS);  // [tstride]
  __shared__ int wtmp[33];
  __shared__ int s_found;

  // ---- AssignFeaturesToGrid: CSR of the grid, cell lists in keypoint order
  for (int i = tid; i < GRID_CELLS; i += NT) cursor[i] = 0;
  if (tid == 0) s_found = 0;
  __syncthreads();
  for (int i = tid; i < n; i += NT) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(keys[i].x, a.grid.min_x), a.grid.inv_w));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(keys[i].y, a.grid.min_y), a.grid.inv_h));
    if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) atomicAdd(&cursor[px * GRID_ROWS + py], 1);
  }
  __syncthreads();
  {
    const int chunk = (GRID_CELLS + NT - 1) / NT;
    const int beg = min(tid * chunk, GRID_CELLS), end = min(beg + chunk, GRID_CELLS);
    int sum = 0;
    for (int i = beg; i < end; i++) sum += cursor[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wtmp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const int v = lane < SEARCH_WARPS ? wtmp[lane] : 0;
      int iv = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, iv, o);
        if (lane >= o) iv += t;
      }
      wtmp[lane] = iv - v;
    }
    __syncthreads();
    int base = wtmp[wid] + incl - sum;
    for (int i = beg; i < end; i++) {
      cellStart[i] = base;
      base += cursor[i];
      cursor[i] = 0;
    }
    if (end == GRID_CELLS && beg < end) cellStart[GRID_CELLS] = base;
  }
  __syncthreads();
  for (int i = tid; i < n; i += NT) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(keys[i].x, a.grid.min_x), a.grid.inv_w));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(keys[i].y, a.grid.min_y), a.grid.inv_h));
    if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) {
      const int c = px * GRID_ROWS + py;
      items[cellStart[c] + atomicAdd(&cursor[c], 1)] = (unsigned short)i;
    }
  }
  __syncthreads();
  for (int c = tid; c < GRID_CELLS; c += NT) {  // insertion (index) order inside each cell
    const int s = cellStart[c], e = cellStart[c + 1];
    for (int i = s + 1; i < e; i++) {
      const unsigned short v = items[i];
      int j = i - 1;
      while (j >= s && items[j] > v) { items[j + 1] = items[j]; j--; }
      items[j + 1] = v;
    }
  }
  __syncthreads();

  const plvi_grid& g = a.grid;
  int found = 0;
  for (int qi = wid; qi < nq; qi += SEARCH_WARPS) {
    const plvi_query qq = q[qi];
    int best = -1, bestDist = 256;
    if (!(qq.flags & 1)) {
      const float x = qq.u, y = qq.v, rad = qq.radius;
      const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, g.min_x), rad), g.inv_w)));
      const int cx1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, g.min_x), rad), g.inv_w)));
      const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, g.min_y), rad), g.inv_h)));
      const int cy1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, g.min_y), rad), g.inv_h)));
      if (!(cx0 >= GRID_COLS || cx1 < 0 || cy0 >= GRID_ROWS || cy1 < 0)) {
        uint32_t qw[8];
#pragma unroll
        for (int k = 0; k < 8; k++) qw[k] = __ldg(reinterpret_cast<const uint32_t*>(qdesc + (size_t)qi * 32) + k);
        const int ny = cy1 - cy0 + 1, nc = (cx1 - cx0 + 1) * ny;
        uint32_t bk = 0xffffffffu;
        int bi = -1;
        for (int c = lane; c < nc; c += 32) {
          const int cell = (cx0 + c / ny) * GRID_ROWS + cy0 + c % ny;
          const int s = cellStart[cell], e = cellStart[cell + 1];
          for (int k = s; k < e; k++) {
            const int i2 = items[k];
            const plvi_keypoint kp = keys[i2];
            if (!(fabsf(__fsub_rn(kp.x, x)) < rad && fabsf(__fsub_rn(kp.y, y)) < rad)) continue;
            if (kp.octave < qq.min_level || kp.octave > qq.max_level) continue;
            if (a.chi2 > 0) {   // reprojection gate of Fuse: float product compared with the double constant
              const float ex = __fsub_rn(x, kp.x), ey = __fsub_rn(y, kp.y);
              const float ur2 = uright ? __ldg(uright + i2) : -1.f;
              if (ur2 >= 0.f) {   // stereo observation: 3 degrees of freedom, chi2 = 7.8 (src/ORBmatcher.cc:1530-1543)
                const float er = __fsub_rn(qurAll[qi], ur2);
                const float e2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(er, er));
                if ((double)__fmul_rn(e2, a.invSigma2[kp.octave & 15]) > 7.8) continue;
              } else {
                const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                if ((double)__fmul_rn(e2, a.invSigma2[kp.octave & 15]) > a.chi2) continue;
              }
            }
            const int d = hamming256_regs(qw, desc + (size_t)i2 * 32);
            const uint32_t key = ((uint32_t)d << 23) | ((uint32_t)min(c, 4095) << 11) | (uint32_t)min(k - s, 2047);
            if (key < bk) { bk = key; bi = i2; }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ok < bk) { bk = ok; bi = oi; }
        }
        if (bi >= 0) { best = bi; bestDist = (int)(bk >> 23); }
      }
    }
    if (lane == 0) {
      const bool ok = best >= 0 && bestDist <= a.th;
      a.bestIdx[(size_t)pair * a.qstride + qi] = ok ? best : -1;
      a.bestDist[(size_t)pair * a.qstride + qi] = bestDist;
      found += ok;
    }
  }
  if (lane == 0 && found) atomicAdd(&s_found, found);
  __syncthreads();
  if (tid == 0) a.nfound[pair] = s_found;
}

// ---------------------------------------------------------------------------------------
// k_line_fuse_search: the per-map-line search of LineMatcher::Fuse (src/LineMatcher.cpp:373-485) over
// KeyFrame::GetLinesInArea (src/KeyFrame.cc:1170-1198).  Warp per query, lanes over the keyframe's keylines; the
// mixed float / double expressions of the reference are evaluated operation by operation; distance = the >> 25
// variant of LineMatcher::DescriptorDistance; min over (distance, index) = first smallest distance.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_line_fuse_search(const plvi_keyline* __restrict__ klAll, const uint8_t* __restrict__ descAll,
                                                          const int* __restrict__ ncount, int stride, const float* __restrict__ qAll,
                                                          const uint8_t* __restrict__ flagsAll, const uint8_t* __restrict__ qdescAll,
                                                          const int* __restrict__ qcount, int qstride, int thLow,
                                                          int* __restrict__ bestIdx, int* __restrict__ bestDist, int* __restrict__ nfound) {
  const int pair = blockIdx.y, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int n = ncount[pair], nq = qcount[pair];
  if (qi >= qstride) return;
  const size_t qo = (size_t)pair * qstride + qi;
  int outIdx = -1, outDist = 0x7fffffff;
  if (qi < nq && !(flagsAll && flagsAll[qo])) {
    const float* q = qAll + qo * 6;
    const float x1 = q[0], y1 = q[1], x2 = q[2], y2 = q[3], r = q[4];
    const int level = (int)q[5];
    const double mx = __dmul_rn(0.5, (double)__fadd_rn(x1, x2)), my = __dmul_rn(0.5, (double)__fadd_rn(y1, y2));
    const float r2 = __fmul_rn(r, r);
    const float slope0 = __fdiv_rn(__fsub_rn(y1, y2), __fsub_rn(x1, x2));
    const double slopeTh = __dmul_rn((double)r, 0.01);
    uint32_t qw[8];
#pragma unroll
    for (int k = 0; k < 8; k++) qw[k] = __ldg(reinterpret_cast<const uint32_t*>(qdescAll + qo * 32) + k);
    const plvi_keyline* kl = klAll + (size_t)pair * stride;
    const uint8_t* desc = descAll + (size_t)pair * stride * 32;
    unsigned long long best = ~0ull;
    for (int k = lane; k < n; k += 32) {
      const plvi_keyline L = kl[k];
      const double dx = __dsub_rn(mx, (double)L.pt_x), dy = __dsub_rn(my, (double)L.pt_y);
      const float distance = (float)__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
      if (distance > r2) continue;
      const float slope = __fsub_rn(slope0, L.angle);
      if ((double)slope > slopeTh) continue;
      if (L.octave < level - 1 || L.octave > level) continue;
      const uint32_t* d = reinterpret_cast<const uint32_t*>(desc + (size_t)k * 32);
      int dist = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) dist += __popc(qw[w] ^ __ldg(d + w)) >> 1;   // (popcount * 0x1010101 >> 24) >> 1
      const unsigned long long key = ((unsigned long long)(unsigned)dist << 32) | (unsigned)k;
      best = best < key ? best : key;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, s);
      best = best < o ? best : o;
    }
    if (best != ~0ull) { outDist = (int)(best >> 32); if (outDist <= thLow) outIdx = (int)(best & 0xffffffffu); }
  }
  if (lane == 0) {
    bestIdx[qo] = outIdx;
    bestDist[qo] = outDist;
    if (outIdx >= 0) atomicAdd(&nfound[pair], 1);
  }
}

struct plvi_matcher {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  int maxPairs = 0, maxTrain = 0, maxQuery = 0;
  // device staging for the host-pointer entry points
  plvi_keypoint* dKeys = nullptr; uint8_t* dDesc = nullptr; uint8_t* dBlocked = nullptr;
  int* dTCount = nullptr; plvi_query* dQ = nullptr; uint8_t* dQDesc = nullptr; int* dQCount = nullptr;
  int* dMatchTrain = nullptr; int* dMatchQuery = nullptr; int* dNMatches = nullptr; int* dMatchedDist = nullptr;
  bool staged = false;            // the staging buffers above exist (allocated by the first host-pointer call)
  // plvi_matcher_set_stereo: side information of the next search (device pointers; host arrays are staged below)
  const float* stereoTrain = nullptr; const float* stereoQuery = nullptr;
  float* dStereoTrain = nullptr; float* dStereoQuery = nullptr;
  int lastLaunches = 0;
};

// Device staging of the host-pointer entry points: ~133 B x max_pairs x max_train, only allocated when such an entry
// point is first used (a matcher driven through device pointers never touches it); dMatchedDist (search mode 2) is
// allocated with the handle.
static int ensure_staging(plvi_matcher* m) {
  if (m->staged) return PLVI_OK;
  const size_t P = m->maxPairs, T = m->maxTrain, Q = m->maxQuery;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  A((void**)&m->dKeys, P * T * sizeof(plvi_keypoint));
  A((void**)&m->dDesc, P * (T > Q ? T : Q) * 32);
  A((void**)&m->dBlocked, P * T);
  A((void**)&m->dTCount, P * sizeof(int));
  A((void**)&m->dQ, P * Q * sizeof(plvi_query));
  A((void**)&m->dQDesc, P * (T > Q ? T : Q) * 32);
  A((void**)&m->dQCount, P * sizeof(int));
  A((void**)&m->dMatchTrain, P * T * sizeof(int));
  A((void**)&m->dMatchQuery, P * (T > Q ? T : Q) * sizeof(int));
  A((void**)&m->dNMatches, P * sizeof(int));
  A((void**)&m->dStereoTrain, P * T * sizeof(float));
  A((void**)&m->dStereoQuery, P * Q * sizeof(float));
  if (e != cudaSuccess) { set_error(std::string("cudaMalloc (matcher staging): ") + cudaGetErrorString(e)); return PLVI_ERR_CUDA; }
  m->staged = true;
  return PLVI_OK;
}
#define PLVI_STAGING(m) do { int rc_ = ensure_staging(m); if (rc_ != PLVI_OK) return rc_; } while (0)

// One stream-ordered allocation holding the inputs and outputs of a *_host entry point: segments are 256-byte aligned,
// inputs are copied in as they are added.
struct HostStage {
  cudaStream_t st; size_t total = 0; unsigned char* buf = nullptr; cudaError_t e = cudaSuccess;
  struct Seg { size_t off, bytes; const void* src; };
  std::vector<Seg> segs;
  explicit HostStage(cudaStream_t s) : st(s) {}
  int add(const void* src, size_t bytes) {   // src may be nullptr (output / scratch segment)
    segs.push_back({total, bytes, src});
    total += (bytes + 255) & ~(size_t)255;
    return (int)segs.size() - 1;
  }
  bool commit() {
    e = cudaMallocAsync(reinterpret_cast<void**>(&buf), total ? total : 256, st);
    for (size_t i = 0; i < segs.size() && e == cudaSuccess; i++)
      if (segs[i].src && segs[i].bytes) e = cudaMemcpyAsync(buf + segs[i].off, segs[i].src, segs[i].bytes, cudaMemcpyHostToDevice, st);
    return e == cudaSuccess;
  }
  template <typename T> T* ptr(int i) const { return reinterpret_cast<T*>(buf + segs[i].off); }
  void fetch(void* dst, int i, size_t bytes) {
    if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(dst, buf + segs[i].off, bytes, cudaMemcpyDeviceToHost, st);
  }
  int finish(int rc) {
    if (buf) cudaFreeAsync(buf, st);
    const cudaError_t s = cudaStreamSynchronize(st);
    if (rc != PLVI_OK) return rc;
    if (e == cudaSuccess) e = s;
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return PLVI_ERR_CUDA; }
    return PLVI_OK;
  }
};

// ---------------------------------------------------------------------------------
// k_line_match_grid: the line search of Frame::ComputeStereoMatches_Lines (src/Frame.cc:1421-1448) =
// GridStructure fill along LineIterator (src/LineIterator.cpp:9-52, src/gridStructure.cpp:14-23,49-60) +
// LineMatcher::matchGrid (src/LineMatcher.cpp:191-272).  One warp per stereo pair.  The per-cell index lists and the
// unordered_set of candidates become one occupancy record per right line in shared memory: the cells a digital line
// passes in one grid row are a contiguous run of columns (the iterator advances one cell along its major axis per step
// and at most one across), so row y of line j is the byte pair (first, last column), (255, 0) when empty; layout
// [row][line] so that the lanes' reads are conflict free (96 B per line instead of a 384 B bitmap: 7 pairs per SM
// instead of 2).  A right line is a candidate iff a run meets the window of the left line's start cell or end cell.  Left lines are visited in order (the
// distances[] / matches_21[] pre-emption couples them), their candidates in parallel, lane j owning right lines
// j, j+32, ...; best / second best are merged by shuffles (the candidate order does not matter: a tie on the best
// distance fails the ratio test).  All double arithmetic is uncontracted, as in the reference's x86-64 build.
// ---------------------------------------------------------------------------------

__global__ void __launch_bounds__(32) k_line_match_grid(const float* __restrict__ seg1, const uint8_t* __restrict__ desc1,
                                                        const int* __restrict__ n1p, int stride1,
                                                        const float* __restrict__ seg2, const uint8_t* __restrict__ desc2,
                                                        const int* __restrict__ n2p, int stride2, double invw, double invh,
                                                        int rows, int cols, int wl, int wr, int wu, int wd,
                                                        int* __restrict__ matches12, int* __restrict__ nmatches,
                                                        const uchar2* __restrict__ occ2, const double* __restrict__ dir2,
                                                        const int* __restrict__ coords1) {
  // occ2 / dir2 / coords1 != nullptr (one pair): the caller hands over what the reference's matchGrid receives -- the
  // filled GridStructure as (first, last) column per right line and grid row [n2][rows], directions2 [n2][2] and the
  // left lines' integer end cells lines1 [n1][4] -- instead of the segments
  extern __shared__ __align__(16) unsigned char lmg_smem[];
  double* dirx = reinterpret_cast<double*>(lmg_smem);                           // [stride2]
  double* diry = dirx + stride2;
  uint32_t* sd2 = reinterpret_cast<uint32_t*>(diry + stride2);                  // [8][stride2]
  int* dist = reinterpret_cast<int*>(sd2 + 8 * (size_t)stride2);                // [stride2]
  int* m21 = dist + stride2;
  uchar2* run = reinterpret_cast<uchar2*>(m21 + stride2);                       // [rows][stride2] (first, last) column
  const int pair = blockIdx.x, lane = threadIdx.x;
  const int n1 = min(n1p[pair], stride1), n2 = min(n2p[pair], stride2);
  seg1 += (size_t)pair * stride1 * 4;
  seg2 += (size_t)pair * stride2 * 4;
  desc1 += (size_t)pair * stride1 * 32;
  desc2 += (size_t)pair * stride2 * 32;
  int* m12 = matches12 + (size_t)pair * stride1;
  for (int i = lane; i < rows * stride2; i += 32) run[i] = make_uchar2(255, 0);
  __syncwarp();
  for (int j = lane; j < n2; j += 32) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(desc2 + 32 * j)), b = __ldg(reinterpret_cast<const uint4*>(desc2 + 32 * j) + 1);
    sd2[0 * stride2 + j] = a.x; sd2[1 * stride2 + j] = a.y; sd2[2 * stride2 + j] = a.z; sd2[3 * stride2 + j] = a.w;
    sd2[4 * stride2 + j] = b.x; sd2[5 * stride2 + j] = b.y; sd2[6 * stride2 + j] = b.z; sd2[7 * stride2 + j] = b.w;
    dist[j] = INT_MAX;
    m21[j] = -1;
    if (occ2) {
      dirx[j] = dir2[2 * j];
      diry[j] = dir2[2 * j + 1];
      for (int y = 0; y < rows; y++) run[y * stride2 + j] = occ2[(size_t)j * rows + y];
      continue;
    }
    const float4 s = *reinterpret_cast<const float4*>(seg2 + 4 * j);
    const double vx = __dmul_rn((double)__fsub_rn(s.z, s.x), invw), vy = __dmul_rn((double)__fsub_rn(s.w, s.y), invh);
    const double mag = __dsqrt_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
    dirx[j] = __ddiv_rn(vx, mag);
    diry[j] = __ddiv_rn(vy, mag);
    // LineIterator over the right line in grid units
    double x1 = __dmul_rn((double)s.x, invw), y1 = __dmul_rn((double)s.y, invh);
    double x2 = __dmul_rn((double)s.z, invw), y2 = __dmul_rn((double)s.w, invh);
    const bool steep = fabs(__dsub_rn(y2, y1)) > fabs(__dsub_rn(x2, x1));
    if (steep) { double t = x1; x1 = y1; y1 = t; t = x2; x2 = y2; y2 = t; }
    if (x1 > x2) { double t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    const double dx = __dsub_rn(x2, x1), dy = fabs(__dsub_rn(y2, y1));
    double error = __ddiv_rn(dx, 2.0);
    const int ystep = (y1 < y2) ? 1 : -1;
    int y = (int)y1;
    const int maxX = (int)x2;
    for (int x = (int)x1; x <= maxX; x++) {
      const int cx = steep ? y : x, cy = steep ? x : y;
      if (cx >= 0 && cx < cols && cy >= 0 && cy < rows) {
        uchar2 r = run[cy * stride2 + j];
        r.x = min((int)r.x, cx);
        r.y = max((int)r.y, cx);
        run[cy * stride2 + j] = r;
      }
      error = __dsub_rn(error, dy);
      if (error < 0) { y += ystep; error = __dadd_rn(error, dx); }
    }
  }
  for (int i = lane; i < n1; i += 32) m12[i] = -1;
  __syncwarp();
  int matches = 0;
  for (int i1 = 0; i1 < n1; i1++) {
    // line_2d holds int pairs (include/LineMatcher.h:41-42): the end points are truncated to grid cells
    int sx, sy, ex, ey;
    if (coords1) {
      sx = coords1[4 * i1]; sy = coords1[4 * i1 + 1]; ex = coords1[4 * i1 + 2]; ey = coords1[4 * i1 + 3];
    } else {
      const float4 s = __ldg(reinterpret_cast<const float4*>(seg1 + 4 * i1));
      sx = (int)__dmul_rn((double)s.x, invw); sy = (int)__dmul_rn((double)s.y, invh);
      ex = (int)__dmul_rn((double)s.z, invw); ey = (int)__dmul_rn((double)s.w, invh);
    }
    double vx = (double)(ex - sx), vy = (double)(ey - sy);
    const double mag = __dsqrt_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
    vx = __ddiv_rn(vx, mag);
    vy = __ddiv_rn(vy, mag);
    const int slo = max(0, sx - wl), shi = min(cols, sx + wr + 1), elo = max(0, ex - wl), ehi = min(cols, ex + wr + 1);   // [lo, hi)
    const int s0 = max(0, sy - wu), s1 = min(rows, sy + wd + 1);
    const int e0 = max(0, ey - wu), e1 = min(rows, ey + wd + 1);
    uint32_t q[8];
    {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(desc1 + 32 * i1)), b = __ldg(reinterpret_cast<const uint4*>(desc1 + 32 * i1) + 1);
      q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    int best = INT_MAX, best2 = INT_MAX, bidx = -1;
    for (int i2 = lane; i2 < n2; i2 += 32) {
      bool hit = false;
      for (int y = s0; y < s1; y++) { const uchar2 r = run[y * stride2 + i2]; hit |= (int)r.x < shi && (int)r.y >= slo; }
      for (int y = e0; y < e1; y++) { const uchar2 r = run[y * stride2 + i2]; hit |= (int)r.x < ehi && (int)r.y >= elo; }
      if (!hit) continue;
      if (fabs(__dadd_rn(__dmul_rn(vx, dirx[i2]), __dmul_rn(vy, diry[i2]))) < 0.75) continue;
      int d = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) d += __popc(q[k] ^ sd2[k * stride2 + i2]);
      if (d < dist[i2]) { dist[i2] = d; m21[i2] = i1; } else continue;
      if (d < best) { best2 = best; best = d; bidx = i2; }
      else if (d < best2) best2 = d;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const int ob = __shfl_xor_sync(0xffffffffu, best, o), ob2 = __shfl_xor_sync(0xffffffffu, best2, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob < best) { best2 = min(best, ob2); best = ob; bidx = oi; }
      else best2 = min(best2, ob);
    }
    if ((double)best < __dmul_rn((double)best2, 0.9)) {
      if (lane == 0) m12[i1] = bidx;
      matches++;
    }
  }
  __syncwarp();
  int dropped = 0;
  for (int i1 = lane; i1 < n1; i1 += 32) {
    const int i2 = m12[i1];
    if (i2 >= 0 && m21[i2] != i1) { m12[i1] = -1; dropped++; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) dropped += __shfl_xor_sync(0xffffffffu, dropped, o);
  if (lane == 0) nmatches[pair] = matches - dropped;
}

extern "C" {

int plvi_matcher_create(plvi_matcher** out, int max_pairs, int max_train, int max_query, int device,
                        void* stream) {
  if (!out || max_pairs < 1 || max_train < 1 || max_query < 1 || max_train > 65535) {
    set_error("plvi_matcher_create: invalid argument (max_train <= 65535)");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(device));
  plvi_matcher* m = new plvi_matcher();
  m->device = device;
  m->maxPairs = max_pairs;
  m->maxTrain = max_train;
  m->maxQuery = max_query;
  if (stream) m->stream = (cudaStream_t)stream;
  else {
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); delete m; return PLVI_ERR_CUDA; }
    m->ownStream = true;
  }
  const cudaError_t e = cudaMalloc((void**)&m->dMatchedDist, (size_t)max_pairs * max_train * sizeof(int));
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
    plvi_matcher_destroy(m);
    return PLVI_ERR_CUDA;
  }
  *out = m;
  return PLVI_OK;
}

void plvi_matcher_destroy(plvi_matcher* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  cudaFree(m->dKeys); cudaFree(m->dDesc); cudaFree(m->dBlocked); cudaFree(m->dTCount);
  cudaFree(m->dQ); cudaFree(m->dQDesc); cudaFree(m->dQCount); cudaFree(m->dMatchTrain);
  cudaFree(m->dMatchQuery); cudaFree(m->dNMatches); cudaFree(m->dMatchedDist);
  cudaFree(m->dStereoTrain); cudaFree(m->dStereoQuery);
  if (m->ownStream && m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

void* plvi_matcher_stream(const plvi_matcher* m) { return m ? (void*)m->stream : nullptr; }
int plvi_matcher_last_launches(const plvi_matcher* m) { return m ? m->lastLaunches : PLVI_ERR_INVALID; }

int plvi_hamming256(plvi_matcher* m, const uint8_t* a, const uint8_t* b, int n, int shift25, int* out,
                    int on_device) {
  if (!m || !a || !b || !out || n < 0) { set_error("plvi_hamming256: invalid argument"); return PLVI_ERR_INVALID; }
  if (n == 0) return PLVI_OK;
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  const uint8_t *da = a, *db = b;
  int* dout = out;
  if (!on_device) {
    if ((size_t)n > (size_t)m->maxPairs * (m->maxTrain > m->maxQuery ? m->maxTrain : m->maxQuery)) {
      set_error("plvi_hamming256: n exceeds matcher capacity");
      return PLVI_ERR_CAPACITY;
    }
    PLVI_STAGING(m);
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dDesc, a, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dQDesc, b, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream));
    da = m->dDesc; db = m->dQDesc; dout = m->dMatchQuery;
  }
  k_hamming_pairs<<<(n + 255) / 256, 256, 0, m->stream>>>(da, db, n, shift25, dout);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  if (!on_device) {
    PLVI_CUDA_TRY(cudaMemcpyAsync(out, dout, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    PLVI_CUDA_TRY(cudaStreamSynchronize(m->stream));
  }
  return PLVI_OK;
}

int plvi_search_by_projection(plvi_matcher* m, int mode, int npairs, const plvi_keypoint* train_keys,
                              const uint8_t* train_desc, const uint8_t* train_blocked,
                              const int* train_counts, int train_stride, const plvi_grid* grid,
                              plvi_query* queries, const uint8_t* query_desc, const int* query_counts,
                              int query_stride, int th_dist, float nnratio, int check_orientation,
                              int* match_train, int* match_query, int* nmatches, int on_device) {
  if (!m || mode < 0 || mode > 2 || npairs < 1 || !train_keys || !train_desc || !train_counts || !grid ||
      !queries || !query_desc || !query_counts || !match_train || !match_query || !nmatches) {
    set_error("plvi_search_by_projection: invalid argument");
    return PLVI_ERR_INVALID;
  }
  if (npairs > m->maxPairs || train_stride > m->maxTrain || query_stride > m->maxQuery) {
    set_error("plvi_search_by_projection: exceeds matcher capacity");
    return PLVI_ERR_CAPACITY;
  }
  if (mode == 2 && query_stride > 65534) {   // the kernel keeps the claiming query of a feature as 16 bits
    set_error("plvi_search_by_projection: initialisation mode takes at most 65534 queries per pair");
    return PLVI_ERR_CAPACITY;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  SearchArgs a;
  a.mode = mode;
  a.grid = *grid;
  a.tstride = train_stride;
  a.qstride = query_stride;
  a.th = th_dist;
  a.nnratio = nnratio;
  a.checkOri = check_orientation;
  a.matchedDist = mode == 2 ? m->dMatchedDist : nullptr;
  a.items = nullptr;
  a.istride = 0;
  a.uright = m->stereoTrain; a.qur = m->stereoQuery;   // side information of this call only
  m->stereoTrain = m->stereoQuery = nullptr;
  const size_t P = npairs, T = train_stride, Q = query_stride;
  cudaStream_t st = m->stream;
  if (on_device) {
    a.keys = train_keys; a.desc = train_desc; a.blocked = train_blocked; a.tcount = train_counts;
    a.q = queries; a.qdesc = query_desc; a.qcount = query_counts;
    a.matchTrain = match_train; a.matchQuery = match_query; a.nmatches = nmatches;
  } else {
    for (size_t p = 0; p < P; p++)
      if (train_counts[p] < 0 || train_counts[p] > train_stride || query_counts[p] < 0 || query_counts[p] > query_stride) {
        set_error("plvi_search_by_projection: count exceeds stride");
        return PLVI_ERR_INVALID;
      }
    PLVI_STAGING(m);
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dKeys, train_keys, P * T * sizeof(plvi_keypoint), cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dDesc, train_desc, P * T * 32, cudaMemcpyHostToDevice, st));
    if (train_blocked) PLVI_CUDA_TRY(cudaMemcpyAsync(m->dBlocked, train_blocked, P * T, cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dTCount, train_counts, P * sizeof(int), cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dQ, queries, P * Q * sizeof(plvi_query), cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dQDesc, query_desc, P * Q * 32, cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dQCount, query_counts, P * sizeof(int), cudaMemcpyHostToDevice, st));
    a.keys = m->dKeys; a.desc = m->dDesc; a.blocked = train_blocked ? m->dBlocked : nullptr;
    a.tcount = m->dTCount; a.q = m->dQ; a.qdesc = m->dQDesc; a.qcount = m->dQCount;
    a.matchTrain = m->dMatchTrain; a.matchQuery = m->dMatchQuery; a.nmatches = m->dNMatches;
  }
  const size_t smem = (GRID_CELLS * 2 + 1) * sizeof(int) + T * 2 + T + Q + T + 16 + (mode == 2 ? T * 4 : 0);   // + octaves of the items; init mode: claimed distances + owners
  if (smem > 227 * 1024) {   // items, flags, octaves (+ claimed distances / owners) of one pair live in shared memory
    set_error("search: train / query strides too large for the shared-memory tables of one pair");
    return PLVI_ERR_CAPACITY;
  }
  if (smem > 48 * 1024)
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_search<<<npairs, SEARCH_WARPS * 32, smem, st>>>(a);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  if (!on_device) {
    PLVI_CUDA_TRY(cudaMemcpyAsync(match_train, m->dMatchTrain, P * T * sizeof(int), cudaMemcpyDeviceToHost, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(match_query, m->dMatchQuery, P * Q * sizeof(int), cudaMemcpyDeviceToHost, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(nmatches, m->dNMatches, P * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (mode == 2)
      PLVI_CUDA_TRY(cudaMemcpyAsync(queries, m->dQ, P * Q * sizeof(plvi_query), cudaMemcpyDeviceToHost, st));
    PLVI_CUDA_TRY(cudaStreamSynchronize(st));
  }
  return PLVI_OK;
}

static int search_by_bow_impl(plvi_matcher* m, int npairs, const plvi_keypoint* train_keys, const uint8_t* train_desc,
                              const uint8_t* train_blocked, const int* train_counts, int train_stride, const int* group_items,
                              int items_stride, const plvi_query* queries, const uint8_t* query_desc, const int* query_counts,
                              int query_stride, int th_dist, float nnratio, int check_orientation, int* match_train,
                              int* match_query, int* nmatches, int on_device) {
  if (!m || npairs < 1 || !train_keys || !train_desc || !train_counts || !group_items || !queries || !query_desc ||
      !query_counts || !match_train || !match_query || !nmatches || items_stride < 1) {
    set_error("plvi_search_by_bow: invalid argument");
    return PLVI_ERR_INVALID;
  }
  if (npairs > m->maxPairs || train_stride > m->maxTrain || query_stride > m->maxQuery || items_stride > m->maxTrain) {
    set_error("plvi_search_by_bow: exceeds matcher capacity");
    return PLVI_ERR_CAPACITY;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  SearchArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = 3;
  a.grid = plvi_grid{0.f, 0.f, 1.f, 1.f};
  a.tstride = train_stride;
  a.qstride = query_stride;
  a.istride = items_stride;
  a.th = th_dist;
  a.nnratio = nnratio;
  a.checkOri = check_orientation;
  const size_t P = npairs, T = train_stride, Q = query_stride;
  cudaStream_t st = m->stream;
  if (on_device) {
    a.keys = train_keys; a.desc = train_desc; a.tcount = train_counts; a.items = group_items;
    a.blocked = train_blocked;
    a.q = const_cast<plvi_query*>(queries); a.qdesc = query_desc; a.qcount = query_counts;
    a.matchTrain = match_train; a.matchQuery = match_query; a.nmatches = nmatches;
  } else {
    PLVI_STAGING(m);
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dKeys, train_keys, P * T * sizeof(plvi_keypoint), cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dDesc, train_desc, P * T * 32, cudaMemcpyHostToDevice, st));
    if (train_blocked) {
      PLVI_CUDA_TRY(cudaMemcpyAsync(m->dBlocked, train_blocked, P * T, cudaMemcpyHostToDevice, st));
      a.blocked = m->dBlocked;
    }
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dTCount, train_counts, P * sizeof(int), cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dMatchedDist, group_items, P * items_stride * sizeof(int), cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dQ, queries, P * Q * sizeof(plvi_query), cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dQDesc, query_desc, P * Q * 32, cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dQCount, query_counts, P * sizeof(int), cudaMemcpyHostToDevice, st));
    a.keys = m->dKeys; a.desc = m->dDesc; a.tcount = m->dTCount; a.items = m->dMatchedDist; a.istride = items_stride;
    a.q = m->dQ; a.qdesc = m->dQDesc; a.qcount = m->dQCount;
    a.matchTrain = m->dMatchTrain; a.matchQuery = m->dMatchQuery; a.nmatches = m->dNMatches;
  }
  const size_t smem = (GRID_CELLS * 2 + 1) * sizeof(int) + T * 2 + T + Q + T + 16;
  if (smem > 227 * 1024) {   // items, flags, octaves (+ claimed distances / owners) of one pair live in shared memory
    set_error("search: train / query strides too large for the shared-memory tables of one pair");
    return PLVI_ERR_CAPACITY;
  }
  if (smem > 48 * 1024)
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_search<<<npairs, SEARCH_WARPS * 32, smem, st>>>(a);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  if (!on_device) {
    PLVI_CUDA_TRY(cudaMemcpyAsync(match_train, m->dMatchTrain, P * T * sizeof(int), cudaMemcpyDeviceToHost, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(match_query, m->dMatchQuery, P * Q * sizeof(int), cudaMemcpyDeviceToHost, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(nmatches, m->dNMatches, P * sizeof(int), cudaMemcpyDeviceToHost, st));
    PLVI_CUDA_TRY(cudaStreamSynchronize(st));
  }
  return PLVI_OK;
}

int plvi_search_by_bow(plvi_matcher* m, int npairs, const plvi_keypoint* train_keys, const uint8_t* train_desc,
                       const int* train_counts, int train_stride, const int* group_items, int items_stride,
                       const plvi_query* queries, const uint8_t* query_desc, const int* query_counts,
                       int query_stride, int th_dist, float nnratio, int check_orientation, int* match_train,
                       int* match_query, int* nmatches, int on_device) {
  return search_by_bow_impl(m, npairs, train_keys, train_desc, nullptr, train_counts, train_stride, group_items, items_stride,
                            queries, query_desc, query_counts, query_stride, th_dist, nnratio, check_orientation, match_train,
                            match_query, nmatches, on_device);
}

int plvi_search_by_bow_kf(plvi_matcher* m, int npairs, const plvi_keypoint* train_keys, const uint8_t* train_desc,
                          const uint8_t* train_blocked, const int* train_counts, int train_stride, const int* group_items,
                          int items_stride, const plvi_query* queries, const uint8_t* query_desc, const int* query_counts,
                          int query_stride, int th_low, float nnratio, int check_orientation, int* match_train,
                          int* match_query, int* nmatches, int on_device) {
  // the keyframe-keyframe variant accepts bestDist1 < TH_LOW (strict, src/ORBmatcher.cc:911); distances are integers
  return search_by_bow_impl(m, npairs, train_keys, train_desc, train_blocked, train_counts, train_stride, group_items,
                            items_stride, queries, query_desc, query_counts, query_stride, th_low - 1, nnratio,
                            check_orientation, match_train, match_query, nmatches, on_device);
}

int plvi_queries_from_keypoints(plvi_matcher* m, const plvi_keypoint* d_kps, const int* d_counts, int npairs,
                                int stride, float th, float scale_factor, plvi_query* d_queries) {
  if (!m || !d_kps || !d_counts || !d_queries || npairs < 1 || stride < 1) {
    set_error("plvi_queries_from_keypoints: invalid argument");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  k_queries_from_keypoints<<<dim3((stride + 255) / 256, npairs), 256, 0, m->stream>>>(d_kps, d_counts, stride, th,
                                                                                     scale_factor, d_queries);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_line_match(plvi_matcher* m, int npairs, const uint8_t* desc1, const int* n1, int stride1,
                    const uint8_t* desc2, const int* n2, int stride2, float nnr, int mutual,
                    int* matches12, int* nmatches, int on_device) {
  if (!m || npairs < 1 || !desc1 || !desc2 || !n1 || !n2 || !matches12 || !nmatches || stride1 < 1 || stride2 < 1) {
    set_error("plvi_line_match: invalid argument");
    return PLVI_ERR_INVALID;
  }
  const int cap = m->maxTrain > m->maxQuery ? m->maxTrain : m->maxQuery;
  if (npairs > m->maxPairs || stride1 > cap || stride2 > cap) {
    set_error("plvi_line_match: exceeds matcher capacity");
    return PLVI_ERR_CAPACITY;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  cudaStream_t st = m->stream;
  const size_t P = npairs;
  const uint8_t *a = desc1, *b = desc2;
  const int *c1 = n1, *c2 = n2;
  int *o = matches12, *nm = nmatches;
  if (!on_device) {
    PLVI_STAGING(m);
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dDesc, desc1, P * stride1 * 32, cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dQDesc, desc2, P * stride2 * 32, cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dTCount, n1, P * sizeof(int), cudaMemcpyHostToDevice, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(m->dQCount, n2, P * sizeof(int), cudaMemcpyHostToDevice, st));
    a = m->dDesc; b = m->dQDesc; c1 = m->dTCount; c2 = m->dQCount; o = m->dMatchQuery; nm = m->dNMatches;
  }
  const size_t smem = (size_t)(stride1 + stride2) * 32 + (size_t)(stride1 + stride2) * sizeof(int);
  if (smem > 200 * 1024) { set_error("plvi_line_match: descriptor sets too large for one CTA"); return PLVI_ERR_CAPACITY; }
  if (smem > 48 * 1024)
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_line_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_line_match<<<npairs, 256, smem, st>>>(a, c1, stride1, b, c2, stride2, nnr, mutual, o, nm);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  if (!on_device) {
    PLVI_CUDA_TRY(cudaMemcpyAsync(matches12, o, P * stride1 * sizeof(int), cudaMemcpyDeviceToHost, st));
    PLVI_CUDA_TRY(cudaMemcpyAsync(nmatches, nm, P * sizeof(int), cudaMemcpyDeviceToHost, st));
    PLVI_CUDA_TRY(cudaStreamSynchronize(st));
  }
  return PLVI_OK;
}

int plvi_line_match_mad(plvi_matcher* m, int npairs, const uint8_t* desc1, const int* n1, int stride1, const uint8_t* desc2,
                        const int* n2, int stride2, const uint8_t* has_line1, const uint8_t* has_line2, double factor,
                        int* matches12, int* nmatches, double* mad) {
  if (!m || npairs < 1 || !desc1 || !desc2 || !n1 || !n2 || !matches12 || !nmatches || !mad || stride1 < 1 || stride2 < 1) {
    set_error("plvi_line_match_mad: invalid argument");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  const size_t smem = (size_t)(stride1 + stride2) * 32 + (size_t)stride1 * 3 * sizeof(int);
  if (smem > 200 * 1024) { set_error("plvi_line_match_mad: descriptor sets too large for one CTA"); return PLVI_ERR_CAPACITY; }
  if (smem > 48 * 1024)
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_line_match_mad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_line_match_mad<<<npairs, 256, smem, m->stream>>>(desc1, n1, stride1, desc2, n2, stride2, has_line1, has_line2, factor,
                                                     matches12, nmatches, mad);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_line_match_grid(plvi_matcher* m, int npairs, const float* d_seg1, const uint8_t* d_desc1, const int* d_n1,
                         int stride1, const float* d_seg2, const uint8_t* d_desc2, const int* d_n2, int stride2,
                         double inv_width, double inv_height, int grid_rows, int grid_cols, int win_left, int win_right,
                         int win_up, int win_down, int* d_matches12, int* d_nmatches) {
  if (!m || npairs < 1 || !d_seg1 || !d_desc1 || !d_n1 || !d_seg2 || !d_desc2 || !d_n2 || !d_matches12 || !d_nmatches ||
      stride1 < 1 || stride2 < 1 || grid_rows < 1 || grid_cols < 1 || grid_rows > 64 || grid_cols > 64 || win_left < 0 ||
      win_right < 0 || win_up < 0 || win_down < 0 || !(inv_width > 0.0) || !(inv_height > 0.0)) {
    set_error("plvi_line_match_grid: invalid argument (grid at most 64 x 64 cells)");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  const size_t smem = (size_t)stride2 * ((size_t)grid_rows * 2 + 2 * sizeof(double) + 32 + 2 * sizeof(int));
  if (smem > 200 * 1024) { set_error("plvi_line_match_grid: right line set too large for one CTA"); return PLVI_ERR_CAPACITY; }
  if (smem > 48 * 1024)
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_line_match_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_line_match_grid<<<npairs, 32, smem, m->stream>>>(d_seg1, d_desc1, d_n1, stride1, d_seg2, d_desc2, d_n2, stride2, inv_width,
                                                    inv_height, grid_rows, grid_cols, win_left, win_right, win_up, win_down,
                                                    d_matches12, d_nmatches, nullptr, nullptr, nullptr);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_line_match_grid_host(plvi_matcher* m, const float* seg1, const uint8_t* desc1, int n1, const float* seg2,
                              const uint8_t* desc2, int n2, double inv_width, double inv_height, int grid_rows,
                              int grid_cols, int win_left, int win_right, int win_up, int win_down, int* matches12,
                              int* nmatches) {
  if (!m || n1 < 0 || n2 < 0 || !matches12 || !nmatches || (n1 && (!seg1 || !desc1)) || (n2 && (!seg2 || !desc2))) {
    set_error("plvi_line_match_grid_host: invalid argument");
    return PLVI_ERR_INVALID;
  }
  *nmatches = 0;
  if (n1 == 0) return PLVI_OK;
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  cudaStream_t st = m->stream;
  const int s1 = n1, s2 = n2 > 0 ? n2 : 1;
  const size_t bSeg1 = (size_t)s1 * 16, bSeg2 = (size_t)s2 * 16, bD1 = (size_t)s1 * 32, bD2 = (size_t)s2 * 32;
  const size_t oSeg2 = bSeg1, oD1 = oSeg2 + bSeg2, oD2 = oD1 + bD1, oCnt = oD2 + bD2, oM = oCnt + 16;
  const size_t total = oM + (size_t)s1 * sizeof(int);
  unsigned char* buf = nullptr;
  PLVI_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&buf), total, st));
  const int cnt[3] = {n1, n2, 0};
  cudaError_t e = cudaMemcpyAsync(buf, seg1, bSeg1, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(buf + oD1, desc1, bD1, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n2) e = cudaMemcpyAsync(buf + oSeg2, seg2, (size_t)n2 * 16, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n2) e = cudaMemcpyAsync(buf + oD2, desc2, (size_t)n2 * 32, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(buf + oCnt, cnt, sizeof(cnt), cudaMemcpyHostToDevice, st);
  int rc = PLVI_OK;
  if (e == cudaSuccess) {
    int* dc = reinterpret_cast<int*>(buf + oCnt);
    rc = plvi_line_match_grid(m, 1, reinterpret_cast<const float*>(buf), buf + oD1, dc, s1,
                              reinterpret_cast<const float*>(buf + oSeg2), buf + oD2, dc + 1, s2, inv_width, inv_height,
                              grid_rows, grid_cols, win_left, win_right, win_up, win_down, reinterpret_cast<int*>(buf + oM),
                              dc + 2);
    if (rc == PLVI_OK) {
      e = cudaMemcpyAsync(matches12, buf + oM, (size_t)n1 * sizeof(int), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(nmatches, dc + 2, sizeof(int), cudaMemcpyDeviceToHost, st);
    }
  }
  cudaFreeAsync(buf, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (rc != PLVI_OK) return rc;
  PLVI_CUDA_TRY(e);
  return PLVI_OK;
}

int plvi_distinctive_descriptors(void* stream, const uint8_t* d_desc, const int* d_counts, int n_points, int stride,
                                 int* d_best_idx, uint8_t* d_best_desc) {
  if (!d_desc || !d_counts || !d_best_idx || n_points < 1 || stride < 1) {
    set_error("plvi_distinctive_descriptors: invalid argument");
    return PLVI_ERR_INVALID;
  }
  if (stride > 4096) { set_error("plvi_distinctive_descriptors: more than 4096 observations per map point"); return PLVI_ERR_CAPACITY; }
  const size_t smem = (size_t)stride * 32 + 32 * 258 * sizeof(unsigned short);
  if (smem > 48 * 1024)
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_distinctive, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_distinctive<<<n_points, 32, smem, (cudaStream_t)stream>>>(d_desc, d_counts, stride, d_best_idx, d_best_desc);
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_search_for_triangulation(plvi_matcher* m, int npairs, const plvi_keypoint* train_keys, const uint8_t* train_desc,
                                  const uint8_t* train_blocked, const int* train_counts, int train_stride, const int* group_items,
                                  int items_stride, const plvi_query* queries, const uint8_t* query_desc, const int* query_counts,
                                  int query_stride, const plvi_epipolar* geometry, int th_low, int check_orientation,
                                  int* match_query, int* nmatches) {
  if (!m || npairs < 1 || !train_keys || !train_desc || !train_counts || !group_items || !queries || !query_desc || !query_counts ||
      !geometry || !match_query || !nmatches || train_stride < 1 || items_stride < 1 || query_stride < 1) {
    set_error("plvi_search_for_triangulation: invalid argument");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  k_search_triangulation<<<npairs, 256, 0, m->stream>>>(train_keys, train_desc, train_blocked, train_counts, train_stride, group_items,
                                                        items_stride, queries, query_desc, query_counts, query_stride, geometry, th_low,
                                                        check_orientation, match_query, nmatches);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_search_in_radius(plvi_matcher* m, int npairs, const plvi_keypoint* train_keys, const uint8_t* train_desc,
                          const int* train_counts, int train_stride, const plvi_grid* grid, const plvi_query* queries,
                          const uint8_t* query_desc, const int* query_counts, int query_stride, const float* inv_level_sigma2,
                          double chi2, int th_dist, int* best_idx, int* best_dist, int* nfound) {
  if (!m || npairs < 1 || !train_keys || !train_desc || !train_counts || !grid || !queries || !query_desc || !query_counts ||
      !inv_level_sigma2 || !best_idx || !best_dist || !nfound || train_stride < 1 || query_stride < 1 || train_stride > 65535) {
    set_error("plvi_search_in_radius: invalid argument");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  RadiusArgs a;
  a.keys = train_keys; a.desc = train_desc; a.tcount = train_counts; a.tstride = train_stride;
  a.grid = *grid;
  a.q = queries; a.qdesc = query_desc; a.qcount = query_counts; a.qstride = query_stride;
  for (int i = 0; i < 16; i++) a.invSigma2[i] = inv_level_sigma2[i];
  a.chi2 = chi2; a.th = th_dist;
  a.bestIdx = best_idx; a.bestDist = best_dist; a.nfound = nfound;
  a.uright = m->stereoTrain; a.qur = m->stereoQuery;   // side information of this call only
  m->stereoTrain = m->stereoQuery = nullptr;
  const size_t smem = (GRID_CELLS * 2 + 1) * sizeof(int) + (size_t)train_stride * 2 + 16;
  if (smem > 48 * 1024)
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_search_in_radius, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_search_in_radius<<<npairs, SEARCH_WARPS * 32, smem, m->stream>>>(a);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_line_fuse_search(plvi_matcher* m, int npairs, const plvi_keyline* keylines, const uint8_t* desc, const int* counts,
                          int stride, const float* queries, const uint8_t* query_flags, const uint8_t* query_desc,
                          const int* query_counts, int query_stride, int th_low, int* best_idx, int* best_dist, int* nfound) {
  if (!m || npairs < 1 || !keylines || !desc || !counts || !queries || !query_desc || !query_counts || !best_idx || !best_dist ||
      !nfound || stride < 1 || query_stride < 1) {
    set_error("plvi_line_fuse_search: invalid argument");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  PLVI_CUDA_TRY(cudaMemsetAsync(nfound, 0, sizeof(int) * npairs, m->stream));
  k_line_fuse_search<<<dim3((query_stride + 7) / 8, npairs), 256, 0, m->stream>>>(keylines, desc, counts, stride, queries, query_flags,
                                                                                 query_desc, query_counts, query_stride, th_low,
                                                                                 best_idx, best_dist, nfound);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_matcher_set_stereo(plvi_matcher* m, const float* train_uright, const float* query_ur, int npairs, int train_stride,
                            int query_stride, int on_device) {
  if (!m) { set_error("plvi_matcher_set_stereo: invalid argument"); return PLVI_ERR_INVALID; }
  m->stereoTrain = m->stereoQuery = nullptr;
  if (!train_uright || !query_ur) return PLVI_OK;   // clears the side information
  if (on_device) { m->stereoTrain = train_uright; m->stereoQuery = query_ur; return PLVI_OK; }
  if (npairs < 1 || npairs > m->maxPairs || train_stride < 1 || train_stride > m->maxTrain || query_stride < 1 ||
      query_stride > m->maxQuery) {
    set_error("plvi_matcher_set_stereo: exceeds matcher capacity");
    return PLVI_ERR_CAPACITY;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  PLVI_STAGING(m);
  PLVI_CUDA_TRY(cudaMemcpyAsync(m->dStereoTrain, train_uright, (size_t)npairs * train_stride * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  PLVI_CUDA_TRY(cudaMemcpyAsync(m->dStereoQuery, query_ur, (size_t)npairs * query_stride * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  // pageable host memory: the copies above have been staged by the runtime when the calls return
  m->stereoTrain = m->dStereoTrain; m->stereoQuery = m->dStereoQuery;
  return PLVI_OK;
}

int plvi_search_in_radius_host(plvi_matcher* m, const plvi_keypoint* train_keys, const uint8_t* train_desc, int n_train,
                               const plvi_grid* grid, const plvi_query* queries, const uint8_t* query_desc, int n_query,
                               const float* inv_level_sigma2, double chi2, int th_dist, const float* train_uright,
                               const float* query_ur, int* best_idx, int* best_dist, int* nfound) {
  if (!m || n_train < 0 || n_query < 0 || !grid || !inv_level_sigma2 || !nfound || (n_train && (!train_keys || !train_desc)) ||
      (n_query && (!queries || !query_desc || !best_idx || !best_dist))) {
    set_error("plvi_search_in_radius_host: invalid argument");
    return PLVI_ERR_INVALID;
  }
  *nfound = 0;
  for (int i = 0; i < n_query; i++) { best_idx[i] = -1; best_dist[i] = 256; }
  if (n_query == 0 || n_train == 0) return PLVI_OK;
  if (n_train > 65535) { set_error("plvi_search_in_radius_host: more than 65535 keypoints"); return PLVI_ERR_CAPACITY; }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  HostStage hs(m->stream);
  const int cnt[2] = {n_train, n_query};
  const int sK = hs.add(train_keys, (size_t)n_train * sizeof(plvi_keypoint)), sD = hs.add(train_desc, (size_t)n_train * 32);
  const int sQ = hs.add(queries, (size_t)n_query * sizeof(plvi_query)), sQD = hs.add(query_desc, (size_t)n_query * 32);
  const int sC = hs.add(cnt, sizeof(cnt));
  const bool stereo = train_uright && query_ur;
  const int sU = hs.add(stereo ? train_uright : nullptr, (size_t)n_train * sizeof(float));
  const int sQU = hs.add(stereo ? query_ur : nullptr, (size_t)n_query * sizeof(float));
  const int sBI = hs.add(nullptr, (size_t)n_query * sizeof(int)), sBD = hs.add(nullptr, (size_t)n_query * sizeof(int));
  const int sNF = hs.add(nullptr, sizeof(int));
  int rc = PLVI_OK;
  if (hs.commit()) {
    if (stereo) { m->stereoTrain = hs.ptr<float>(sU); m->stereoQuery = hs.ptr<float>(sQU); }
    rc = plvi_search_in_radius(m, 1, hs.ptr<plvi_keypoint>(sK), hs.ptr<uint8_t>(sD), hs.ptr<int>(sC), n_train, grid,
                               hs.ptr<plvi_query>(sQ), hs.ptr<uint8_t>(sQD), hs.ptr<int>(sC) + 1, n_query, inv_level_sigma2, chi2,
                               th_dist, hs.ptr<int>(sBI), hs.ptr<int>(sBD), hs.ptr<int>(sNF));
    if (rc == PLVI_OK) {
      hs.fetch(best_idx, sBI, (size_t)n_query * sizeof(int));
      hs.fetch(best_dist, sBD, (size_t)n_query * sizeof(int));
      hs.fetch(nfound, sNF, sizeof(int));
    }
  }
  return hs.finish(rc);
}

int plvi_search_for_triangulation_host(plvi_matcher* m, const plvi_keypoint* train_keys, const uint8_t* train_desc,
                                       const uint8_t* train_blocked, int n_train, const int* group_items, int n_items,
                                       const plvi_query* queries, const uint8_t* query_desc, int n_query,
                                       const plvi_epipolar* geometry, int th_low, int check_orientation, int* match_query,
                                       int* nmatches) {
  if (!m || n_train < 0 || n_items < 0 || n_query < 0 || !geometry || !nmatches || (n_train && (!train_keys || !train_desc)) ||
      (n_items && !group_items) || (n_query && (!queries || !query_desc || !match_query))) {
    set_error("plvi_search_for_triangulation_host: invalid argument");
    return PLVI_ERR_INVALID;
  }
  *nmatches = 0;
  for (int i = 0; i < n_query; i++) match_query[i] = -1;
  if (n_query == 0 || n_train == 0 || n_items == 0) return PLVI_OK;
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  HostStage hs(m->stream);
  const int cnt[2] = {n_train, n_query};
  const int sK = hs.add(train_keys, (size_t)n_train * sizeof(plvi_keypoint)), sD = hs.add(train_desc, (size_t)n_train * 32);
  const int sB = hs.add(train_blocked, train_blocked ? (size_t)n_train : 0), sI = hs.add(group_items, (size_t)n_items * sizeof(int));
  const int sQ = hs.add(queries, (size_t)n_query * sizeof(plvi_query)), sQD = hs.add(query_desc, (size_t)n_query * 32);
  const int sC = hs.add(cnt, sizeof(cnt)), sG = hs.add(geometry, sizeof(plvi_epipolar));
  const int sM = hs.add(nullptr, (size_t)n_query * sizeof(int)), sN = hs.add(nullptr, sizeof(int));
  int rc = PLVI_OK;
  if (hs.commit()) {
    rc = plvi_search_for_triangulation(m, 1, hs.ptr<plvi_keypoint>(sK), hs.ptr<uint8_t>(sD), train_blocked ? hs.ptr<uint8_t>(sB) : nullptr,
                                       hs.ptr<int>(sC), n_train, hs.ptr<int>(sI), n_items, hs.ptr<plvi_query>(sQ), hs.ptr<uint8_t>(sQD),
                                       hs.ptr<int>(sC) + 1, n_query, hs.ptr<plvi_epipolar>(sG), th_low, check_orientation,
                                       hs.ptr<int>(sM), hs.ptr<int>(sN));
    if (rc == PLVI_OK) { hs.fetch(match_query, sM, (size_t)n_query * sizeof(int)); hs.fetch(nmatches, sN, sizeof(int)); }
  }
  return hs.finish(rc);
}

int plvi_line_fuse_search_host(plvi_matcher* m, const plvi_keyline* keylines, const uint8_t* desc, int n, const float* queries,
                               const uint8_t* query_flags, const uint8_t* query_desc, int n_query, int th_low, int* best_idx,
                               int* best_dist, int* nfound) {
  if (!m || n < 0 || n_query < 0 || !nfound || (n && (!keylines || !desc)) ||
      (n_query && (!queries || !query_desc || !best_idx || !best_dist))) {
    set_error("plvi_line_fuse_search_host: invalid argument");
    return PLVI_ERR_INVALID;
  }
  *nfound = 0;
  for (int i = 0; i < n_query; i++) { best_idx[i] = -1; best_dist[i] = 0x7fffffff; }
  if (n_query == 0 || n == 0) return PLVI_OK;
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  HostStage hs(m->stream);
  const int cnt[2] = {n, n_query};
  const int sK = hs.add(keylines, (size_t)n * sizeof(plvi_keyline)), sD = hs.add(desc, (size_t)n * 32);
  const int sQ = hs.add(queries, (size_t)n_query * 6 * sizeof(float)), sF = hs.add(query_flags, query_flags ? (size_t)n_query : 0);
  const int sQD = hs.add(query_desc, (size_t)n_query * 32), sC = hs.add(cnt, sizeof(cnt));
  const int sBI = hs.add(nullptr, (size_t)n_query * sizeof(int)), sBD = hs.add(nullptr, (size_t)n_query * sizeof(int));
  const int sNF = hs.add(nullptr, sizeof(int));
  int rc = PLVI_OK;
  if (hs.commit()) {
    rc = plvi_line_fuse_search(m, 1, hs.ptr<plvi_keyline>(sK), hs.ptr<uint8_t>(sD), hs.ptr<int>(sC), n, hs.ptr<float>(sQ),
                               query_flags ? hs.ptr<uint8_t>(sF) : nullptr, hs.ptr<uint8_t>(sQD), hs.ptr<int>(sC) + 1, n_query, th_low,
                               hs.ptr<int>(sBI), hs.ptr<int>(sBD), hs.ptr<int>(sNF));
    if (rc == PLVI_OK) {
      hs.fetch(best_idx, sBI, (size_t)n_query * sizeof(int));
      hs.fetch(best_dist, sBD, (size_t)n_query * sizeof(int));
      hs.fetch(nfound, sNF, sizeof(int));
    }
  }
  return hs.finish(rc);
}

int plvi_line_match_mad_host(plvi_matcher* m, const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, const uint8_t* has_line1,
                             const uint8_t* has_line2, double factor, int* matches12, int* nmatches, double* mad) {
  if (!m || n1 < 0 || n2 < 0 || !nmatches || !mad || (n1 && (!desc1 || !matches12)) || (n2 && !desc2)) {
    set_error("plvi_line_match_mad_host: invalid argument");
    return PLVI_ERR_INVALID;
  }
  *nmatches = 0; mad[0] = mad[1] = 0.0;
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  if (n1 == 0 || n2 < 2) return PLVI_OK;
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  HostStage hs(m->stream);
  const int cnt[2] = {n1, n2};
  const bool masks = has_line1 && has_line2;
  const int sA = hs.add(desc1, (size_t)n1 * 32), sB = hs.add(desc2, (size_t)n2 * 32), sC = hs.add(cnt, sizeof(cnt));
  const int sH1 = hs.add(masks ? has_line1 : nullptr, (size_t)n1), sH2 = hs.add(masks ? has_line2 : nullptr, (size_t)n2);
  const int sM = hs.add(nullptr, (size_t)n1 * sizeof(int)), sN = hs.add(nullptr, sizeof(int)), sMad = hs.add(nullptr, 2 * sizeof(double));
  int rc = PLVI_OK;
  if (hs.commit()) {
    rc = plvi_line_match_mad(m, 1, hs.ptr<uint8_t>(sA), hs.ptr<int>(sC), n1, hs.ptr<uint8_t>(sB), hs.ptr<int>(sC) + 1, n2,
                             masks ? hs.ptr<uint8_t>(sH1) : nullptr, masks ? hs.ptr<uint8_t>(sH2) : nullptr, factor, hs.ptr<int>(sM),
                             hs.ptr<int>(sN), hs.ptr<double>(sMad));
    if (rc == PLVI_OK) { hs.fetch(matches12, sM, (size_t)n1 * sizeof(int)); hs.fetch(nmatches, sN, sizeof(int)); hs.fetch(mad, sMad, 2 * sizeof(double)); }
  }
  return hs.finish(rc);
}

int plvi_distinctive_descriptors_host(plvi_matcher* m, const uint8_t* desc, const int* counts, int n_points, int stride,
                                      int* best_idx, uint8_t* best_desc) {
  if (!m || !desc || !counts || n_points < 1 || stride < 1 || !best_idx) {
    set_error("plvi_distinctive_descriptors_host: invalid argument");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  HostStage hs(m->stream);
  const int sD = hs.add(desc, (size_t)n_points * stride * 32), sC = hs.add(counts, (size_t)n_points * sizeof(int));
  const int sI = hs.add(nullptr, (size_t)n_points * sizeof(int)), sO = hs.add(nullptr, (size_t)n_points * 32);
  int rc = PLVI_OK;
  if (hs.commit()) {
    rc = plvi_distinctive_descriptors(m->stream, hs.ptr<uint8_t>(sD), hs.ptr<int>(sC), n_points, stride, hs.ptr<int>(sI), hs.ptr<uint8_t>(sO));
    if (rc == PLVI_OK) { hs.fetch(best_idx, sI, (size_t)n_points * sizeof(int)); if (best_desc) hs.fetch(best_desc, sO, (size_t)n_points * 32); }
  }
  return hs.finish(rc);
}

int plvi_line_match_grid_occ_host(plvi_matcher* m, const int* lines1, const uint8_t* desc1, int n1, const uint8_t* occ2,
                                  const double* directions2, const uint8_t* desc2, int n2, int grid_rows, int grid_cols,
                                  int win_left, int win_right, int win_up, int win_down, int* matches12, int* nmatches) {
  if (!m || n1 < 0 || n2 < 0 || !matches12 || !nmatches || (n1 && (!lines1 || !desc1)) || (n2 && (!occ2 || !directions2 || !desc2)) ||
      grid_rows < 1 || grid_cols < 1 || grid_rows > 64 || grid_cols > 64 || win_left < 0 || win_right < 0 || win_up < 0 || win_down < 0) {
    set_error("plvi_line_match_grid_occ_host: invalid argument (grid at most 64 x 64 cells)");
    return PLVI_ERR_INVALID;
  }
  *nmatches = 0;
  for (int i = 0; i < n1; i++) matches12[i] = -1;
  if (n1 == 0 || n2 == 0) return PLVI_OK;
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  const size_t smem = (size_t)n2 * ((size_t)grid_rows * 2 + 2 * sizeof(double) + 32 + 2 * sizeof(int));
  if (smem > 200 * 1024) { set_error("plvi_line_match_grid_occ_host: right line set too large for one CTA"); return PLVI_ERR_CAPACITY; }
  if (smem > 48 * 1024)
    PLVI_CUDA_TRY(cudaFuncSetAttribute(k_line_match_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HostStage hs(m->stream);
  const int cnt[2] = {n1, n2};
  const int sL = hs.add(lines1, (size_t)n1 * 4 * sizeof(int)), sD1 = hs.add(desc1, (size_t)n1 * 32);
  const int sO = hs.add(occ2, (size_t)n2 * grid_rows * 2), sDir = hs.add(directions2, (size_t)n2 * 2 * sizeof(double));
  const int sD2 = hs.add(desc2, (size_t)n2 * 32), sC = hs.add(cnt, sizeof(cnt));
  const int sM = hs.add(nullptr, (size_t)n1 * sizeof(int)), sN = hs.add(nullptr, sizeof(int));
  if (hs.commit()) {
    k_line_match_grid<<<1, 32, smem, m->stream>>>(nullptr, hs.ptr<uint8_t>(sD1), hs.ptr<int>(sC), n1, nullptr, hs.ptr<uint8_t>(sD2),
                                                  hs.ptr<int>(sC) + 1, n2, 1.0, 1.0, grid_rows, grid_cols, win_left, win_right, win_up,
                                                  win_down, hs.ptr<int>(sM), hs.ptr<int>(sN), hs.ptr<uchar2>(sO), hs.ptr<double>(sDir),
                                                  hs.ptr<int>(sL));
    m->lastLaunches = 1;
    hs.e = cudaGetLastError();
    hs.fetch(matches12, sM, (size_t)n1 * sizeof(int));
    hs.fetch(nmatches, sN, sizeof(int));
  }
  return hs.finish(PLVI_OK);
}

int plvi_pair_queries(plvi_matcher* m, const plvi_keypoint* d_kps, const int* d_counts, int npairs, int stride,
                      int out_stride, float th, float scale_factor, const float* affine6, const float* bounds4, float init_window,
                      plvi_query* d_queries_proj, plvi_query* d_queries_init, int* d_qcount, int* d_tcount) {
  if (!m || !d_kps || !d_counts || !affine6 || !bounds4 || !d_queries_proj || !d_qcount || !d_tcount || npairs < 1 || stride < 1 ||
      out_stride < stride) {
    set_error("plvi_pair_queries: invalid argument");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  k_pair_queries<<<dim3((stride + 255) / 256, npairs), 256, 0, m->stream>>>(d_kps, d_counts, stride, out_stride, th, scale_factor, affine6[0],
                                                                            affine6[1], affine6[2], affine6[3], affine6[4], affine6[5],
                                                                            bounds4[0], bounds4[1], bounds4[2], bounds4[3], init_window,
                                                                            d_queries_proj, d_queries_init, d_qcount, d_tcount);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_gather_i32(void* stream, const int* d_src, int n, int first, int step, int* d_dst) {
  if (!d_src || !d_dst || n < 0 || step < 1 || first < 0) { set_error("plvi_gather_i32: invalid argument"); return PLVI_ERR_INVALID; }
  if (n == 0) return PLVI_OK;
  k_gather_i32<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_src, n, first, step, d_dst);
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_line_stereo_depth(plvi_matcher* m, int npairs, const float* d_seg1, const int* d_n1, int stride1, const float* d_seg2,
                           const int* d_n2, int stride2, const int* d_matches12, const float* d_seg1_un, float mbf, float* d_disparity,
                           float* d_depth, double* d_le, int* d_ndepth) {
  if (!m || npairs < 1 || !d_seg1 || !d_n1 || !d_seg2 || !d_n2 || !d_matches12 || !d_disparity || !d_depth || !d_ndepth || stride1 < 1 ||
      stride2 < 1 || (d_le && !d_seg1_un)) {
    set_error("plvi_line_stereo_depth: invalid argument");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  k_line_stereo_depth<<<npairs, 128, 0, m->stream>>>(reinterpret_cast<const float4*>(d_seg1), d_n1, stride1,
                                                     reinterpret_cast<const float4*>(d_seg2), d_n2, stride2, d_matches12,
                                                     reinterpret_cast<const float4*>(d_seg1_un), mbf,
                                                     reinterpret_cast<float2*>(d_disparity), reinterpret_cast<float2*>(d_depth), d_le,
                                                     d_ndepth);
  m->lastLaunches = 1;
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_line_stereo_depth_host(plvi_matcher* m, const float* seg1, int n1, const float* seg2, int n2, const int* matches12,
                                const float* seg1_un, float mbf, float* disparity, float* depth, double* le, int* ndepth) {
  if (!m || n1 < 0 || n2 < 0 || !ndepth || (n1 && (!seg1 || !matches12 || !disparity || !depth)) || (n2 && !seg2) || (n1 && le && !seg1_un)) {
    set_error("plvi_line_stereo_depth_host: invalid argument");
    return PLVI_ERR_INVALID;
  }
  *ndepth = 0;
  if (n1 == 0) return PLVI_OK;
  PLVI_CUDA_TRY(cudaSetDevice(m->device));
  cudaStream_t st = m->stream;
  const int s2 = n2 > 0 ? n2 : 1;
  const auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };   // float4 / float2 / double accesses
  const size_t bS1 = (size_t)n1 * 16, bS2 = (size_t)s2 * 16, bM = (size_t)n1 * 4, bO = (size_t)n1 * 8, bLe = (size_t)n1 * 24;
  const size_t oLe = 0, oS1 = up16(oLe + bLe), oSu = oS1 + bS1, oS2 = oSu + bS1, oM = oS2 + bS2, oDs = up16(oM + bM), oDp = up16(oDs + bO),
               oC = up16(oDp + bO);
  unsigned char* buf = nullptr;
  PLVI_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&buf), oC + 16, st));
  const int cnt[3] = {n1, 0, n2};
  cudaError_t e = cudaMemcpyAsync(buf + oS1, seg1, bS1, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(buf + oSu, le ? seg1_un : seg1, bS1, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n2) e = cudaMemcpyAsync(buf + oS2, seg2, (size_t)n2 * 16, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(buf + oM, matches12, bM, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(buf + oC, cnt, sizeof(cnt), cudaMemcpyHostToDevice, st);
  int rc = PLVI_OK;
  if (e == cudaSuccess) {
    int* dc = reinterpret_cast<int*>(buf + oC);
    rc = plvi_line_stereo_depth(m, 1, reinterpret_cast<const float*>(buf + oS1), dc, n1, reinterpret_cast<const float*>(buf + oS2), dc + 2, s2,
                                reinterpret_cast<const int*>(buf + oM), reinterpret_cast<const float*>(buf + oSu), mbf,
                                reinterpret_cast<float*>(buf + oDs), reinterpret_cast<float*>(buf + oDp),
                                le ? reinterpret_cast<double*>(buf + oLe) : nullptr, dc + 1);
    if (rc == PLVI_OK) {
      e = cudaMemcpyAsync(disparity, buf + oDs, bO, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(depth, buf + oDp, bO, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess && le) e = cudaMemcpyAsync(le, buf + oLe, bLe, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(ndepth, dc + 1, sizeof(int), cudaMemcpyDeviceToHost, st);
    }
  }
  cudaFreeAsync(buf, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return PLVI_ERR_CUDA; }
  return rc;
}

}  // extern "C"
