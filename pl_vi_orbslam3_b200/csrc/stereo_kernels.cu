// Frame::ComputeStereoMatches (src/Frame.cc:1228-1406) on the device-resident pyramids of two ORB extractor handles.
//   k_stereo_match   warp per left keypoint: best right keypoint of its row band by Hamming distance (levels +-1,
//                    disparity range, first smallest distance in right-keypoint order), 11x11 SAD refinement over
//                    +-5 px on the left keypoint's pyramid level (lane = shift), parabola fit, depth.
//   k_stereo_filter  CTA per frame: median of the SAD values by rank counting, matches with SAD >= 1.5 * 1.4 * median
//                    are removed.
// Float expressions use explicit _rn operations in the reference's evaluation order (no contraction).
#include "plvi_internal.cuh"

namespace plvi {

struct StereoArgs {
  OrbPtrs L, R;
  int w[PLVI_MAX_LEVELS], h[PLVI_MAX_LEVELS];
  float scale[PLVI_MAX_LEVELS], invScale[PLVI_MAX_LEVELS];
  int nlevels;
  const plvi_keypoint* kl; const uint8_t* dl; const int* nl;
  const plvi_keypoint* kr; const uint8_t* dr; const int* nr;
  int stride;
  float mb, mbf;
  float* uRight; float* depth; int* sad; int* nstereo;
};

__global__ void __launch_bounds__(256) k_stereo_match(const StereoArgs a) {
  const int f = blockIdx.y, lane = threadIdx.x & 31;
  const int iL = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (iL >= a.stride) return;
  const size_t o = (size_t)f * a.stride + iL;
  float outU = -1.0f, outD = -1.0f;
  int outSad = -1;
  const int nL = min(a.nl[f], a.stride), nR = min(a.nr[f], a.stride);
  // octaves come straight from the caller's keypoint records: a record outside [0, nlevels) is skipped
  if (iL < nL && (unsigned)a.kl[o].octave < (unsigned)a.nlevels) {
    const plvi_keypoint kpL = a.kl[o];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL, nRows = a.h[0];
    const float maxD = __fdiv_rn(a.mbf, a.mb);
    const float minU = __fsub_rn(uL, maxD), maxU = uL;
    if (row >= 0 && row < nRows && !(maxU < 0)) {
      uint32_t q[8];
#pragma unroll
      for (int k = 0; k < 8; k++) q[k] = __ldg(reinterpret_cast<const uint32_t*>(a.dl + o * 32) + k);
      const plvi_keypoint* kr = a.kr + (size_t)f * a.stride;
      const uint8_t* dr = a.dr + (size_t)f * a.stride * 32;
      unsigned best = 0xffffffffu;
      for (int iR = lane; iR < nR; iR += 32) {
        const plvi_keypoint kpR = kr[iR];
        if ((unsigned)kpR.octave >= (unsigned)a.nlevels) continue;
        const float r = __fmul_rn(2.0f, a.scale[kpR.octave]);
        const int maxr = (int)ceilf(__fadd_rn(kpR.y, r)), minr = (int)floorf(__fsub_rn(kpR.y, r));
        if (row < minr || row > maxr) continue;                       // vRowIndices[row] holds iR
        if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
        if (!(kpR.x >= minU && kpR.x <= maxU)) continue;
        const uint32_t* d = reinterpret_cast<const uint32_t*>(dr + (size_t)iR * 32);
        int dist = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) dist += __popc(q[k] ^ __ldg(d + k));
        const unsigned key = ((unsigned)dist << 16) | (unsigned)iR;   // first smallest distance in iR order
        best = min(best, key);
      }
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, s));
      const int bestDist = (int)(best >> 16), bestIdxR = (int)(best & 0xffffu);
      if (best != 0xffffffffu && bestDist < 100 && bestDist < (100 + 50) / 2) {
        const float uR0 = kr[bestIdxR].x;
        const float sfac = a.invScale[levelL];
        const float scaleduL = roundf(__fmul_rn(kpL.x, sfac)), scaledvL = roundf(__fmul_rn(kpL.y, sfac));
        const float scaleduR0 = roundf(__fmul_rn(uR0, sfac));
        const int w = 5, Ls = 5;
        const int W = a.w[levelL], H = a.h[levelL];
        const int cuL = (int)scaleduL, cvL = (int)scaledvL, cuR = (int)scaleduR0;
        const float iniu = __fsub_rn(__fadd_rn(scaleduR0, (float)Ls), (float)w);
        const float endu = __fadd_rn(__fadd_rn(__fadd_rn(scaleduR0, (float)Ls), (float)w), 1.0f);
        const bool inside = cvL - w >= 0 && cvL + w < H && cuL - w >= 0 && cuL + w < W && !(iniu < 0 || endu >= (float)W) &&
                            cuR - Ls - w >= 0;
        if (inside) {
          const u8* IL = a.L.img[levelL] + (size_t)f * a.L.ifs[levelL];
          const u8* IR = a.R.img[levelL] + (size_t)f * a.R.ifs[levelL];
          const int pL = a.L.ipitch[levelL], pR = a.R.ipitch[levelL];
          int sad = 0x7fffffff;
          if (lane < 2 * Ls + 1) {   // lane = shift index: incR = lane - L
            const int incR = lane - Ls;
            const int cL = IL[(size_t)cvL * pL + cuL], cR = IR[(size_t)cvL * pR + cuR + incR];
            sad = 0;
            for (int dy = -w; dy <= w; dy++) {
              const u8* rl = IL + (size_t)(cvL + dy) * pL + cuL - w;
              const u8* rr = IR + (size_t)(cvL + dy) * pR + cuR + incR - w;
#pragma unroll
              for (int dx = 0; dx <= 2 * w; dx++) sad += abs(((int)rl[dx] - cL) - ((int)rr[dx] - cR));
            }
          }
          // first smallest SAD over incR = -L .. L
          int bestSad = 0x7fffffff, bestinc = 0;
          float d1 = 0.f, d2 = 0.f, d3 = 0.f;
          int sads[2 * 5 + 1];
#pragma unroll
          for (int k = 0; k <= 2 * Ls; k++) sads[k] = __shfl_sync(0xffffffffu, sad, k);
#pragma unroll
          for (int k = 0; k <= 2 * Ls; k++)
            if (sads[k] < bestSad) { bestSad = sads[k]; bestinc = k - Ls; }
          if (bestinc != -Ls && bestinc != Ls) {
#pragma unroll
            for (int k = 1; k < 2 * Ls; k++)
              if (k - Ls == bestinc) { d1 = (float)sads[k - 1]; d2 = (float)sads[k]; d3 = (float)sads[k + 1]; }
            const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2)));
            const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), den);
            if (!(deltaR < -1 || deltaR > 1)) {
              float bestuR = __fmul_rn(a.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
              float disparity = __fsub_rn(uL, bestuR);
              if (disparity >= 0 && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                outD = __fdiv_rn(a.mbf, disparity);
                outU = bestuR;
                outSad = bestSad;
              }
            }
          }
        }
      }
    }
  }
  if (lane == 0) { a.uRight[o] = outU; a.depth[o] = outD; a.sad[o] = outSad; }
}

__global__ void __launch_bounds__(256) k_stereo_filter(const StereoArgs a) {
  const int f = blockIdx.x, tid = threadIdx.x;
  const int nL = min(a.nl[f], a.stride);
  const int* sad = a.sad + (size_t)f * a.stride;
  __shared__ int s_cnt, s_median, s_kept;
  if (tid == 0) { s_cnt = 0; s_median = -1; s_kept = 0; }
  __syncthreads();
  int c = 0;
  for (int i = tid; i < nL; i += 256) c += sad[i] >= 0;
  if (c) atomicAdd(&s_cnt, c);
  __syncthreads();
  const int cnt = s_cnt;
  if (cnt == 0) { if (tid == 0) a.nstereo[f] = 0; return; }
  // sorted pair list (sad, iL): the element of rank cnt / 2 is the median
  for (int i = tid; i < nL; i += 256) {
    const int v = sad[i];
    if (v < 0) continue;
    int rank = 0;
    for (int j = 0; j < nL; j++) {
      const int u = sad[j];
      rank += (u >= 0) && (u < v || (u == v && j < i));
    }
    if (rank == cnt / 2) s_median = v;
  }
  __syncthreads();
  const float thDist = __fmul_rn(1.5f * 1.4f, (float)s_median);
  int kept = 0;
  for (int i = tid; i < nL; i += 256) {
    const int v = sad[i];
    if (v < 0) continue;
    if ((float)v < thDist) kept++;
    else { a.uRight[(size_t)f * a.stride + i] = -1.0f; a.depth[(size_t)f * a.stride + i] = -1.0f; }
  }
  if (kept) atomicAdd(&s_kept, kept);
  __syncthreads();
  if (tid == 0) a.nstereo[f] = s_kept;
}

int launch_stereo(const OrbGeom& g, const OrbPtrs& L, const OrbPtrs& R, const float* scale, const float* invScale, int n,
                  const plvi_keypoint* kl, const uint8_t* dl, const int* nl, const plvi_keypoint* kr, const uint8_t* dr,
                  const int* nr, int stride, float mb, float mbf, float* uRight, float* depth, int* sad, int* nstereo,
                  cudaStream_t st) {
  StereoArgs a;
  a.L = L; a.R = R;
  a.nlevels = g.nlevels;
  for (int l = 0; l < g.nlevels; l++) { a.w[l] = g.lv[l].w; a.h[l] = g.lv[l].h; a.scale[l] = scale[l]; a.invScale[l] = invScale[l]; }
  a.kl = kl; a.dl = dl; a.nl = nl; a.kr = kr; a.dr = dr; a.nr = nr; a.stride = stride;
  a.mb = mb; a.mbf = mbf;
  a.uRight = uRight; a.depth = depth; a.sad = sad; a.nstereo = nstereo;
  k_stereo_match<<<dim3((stride + 7) / 8, n), 256, 0, st>>>(a);
  k_stereo_filter<<<n, 256, 0, st>>>(a);
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

}  // namespace plvi
