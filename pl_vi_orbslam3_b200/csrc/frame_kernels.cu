// Frame post-processing that follows the extractors:
//   k_undistort       cv::undistortPoints as called by Frame::UndistortKeyPoints / UndistortKeyLines
//                     (src/Frame.cc:1124-1197): normalise with K, 5 fixed-point iterations of the
//                     plumb-bob model in double, re-project with P = K; float in / float out
//   k_assign_grid     Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:644-675,1077-1087):
//                     64 x 48 cell lists in keypoint order, written as a CSR (cell_start, items)
// All pointers are device pointers; nothing here allocates.
#include "plvi_internal.cuh"

namespace plvi {

struct CamArgs {
  double fx, fy, cx, cy, ifx, ify;
  double k[14];
  double nfx, nfy, ncx, ncy;
  int iters;
  int identity;   // mDistCoef[0] == 0: the reference copies the input
};

__device__ __forceinline__ float2 undistort_one(float u, float v, const CamArgs& c) {
  if (c.identity) return make_float2(u, v);
  // explicit _rn operations: the reference (OpenCV, -ffp-contract off on x86) does not fuse
  double x = __dmul_rn(__dsub_rn((double)u, c.cx), c.ifx), y = __dmul_rn(__dsub_rn((double)v, c.cy), c.ify);
  const double x0 = x, y0 = y;
  for (int j = 0; j < c.iters; j++) {
    const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
    const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k[7], r2), c.k[6]), r2), c.k[5]), r2));
    const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k[4], r2), c.k[1]), r2), c.k[0]), r2));
    const double icdist = __ddiv_rn(num, den);
    if (icdist < 0) {
      x = __dmul_rn(__dsub_rn((double)u, c.cx), c.ifx);
      y = __dmul_rn(__dsub_rn((double)v, c.cy), c.ify);
      break;
    }
    // deltaX = 2*k2*x*y + k3*(r2 + 2*x*x) + k8*r2 + k9*r2*r2 (left to right, as written in OpenCV)
    double dX = __dmul_rn(__dmul_rn(__dmul_rn(2.0, c.k[2]), x), y);
    dX = __dadd_rn(dX, __dmul_rn(c.k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
    dX = __dadd_rn(dX, __dmul_rn(c.k[8], r2));
    dX = __dadd_rn(dX, __dmul_rn(__dmul_rn(c.k[9], r2), r2));
    double dY = __dmul_rn(c.k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y)));
    dY = __dadd_rn(dY, __dmul_rn(__dmul_rn(__dmul_rn(2.0, c.k[3]), x), y));
    dY = __dadd_rn(dY, __dmul_rn(c.k[10], r2));
    dY = __dadd_rn(dY, __dmul_rn(__dmul_rn(c.k[11], r2), r2));
    x = __dmul_rn(__dsub_rn(x0, dX), icdist);
    y = __dmul_rn(__dsub_rn(y0, dY), icdist);
  }
  // P = [nfx 0 ncx; 0 nfy ncy; 0 0 1]: xx = nfx*x + 0*y + ncx, ww = 1/(0*x + 0*y + 1)
  const double xx = __dadd_rn(__dadd_rn(__dmul_rn(c.nfx, x), __dmul_rn(0.0, y)), c.ncx);
  const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(c.nfy, y)), c.ncy);
  const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
  return make_float2((float)__dmul_rn(xx, ww), (float)__dmul_rn(yy, ww));
}

// Generic record walker: `npts` (x, y) float pairs at byte offsets off[] inside records of
// `recBytes` bytes; the rest of the record is copied.
template <int REC_WORDS, int NPTS>
__global__ void __launch_bounds__(256) k_undistort(const uint32_t* __restrict__ in, const int* __restrict__ counts, int stride,
                                                   uint32_t* __restrict__ out, const __grid_constant__ CamArgs cam, int off0,
                                                   int off1) {
  const int f = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= counts[f]) return;
  const uint32_t* r = in + ((size_t)f * stride + i) * REC_WORDS;
  uint32_t w[REC_WORDS];
#pragma unroll
  for (int k = 0; k < REC_WORDS; k++) w[k] = r[k];
  const int offs[2] = {off0, off1};
#pragma unroll
  for (int p = 0; p < NPTS; p++) {
    const float2 q = undistort_one(__uint_as_float(w[offs[p]]), __uint_as_float(w[offs[p] + 1]), cam);
    w[offs[p]] = __float_as_uint(q.x);
    w[offs[p] + 1] = __float_as_uint(q.y);
  }
  uint32_t* o = out + ((size_t)f * stride + i) * REC_WORDS;
#pragma unroll
  for (int k = 0; k < REC_WORDS; k++) o[k] = w[k];
}

#define FG_COLS 64
#define FG_ROWS 48
#define FG_CELLS (FG_COLS * FG_ROWS)

__global__ void __launch_bounds__(256) k_assign_grid(const plvi_keypoint* __restrict__ keysAll, const int* __restrict__ counts,
                                                     int stride, plvi_grid grid, int* __restrict__ cellStartAll,
                                                     int* __restrict__ itemsAll) {
  __shared__ int cnt[FG_CELLS];
  __shared__ int wsum[8];
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = min(counts[f], stride);
  const plvi_keypoint* keys = keysAll + (size_t)f * stride;
  int* cellStart = cellStartAll + (size_t)f * (FG_CELLS + 1);
  int* items = itemsAll + (size_t)f * stride;
  for (int i = tid; i < FG_CELLS; i += 256) cnt[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += 256) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(keys[i].x, grid.min_x), grid.inv_w));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(keys[i].y, grid.min_y), grid.inv_h));
    if (px >= 0 && px < FG_COLS && py >= 0 && py < FG_ROWS) atomicAdd(&cnt[px * FG_ROWS + py], 1);
  }
  __syncthreads();
  {  // exclusive scan, 12 cells per thread
    const int beg = tid * 12, end = beg + 12;
    int sum = 0;
    for (int i = beg; i < end; i++) sum += cnt[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    int base = incl - sum;
    for (int k = 0; k < wid; k++) base += wsum[k];
    for (int i = beg; i < end; i++) {
      const int c = cnt[i];
      cellStart[i] = base;
      cnt[i] = base;      // becomes the insertion cursor
      base += c;
    }
    if (tid == 255) cellStart[FG_CELLS] = base;
  }
  __syncthreads();
  for (int i = tid; i < n; i += 256) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(keys[i].x, grid.min_x), grid.inv_w));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(keys[i].y, grid.min_y), grid.inv_h));
    if (px >= 0 && px < FG_COLS && py >= 0 && py < FG_ROWS) items[atomicAdd(&cnt[px * FG_ROWS + py], 1)] = i;
  }
  __syncthreads();
  // insertion (index) order inside each cell, as push_back in a loop over i produces it
  for (int c = tid; c < FG_CELLS; c += 256) {
    const int s = cellStart[c], e = cnt[c];
    for (int i = s + 1; i < e; i++) {
      const int v = items[i];
      int j = i - 1;
      while (j >= s && items[j] > v) { items[j + 1] = items[j]; j--; }
      items[j + 1] = v;
    }
  }
}

static int make_cam(const plvi_camera* cam, CamArgs& c) {
  if (!cam || cam->fx == 0 || cam->fy == 0) { set_error("plvi_camera: null or zero focal length"); return PLVI_ERR_INVALID; }
  c.fx = cam->fx; c.fy = cam->fy; c.cx = cam->cx; c.cy = cam->cy;
  c.ifx = 1.0 / cam->fx; c.ify = 1.0 / cam->fy;
  for (int i = 0; i < 14; i++) c.k[i] = cam->dist[i];
  c.nfx = cam->new_fx; c.nfy = cam->new_fy; c.ncx = cam->new_cx; c.ncy = cam->new_cy;
  c.iters = cam->iters > 0 ? cam->iters : 5;   // cv::undistortPoints default TermCriteria(COUNT, 5, 0.01)
  c.identity = cam->dist[0] == 0.0;            // Frame::UndistortKeyPoints: mDistCoef.at<float>(0) == 0.0
  return PLVI_OK;
}

}  // namespace plvi

using namespace plvi;

extern "C" {

int plvi_undistort_keypoints(void* stream, const plvi_keypoint* d_in, const int* d_counts, int n_frames, int stride,
                             const plvi_camera* cam, plvi_keypoint* d_out) {
  if (!d_in || !d_out || !d_counts || n_frames < 1 || stride < 1) { set_error("plvi_undistort_keypoints: invalid argument"); return PLVI_ERR_INVALID; }
  CamArgs c;
  int rc = make_cam(cam, c);
  if (rc) return rc;
  static_assert(sizeof(plvi_keypoint) == 28, "keypoint layout");
  k_undistort<7, 1><<<dim3((stride + 255) / 256, n_frames), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint32_t*>(d_in), d_counts, stride, reinterpret_cast<uint32_t*>(d_out), c, 0, 0);
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_undistort_keylines(void* stream, const plvi_keyline* d_in, const int* d_counts, int n_frames, int stride,
                            const plvi_camera* cam, plvi_keyline* d_out) {
  if (!d_in || !d_out || !d_counts || n_frames < 1 || stride < 1) { set_error("plvi_undistort_keylines: invalid argument"); return PLVI_ERR_INVALID; }
  CamArgs c;
  int rc = make_cam(cam, c);
  if (rc) return rc;
  static_assert(sizeof(plvi_keyline) == 68, "keyline layout");
  // startPointX/Y are words 7-8, endPointX/Y words 9-10 of the 17-word record
  k_undistort<17, 2><<<dim3((stride + 255) / 256, n_frames), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint32_t*>(d_in), d_counts, stride, reinterpret_cast<uint32_t*>(d_out), c, 7, 9);
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

int plvi_assign_features_to_grid(void* stream, const plvi_keypoint* d_keys, const int* d_counts, int n_frames, int stride,
                                 const plvi_grid* grid, int* d_cell_start, int* d_cell_items) {
  if (!d_keys || !d_counts || !grid || !d_cell_start || !d_cell_items || n_frames < 1 || stride < 1) {
    set_error("plvi_assign_features_to_grid: invalid argument");
    return PLVI_ERR_INVALID;
  }
  k_assign_grid<<<n_frames, 256, 0, (cudaStream_t)stream>>>(d_keys, d_counts, stride, *grid, d_cell_start, d_cell_items);
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}

}  // extern "C"
