// C ABI for the ORB extractor (include/plvi.h): handle lifetime, geometry planning
// (ORBextractor ctor, src/ORBextractor.cc:408-468; cell grid :771-785; octree roots
// :541-543), device scratch, H2D/D2H staging.
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include <algorithm>

#include "plvi_internal.cuh"

namespace plvi {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

static inline int round_even_f(float v) { return (int)nearbyintf(v); }

}  // namespace plvi

using namespace plvi;

struct plvi_orb {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0;
  float scaleFactor = 1.f;
  int maxW = 0, maxH = 0, maxBatch = 0;
  std::vector<float> scale, invScale, sigma2, invSigma2;
  std::vector<int> quota;
  int curW = -1, curH = -1;
  OrbGeom geom;        // geometry of the current (w,h)
  OrbGeom capGeom;     // geometry of (maxW,maxH): allocation capacities
  size_t lvlCapBytes[PLVI_MAX_LEVELS] = {0};  // per-frame bytes per level at capacity
  u8* dImg[PLVI_MAX_LEVELS] = {nullptr};
  AsyncInput ain;   // input staging of the host-buffer entry points
  u8* dBlur[PLVI_MAX_LEVELS] = {nullptr};
  OrbScratch scr = {};
  int2* dRsTab = nullptr;
  FastTile* dFastTiles = nullptr;
  BlurTile* dBlurTiles = nullptr;
  size_t rsCap = 0, fastTileCap = 0, blurTileCap = 0;
  plvi_keypoint* dKps = nullptr;     // result set 0 (device copies of the host-buffer entry points' results)
  uint8_t* dDesc = nullptr;
  int* dCounts = nullptr;
  int* dMono = nullptr;
  plvi_keypoint* dKpsB = nullptr;    // result set 1, allocated by the second host-buffer call
  uint8_t* dDescB = nullptr;
  int* dCountsB = nullptr;
  int* dMonoB = nullptr;
  AsyncOutput aout;
  int cap = 0;
  int lastN = 0, lastLaunches = 0;
  cudaEvent_t waitAfterPyramid = nullptr;   // one-shot (plvi_orb_wait_event_after_pyramid)
  bool pyrValid = false;                    // plvi_orb_pyramid_device built the pyramid of pyrKey ahead of the extraction call
  uint64_t pyrKey[6] = {0, 0, 0, 0, 0, 0};
  StageProf prof;
  GraphCache graphs;
  int* dStereoSad = nullptr;      // scratch of plvi_orb_stereo_matches
  size_t stereoSadCap = 0;
  cudaEvent_t stereoEv = nullptr;
  std::string profText;
  OrbPtrs lastPtrs = {};
};

namespace {

struct HostTables {
  std::vector<int2> rs;
  std::vector<FastTile> fast;
  std::vector<BlurTile> blur;
};

void linear_rows(int ssize, int dsize, std::vector<int2>& out) {
  const double scale = 1.0 / ((double)dsize / ssize);
  for (int d = 0; d < dsize; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    const int a0 = round_even_f((1.f - f) * 2048.f), a1 = round_even_f(f * 2048.f);
    out.push_back(make_int2(s, (a0 & 0xffff) | (a1 << 16)));
  }
}

// Fills geometry for a w x h input.  Returns PLVI_OK or PLVI_ERR_INVALID when the
// reference itself would be undefined (a level too small for one 30 px cell, or a
// portrait aspect that gives zero octree roots).
int make_geom(const plvi_orb* h, int w, int hh, OrbGeom& g, HostTables* tab) {
  memset(&g, 0, sizeof(g));
  g.nlevels = h->nlevels;
  g.iniTh = h->iniTh;
  g.minTh = h->minTh;
  int candOff = 0, kpOff = 0, rsOff = 0, maxNodes = kMaxNodesMin;
  for (int l = 0; l < h->nlevels; l++) {
    OrbLevel& L = g.lv[l];
    L.w = round_even_f((float)w * h->invScale[l]);
    L.h = round_even_f((float)hh * h->invScale[l]);
    L.pitch = (L.w + 63) & ~63;
    const float width = (float)(L.w - 2 * kEdge), height = (float)(L.h - 2 * kEdge);
    L.nCols = (int)(width / 30.f);
    L.nRows = (int)(height / 30.f);
    if (L.nCols < 1 || L.nRows < 1) {
      set_error("image too small: a pyramid level has no 30 px FAST cell");
      return PLVI_ERR_INVALID;
    }
    L.wCell = (int)ceilf(width / L.nCols);
    L.hCell = (int)ceilf(height / L.nRows);
    if (L.wCell > 64 || L.hCell > 64 || L.nCols > 255 || L.nRows > 255 || L.w > 4000 || L.h > 4000) {
      set_error("image dimensions outside supported range");
      return PLVI_ERR_INVALID;
    }
    L.quota = h->quota[l];
    L.nIni = (int)roundf((float)(L.w - 2 * kEdge) / (float)(L.h - 2 * kEdge));
    if (L.nIni < 1) {
      set_error("portrait aspect ratio: DistributeOctTree has zero root nodes (undefined in the reference)");
      return PLVI_ERR_INVALID;
    }
    L.hX = (float)(L.w - 2 * kEdge) / (float)L.nIni;
    L.scale = h->scale[l];
    L.sizeField = (int)(31 * h->scale[l]);
    L.kpOff = kpOff;
    L.kpCap = std::max(L.quota + 3, 4 * L.nIni);
    kpOff += L.kpCap;
    maxNodes = std::max(maxNodes, L.kpCap + 8);
    L.candOff = candOff;
    // NMS leaves at most ceil(w/2)*ceil(h/2) survivors per cell; +25% so that a smaller
    // image's slightly different cell rounding still fits the capacity layout
    L.candCap = (L.nCols * L.nRows * ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2)) * 5 / 4;
    candOff += (L.candCap + 3) & ~3;
    L.rsOff = rsOff;
    if (l > 0) {
      rsOff += L.w + L.h;
      if (tab) {
        linear_rows(g.lv[l - 1].w, L.w, tab->rs);
        linear_rows(g.lv[l - 1].h, L.h, tab->rs);
      }
    }
    if (tab) {
      for (int r = 0; r < L.nRows; r++)
        for (int c = 0; c < L.nCols; c += 4)
          tab->fast.push_back(FastTile{(unsigned short)l, (unsigned short)r, (unsigned short)c,
                                       (unsigned short)std::min(4, L.nCols - c)});
      for (int ty = 0; ty < (L.h + 31) / 32; ty++)
        for (int tx = 0; tx < (L.w + 127) / 128; tx++)
          tab->blur.push_back(BlurTile{(unsigned short)l, (unsigned short)tx, (unsigned short)ty, 0});
    }
  }
  g.candTotal = candOff;
  g.kpTotal = kpOff;
  g.maxNodes = maxNodes;
  return PLVI_OK;
}

int ensure_geom(plvi_orb* h, int w, int hh) {
  if (w == h->curW && hh == h->curH) return PLVI_OK;
  if (w > h->maxW || hh > h->maxH) {
    set_error("image larger than the handle's max_width/max_height");
    return PLVI_ERR_CAPACITY;
  }
  HostTables tab;
  OrbGeom g;
  int rc = make_geom(h, w, hh, g, &tab);
  if (rc) return rc;
  if (tab.rs.size() > h->rsCap || tab.fast.size() > h->fastTileCap || tab.blur.size() > h->blurTileCap ||
      g.candTotal > h->capGeom.candTotal || g.kpTotal > h->capGeom.kpTotal) {
    set_error("internal: geometry exceeds allocated capacity");
    return PLVI_ERR_CAPACITY;
  }
  // tables are read by kernels of earlier batches: drain the stream first
  PLVI_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->graphs.clear();   // captured graphs hold the old geometry
  PLVI_CUDA_TRY(cudaMemcpy(h->dRsTab, tab.rs.data(), tab.rs.size() * sizeof(int2), cudaMemcpyHostToDevice));
  PLVI_CUDA_TRY(cudaMemcpy(h->dFastTiles, tab.fast.data(), tab.fast.size() * sizeof(FastTile), cudaMemcpyHostToDevice));
  PLVI_CUDA_TRY(cudaMemcpy(h->dBlurTiles, tab.blur.data(), tab.blur.size() * sizeof(BlurTile), cudaMemcpyHostToDevice));
  g.candTotal = h->capGeom.candTotal;  // keep the allocated per-frame strides
  g.kpTotal = h->capGeom.kpTotal;
  // per-level offsets must follow the capacity layout so frames never overlap
  for (int l = 0; l < g.nlevels; l++) {
    g.lv[l].candOff = h->capGeom.lv[l].candOff;
    g.lv[l].kpOff = h->capGeom.lv[l].kpOff;
    g.lv[l].kpCap = std::min(g.lv[l].kpCap, h->capGeom.lv[l].kpCap);
    g.lv[l].candCap = std::min(g.lv[l].candCap, h->capGeom.lv[l].candCap);
  }
  g.maxNodes = h->capGeom.maxNodes;
  h->geom = g;
  h->scr.rsTab = h->dRsTab;
  h->scr.fastTiles = h->dFastTiles;
  h->scr.nFastTiles = (int)tab.fast.size();
  h->scr.blurTiles = h->dBlurTiles;
  h->scr.nBlurTiles = (int)tab.blur.size();
  rc = orb_kernel_attrs(h->geom, &h->scr.fastSmem, &h->scr.octSmem);
  if (rc) return rc;
  h->curW = w;
  h->curH = hh;
  return PLVI_OK;
}

void fill_ptrs(plvi_orb* h, const u8* l0, int l0pitch, size_t l0fs, OrbPtrs& p) {
  memset(&p, 0, sizeof(p));
  for (int l = 0; l < h->nlevels; l++) {
    const OrbLevel& L = h->geom.lv[l];
    p.img[l] = h->dImg[l];
    p.ipitch[l] = L.pitch;
    p.ifs[l] = (size_t)L.pitch * L.h;
    p.blur[l] = h->dBlur[l];
    p.bfs[l] = (size_t)L.pitch * L.h;
  }
  if (l0) {
    p.img[0] = l0;
    p.ipitch[0] = l0pitch;
    p.ifs[0] = l0fs;
  }
}

int check_batch(plvi_orb* h, const void* imgs, int n, int w, int hh, int stride) {
  if (!h) { set_error("null handle"); return PLVI_ERR_INVALID; }
  if (!imgs || w <= 0 || hh <= 0 || n <= 0) { set_error("empty image batch"); return PLVI_ERR_EMPTY; }
  if (stride < w) { set_error("stride < width"); return PLVI_ERR_INVALID; }
  if (n > h->maxBatch) { set_error("batch larger than max_batch"); return PLVI_ERR_CAPACITY; }
  return PLVI_OK;
}

}  // namespace

extern "C" {

const char* plvi_last_error(void) { return g_err.c_str(); }

int plvi_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return PLVI_ERR_CUDA;
  }
  return n;
}

int plvi_orb_create(plvi_orb** out, int nfeatures, float scale_factor, int nlevels, int ini_th,
                    int min_th, int max_width, int max_height, int max_batch, int device,
                    void* stream) {
  if (!out || nfeatures < 1 || nlevels < 1 || nlevels > PLVI_MAX_LEVELS || !(scale_factor > 1.0f) ||
      ini_th < 1 || min_th < 1 || ini_th > 255 || min_th > ini_th || max_batch < 1 || max_width < 1 ||
      max_height < 1) {
    set_error("plvi_orb_create: invalid argument");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(device));
  plvi_orb* h = new plvi_orb();
  h->device = device;
  h->nfeatures = nfeatures;
  h->scaleFactor = scale_factor;
  h->nlevels = nlevels;
  h->iniTh = ini_th;
  h->minTh = min_th;
  h->maxW = max_width;
  h->maxH = max_height;
  h->maxBatch = max_batch;
  // scale tables and per-level quotas, float32 like the reference ctor
  h->scale.assign(nlevels, 1.f);
  h->sigma2.assign(nlevels, 1.f);
  for (int i = 1; i < nlevels; i++) {
    h->scale[i] = h->scale[i - 1] * scale_factor;
    h->sigma2[i] = h->scale[i] * h->scale[i];
  }
  h->invScale.resize(nlevels);
  h->invSigma2.resize(nlevels);
  for (int i = 0; i < nlevels; i++) {
    h->invScale[i] = 1.0f / h->scale[i];
    h->invSigma2[i] = 1.0f / h->sigma2[i];
  }
  h->quota.resize(nlevels);
  {
    const float factor = 1.0f / scale_factor;
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
      h->quota[l] = round_even_f(nDesired);
      sum += h->quota[l];
      nDesired *= factor;
    }
    h->quota[nlevels - 1] = std::max(nfeatures - sum, 0);
  }
  HostTables tab;
  int rc = make_geom(h, max_width, max_height, h->capGeom, &tab);
  if (rc) { delete h; return rc; }
  h->cap = h->capGeom.kpTotal;
  if (stream) {
    h->stream = (cudaStream_t)stream;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); delete h; return PLVI_ERR_CUDA; }
    h->ownStream = true;
  }
  const size_t B = (size_t)max_batch;
  auto fail = [&](cudaError_t e) {
    set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
    plvi_orb_destroy(h);
    return PLVI_ERR_CUDA;
  };
  cudaError_t e;
  for (int l = 0; l < nlevels; l++) {
    const OrbLevel& L = h->capGeom.lv[l];
    h->lvlCapBytes[l] = (size_t)L.pitch * L.h;
    // +256: k_blur7 / k_resize store whole 32-bit words inside the row pitch only, but keep slack
    if ((e = cudaMalloc(&h->dImg[l], B * h->lvlCapBytes[l] + 256)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&h->dBlur[l], B * h->lvlCapBytes[l] + 256)) != cudaSuccess) return fail(e);
  }
  // slack: smaller images may lay tables out with more rows per byte than the max geometry
  h->rsCap = tab.rs.size() + 64 * nlevels;
  h->fastTileCap = tab.fast.size() * 2 + 64;
  h->blurTileCap = tab.blur.size() * 2 + 64;
  if ((e = cudaMalloc(&h->dRsTab, h->rsCap * sizeof(int2))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->dFastTiles, h->fastTileCap * sizeof(FastTile))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->dBlurTiles, h->blurTileCap * sizeof(BlurTile))) != cudaSuccess) return fail(e);
  const size_t ct = (size_t)h->capGeom.candTotal, kt = (size_t)h->capGeom.kpTotal;
  if ((e = cudaMalloc(&h->scr.cand, B * ct * sizeof(uint32_t))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->scr.knode, B * ct * sizeof(uint16_t))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->scr.candCount, B * nlevels * sizeof(int))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->scr.lvlKp, B * kt * sizeof(uint32_t))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->scr.lvlCount, B * nlevels * sizeof(int))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->scr.slot, B * kt * sizeof(int))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->dKps, B * kt * sizeof(plvi_keypoint))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->dDesc, B * kt * 32)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->dCounts, B * sizeof(int))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->dMono, B * sizeof(int))) != cudaSuccess) return fail(e);
  *out = h;
  return PLVI_OK;
}

void plvi_orb_destroy(plvi_orb* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->ain.destroy();
  for (int l = 0; l < PLVI_MAX_LEVELS; l++) {
    cudaFree(h->dImg[l]);
    cudaFree(h->dBlur[l]);
  }
  cudaFree(h->dRsTab);
  if (h->dStereoSad) cudaFree(h->dStereoSad);
  if (h->stereoEv) cudaEventDestroy(h->stereoEv);
  cudaFree(h->dFastTiles);
  cudaFree(h->dBlurTiles);
  cudaFree(h->scr.cand);
  cudaFree(h->scr.knode);
  cudaFree(h->scr.candCount);
  cudaFree(h->scr.lvlKp);
  cudaFree(h->scr.lvlCount);
  cudaFree(h->scr.slot);
  h->aout.destroy();
  cudaFree(h->dKps);
  cudaFree(h->dDesc);
  cudaFree(h->dCounts);
  cudaFree(h->dMono);
  cudaFree(h->dKpsB);
  cudaFree(h->dDescB);
  cudaFree(h->dCountsB);
  cudaFree(h->dMonoB);
  if (h->ownStream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int plvi_orb_capacity(const plvi_orb* h) { return h ? h->cap : PLVI_ERR_INVALID; }
int plvi_orb_levels(const plvi_orb* h) { return h ? h->nlevels : PLVI_ERR_INVALID; }
float plvi_orb_scale_factor(const plvi_orb* h) { return h ? h->scaleFactor : 0.f; }
void* plvi_orb_stream(const plvi_orb* h) { return h ? (void*)h->stream : nullptr; }
int plvi_orb_last_launches(const plvi_orb* h) { return h ? h->lastLaunches : PLVI_ERR_INVALID; }
int plvi_orb_stereo_matches(plvi_orb* left, plvi_orb* right, int n, const plvi_keypoint* d_kps_l, const uint8_t* d_desc_l,
                            const int* d_counts_l, const plvi_keypoint* d_kps_r, const uint8_t* d_desc_r, const int* d_counts_r,
                            int stride, float mb, float mbf, float* d_u_right, float* d_depth, int* d_nstereo) {
  if (!left || !right || n < 1 || !d_kps_l || !d_desc_l || !d_counts_l || !d_kps_r || !d_desc_r || !d_counts_r || stride < 1 ||
      stride > 65535 || !(mb > 0) || !d_u_right || !d_depth || !d_nstereo) {
    set_error("plvi_orb_stereo_matches: invalid argument");
    return PLVI_ERR_INVALID;
  }
  if (left->device != right->device || left->nlevels != right->nlevels || left->curW != right->curW || left->curH != right->curH ||
      left->curW < 0 || n > left->lastN || n > right->lastN || left->scaleFactor != right->scaleFactor) {
    set_error("plvi_orb_stereo_matches: the two extractors must hold the pyramids of the same batch geometry");
    return PLVI_ERR_INVALID;
  }
  PLVI_CUDA_TRY(cudaSetDevice(left->device));
  const size_t need = (size_t)n * stride;
  if (need > left->stereoSadCap) {
    PLVI_CUDA_TRY(cudaStreamSynchronize(left->stream));
    if (left->dStereoSad) cudaFree(left->dStereoSad);
    left->dStereoSad = nullptr; left->stereoSadCap = 0;
    PLVI_CUDA_TRY(cudaMalloc((void**)&left->dStereoSad, need * sizeof(int)));
    left->stereoSadCap = need;
  }
  if (!left->stereoEv) PLVI_CUDA_TRY(cudaEventCreateWithFlags(&left->stereoEv, cudaEventDisableTiming));
  // the right pyramid is produced on the right handle's stream
  PLVI_CUDA_TRY(cudaEventRecord(left->stereoEv, right->stream));
  PLVI_CUDA_TRY(cudaStreamWaitEvent(left->stream, left->stereoEv, 0));
  return launch_stereo(left->geom, left->lastPtrs, right->lastPtrs, left->scale.data(), left->invScale.data(), n, d_kps_l, d_desc_l,
                       d_counts_l, d_kps_r, d_desc_r, d_counts_r, stride, mb, mbf, d_u_right, d_depth, left->dStereoSad, d_nstereo,
                       left->stream);
}

// Host-pointer form for one stereo frame (the reference's Frame::ComputeStereoMatches works on the host vectors
// mvKeys / mvKeysRight / mDescriptors / mDescriptorsRight and fills mvuRight / mvDepth): staged through device
// scratch owned by the left handle, synchronous.
int plvi_orb_stereo_matches_host(plvi_orb* left, plvi_orb* right, const plvi_keypoint* kps_l, const uint8_t* desc_l, int n_l,
                                 const plvi_keypoint* kps_r, const uint8_t* desc_r, int n_r, float mb, float mbf, float* u_right,
                                 float* depth, int* nstereo) {
  if (!left || !right || n_l < 0 || n_r < 0 || (n_l && (!kps_l || !desc_l || !u_right || !depth)) || (n_r && (!kps_r || !desc_r))) {
    set_error("plvi_orb_stereo_matches_host: invalid argument");
    return PLVI_ERR_INVALID;
  }
  if (nstereo) *nstereo = 0;
  if (n_l == 0) return PLVI_OK;
  PLVI_CUDA_TRY(cudaSetDevice(left->device));
  const int stride = std::max(std::max(n_l, n_r), 1);
  const size_t kb = (size_t)stride * sizeof(plvi_keypoint), db = (size_t)stride * 32;
  const size_t total = 2 * kb + 2 * db + 2 * (size_t)stride * sizeof(float) + 4 * sizeof(int) + 256;
  char* buf = nullptr;
  PLVI_CUDA_TRY(cudaMalloc((void**)&buf, total));
  char* p = buf;
  auto take = [&](size_t n) { char* r = p; p += (n + 15) & ~(size_t)15; return r; };
  plvi_keypoint* dkl = (plvi_keypoint*)take(kb); plvi_keypoint* dkr = (plvi_keypoint*)take(kb);
  uint8_t* ddl = (uint8_t*)take(db); uint8_t* ddr = (uint8_t*)take(db);
  float* dur = (float*)take((size_t)stride * sizeof(float)); float* ddp = (float*)take((size_t)stride * sizeof(float));
  int* dcnt = (int*)take(4 * sizeof(int));
  const int cnt[3] = {n_l, n_r, 0};
  cudaStream_t st = left->stream;
  int rc = PLVI_OK;
  cudaError_t e = cudaMemcpyAsync(dkl, kps_l, (size_t)n_l * sizeof(plvi_keypoint), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ddl, desc_l, (size_t)n_l * 32, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_r) e = cudaMemcpyAsync(dkr, kps_r, (size_t)n_r * sizeof(plvi_keypoint), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n_r) e = cudaMemcpyAsync(ddr, desc_r, (size_t)n_r * 32, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dcnt, cnt, sizeof(cnt), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    rc = plvi_orb_stereo_matches(left, right, 1, dkl, ddl, dcnt, dkr, ddr, dcnt + 1, stride, mb, mbf, dur, ddp, dcnt + 2);
  int ns = 0;
  if (e == cudaSuccess && rc == PLVI_OK) e = cudaMemcpyAsync(u_right, dur, (size_t)n_l * sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && rc == PLVI_OK) e = cudaMemcpyAsync(depth, ddp, (size_t)n_l * sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && rc == PLVI_OK) e = cudaMemcpyAsync(&ns, dcnt + 2, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(buf);
  if (rc != PLVI_OK) return rc;
  PLVI_CUDA_TRY(e);
  if (nstereo) *nstereo = ns;
  return PLVI_OK;
}

int plvi_orb_wait_event(plvi_orb* h, void* cuda_event) {
  if (!h || !cuda_event) return PLVI_ERR_INVALID;
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  PLVI_CUDA_TRY(cudaStreamWaitEvent(h->stream, (cudaEvent_t)cuda_event, 0));
  return PLVI_OK;
}
// one thread polls a device counter (another pipeline's progress) with a time limit: stream-ordered work behind it starts
// when the counter has reached the target
__global__ void k_wait_counter(const int* counter, int target, unsigned long long maxNs) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (*(volatile const int*)counter < target) {
    __nanosleep(2000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > maxNs) break;
  }
}
int plvi_orb_wait_counter(plvi_orb* h, const int* d_counter, int target) {
  if (!h || !d_counter) return PLVI_ERR_INVALID;
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  k_wait_counter<<<1, 1, 0, h->stream>>>(d_counter, target, 30ull * 1000 * 1000);
  PLVI_CUDA_TRY(cudaGetLastError());
  return PLVI_OK;
}
int plvi_orb_wait_event_after_pyramid(plvi_orb* h, void* cuda_event) {
  if (!h) return PLVI_ERR_INVALID;
  h->waitAfterPyramid = (cudaEvent_t)cuda_event;
  return PLVI_OK;
}
int plvi_orb_graph_stats(const plvi_orb* h, int* captures) {
  if (!h) return PLVI_ERR_INVALID;
  if (captures) *captures = (int)h->graphs.captures;
  return (int)h->graphs.replays;
}

int plvi_orb_scale_factors(const plvi_orb* h, float* scale, float* inv_scale, float* sigma2,
                           float* inv_sigma2) {
  if (!h) return PLVI_ERR_INVALID;
  for (int i = 0; i < h->nlevels; i++) {
    if (scale) scale[i] = h->scale[i];
    if (inv_scale) inv_scale[i] = h->invScale[i];
    if (sigma2) sigma2[i] = h->sigma2[i];
    if (inv_sigma2) inv_sigma2[i] = h->invSigma2[i];
  }
  return PLVI_OK;
}

int plvi_orb_features_per_level(const plvi_orb* h, int* quota) {
  if (!h || !quota) return PLVI_ERR_INVALID;
  for (int i = 0; i < h->nlevels; i++) quota[i] = h->quota[i];
  return PLVI_OK;
}

int plvi_orb_level_sizes(const plvi_orb* h, int w, int hh, int* lw, int* lh) {
  if (!h || !lw || !lh) return PLVI_ERR_INVALID;
  for (int i = 0; i < h->nlevels; i++) {
    lw[i] = round_even_f((float)w * h->invScale[i]);
    lh[i] = round_even_f((float)hh * h->invScale[i]);
  }
  return PLVI_OK;
}

// The per-batch launch sequence, replayed from a captured CUDA graph when possible (GraphCache).
static int run_orb_pipeline(plvi_orb* h, const OrbPtrs& p, int n, int lap0, int lap1, plvi_keypoint* d_kps,
                            uint8_t* d_desc, int* d_counts, int* d_mono) {
  const cudaEvent_t waitEv = h->waitAfterPyramid;
  h->waitAfterPyramid = nullptr;
  // the pyramid of exactly these images was built ahead (plvi_orb_pyramid_device): skip that stage
  const uint64_t pk[6] = {(uint64_t)(uintptr_t)p.img[0], (uint64_t)n, (uint64_t)h->curW, (uint64_t)h->curH, (uint64_t)p.ipitch[0], (uint64_t)p.ifs[0]};
  const bool havePyr = h->pyrValid && std::equal(pk, pk + 6, h->pyrKey);
  h->pyrValid = false;
  const int stages = havePyr ? 2 : 3;
  auto record = [&](int* launches) {
    return launch_orb_pipeline(h->geom, p, h->scr, n, lap0, lap1, d_kps, d_desc, d_counts, d_mono, h->cap, h->stream,
                               launches, &h->prof, waitEv, stages);
  };
  if (h->prof.on || !h->graphs.on()) return record(&h->lastLaunches);
  std::vector<uint64_t> key = {(uint64_t)n, (uint64_t)h->curW, (uint64_t)h->curH, (uint64_t)(uintptr_t)p.img[0],
                               (uint64_t)p.ipitch[0], (uint64_t)p.ifs[0], (uint64_t)(uint32_t)lap0, (uint64_t)(uint32_t)lap1,
                               (uint64_t)(uintptr_t)d_kps, (uint64_t)(uintptr_t)d_desc, (uint64_t)(uintptr_t)d_counts,
                               (uint64_t)(uintptr_t)d_mono, (uint64_t)(uintptr_t)waitEv, (uint64_t)stages};
  return h->graphs.run(h->stream, key, &h->lastLaunches, record);
}

int plvi_orb_pyramid_device(plvi_orb* h, const uint8_t* d_imgs, int n, int w, int hh, int stride, size_t frame_stride) {
  int rc = check_batch(h, d_imgs, n, w, hh, stride);
  if (rc) return rc;
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  if ((rc = ensure_geom(h, w, hh))) return rc;
  OrbPtrs p;
  fill_ptrs(h, d_imgs, stride, frame_stride, p);
  int nl = 0;
  StageProf off;
  rc = launch_orb_pipeline(h->geom, p, h->scr, n, 0, 0, nullptr, nullptr, nullptr, nullptr, h->cap, h->stream, &nl, &off, nullptr, 1);
  if (rc) return rc;
  const uint64_t pk[6] = {(uint64_t)(uintptr_t)p.img[0], (uint64_t)n, (uint64_t)h->curW, (uint64_t)h->curH, (uint64_t)p.ipitch[0], (uint64_t)p.ifs[0]};
  std::copy(pk, pk + 6, h->pyrKey);
  h->pyrValid = true;
  return PLVI_OK;
}

int plvi_orb_extract_batch_device(plvi_orb* h, const uint8_t* d_imgs, int n, int w, int hh,
                                  int stride, size_t frame_stride, int lap0, int lap1,
                                  plvi_keypoint* d_kps, uint8_t* d_desc, int* d_counts,
                                  int* d_mono) {
  int rc = check_batch(h, d_imgs, n, w, hh, stride);
  if (rc) return rc;
  if (!d_kps || !d_desc || !d_counts || !d_mono) { set_error("null output"); return PLVI_ERR_INVALID; }
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  if ((rc = ensure_geom(h, w, hh))) return rc;
  OrbPtrs p;
  fill_ptrs(h, d_imgs, stride, frame_stride, p);
  h->lastPtrs = p;
  h->lastN = n;
  return run_orb_pipeline(h, p, n, lap0, lap1, d_kps, d_desc, d_counts, d_mono);
}

// result set of the current host-buffer call (second set allocated on first use) + its copy to the caller's buffers on
// the device-to-host stream
static int orb_result_set(plvi_orb* h, plvi_keypoint** k, uint8_t** d, int** c, int** m) {
  int rc = h->aout.begin(h->stream);
  if (rc) return rc;
  if (h->aout.sel == 1 && !h->dKpsB) {
    const size_t B = h->maxBatch, kt = h->cap;
    PLVI_CUDA_TRY(cudaMalloc(&h->dKpsB, B * kt * sizeof(plvi_keypoint)));
    PLVI_CUDA_TRY(cudaMalloc(&h->dDescB, B * kt * 32));
    PLVI_CUDA_TRY(cudaMalloc(&h->dCountsB, B * sizeof(int)));
    PLVI_CUDA_TRY(cudaMalloc(&h->dMonoB, B * sizeof(int)));
  }
  const bool b = h->aout.sel == 1;
  *k = b ? h->dKpsB : h->dKps; *d = b ? h->dDescB : h->dDesc; *c = b ? h->dCountsB : h->dCounts; *m = b ? h->dMonoB : h->dMono;
  return PLVI_OK;
}
static int orb_copy_back(plvi_orb* h, int n, const plvi_keypoint* dk, const uint8_t* dd, const int* dc, const int* dm, plvi_keypoint* kps,
                         uint8_t* desc, int* counts, int* mono_idx) {
  int rc = h->aout.start_copy(h->stream);
  if (rc) return rc;
  const size_t rows = (size_t)n * h->cap;
  cudaStream_t s = h->aout.d2h;
  PLVI_CUDA_TRY(cudaMemcpyAsync(counts, dc, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  PLVI_CUDA_TRY(cudaMemcpyAsync(mono_idx, dm, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  PLVI_CUDA_TRY(cudaMemcpyAsync(kps, dk, rows * sizeof(plvi_keypoint), cudaMemcpyDeviceToHost, s));
  PLVI_CUDA_TRY(cudaMemcpyAsync(desc, dd, rows * 32, cudaMemcpyDeviceToHost, s));
  return h->aout.end_copy();
}

int plvi_orb_extract_batch_async(plvi_orb* h, const uint8_t* imgs, int n, int w, int hh, int stride,
                                 size_t frame_stride, int lap0, int lap1, plvi_keypoint* kps,
                                 uint8_t* desc, int* counts, int* mono_idx) {
  int rc = check_batch(h, imgs, n, w, hh, stride);
  if (rc) return rc;
  if (!kps || !desc || !counts || !mono_idx) { set_error("null output"); return PLVI_ERR_INVALID; }
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  if ((rc = ensure_geom(h, w, hh))) return rc;
  const OrbLevel& L0 = h->geom.lv[0];
  u8* dIn = nullptr;
  if ((rc = h->ain.begin(h->dImg[0], (size_t)h->maxBatch * h->lvlCapBytes[0], &dIn))) return rc;
  int inPitch = 0;
  size_t inFs = 0;
  if ((rc = h->ain.upload(dIn, L0.pitch, L0.h, imgs, n, w, hh, stride, frame_stride, &inPitch, &inFs))) return rc;
  if ((rc = h->ain.uploaded(h->stream))) return rc;
  OrbPtrs p;
  fill_ptrs(h, dIn, inPitch, inFs, p);
  h->lastPtrs = p;
  h->lastN = n;
  plvi_keypoint* dk; uint8_t* dd; int* dc; int* dm;
  if ((rc = orb_result_set(h, &dk, &dd, &dc, &dm))) return rc;
  rc = run_orb_pipeline(h, p, n, lap0, lap1, dk, dd, dc, dm);
  if (rc) return rc;
  if ((rc = h->ain.finish(h->stream))) return rc;
  return orb_copy_back(h, n, dk, dd, dc, dm, kps, desc, counts, mono_idx);
}

int plvi_line_share_input(plvi_line* h, void* reader_stream, const uint8_t** d_img, int* pitch, size_t* frame_stride, int* n, int* w,
                          int* hh);
int plvi_line_share_done(plvi_line* h, void* reader_stream);

int plvi_orb_extract_batch_async_from_line(plvi_orb* h, plvi_line* src, int lap0, int lap1, plvi_keypoint* kps, uint8_t* desc,
                                           int* counts, int* mono_idx) {
  if (!h || !src || !kps || !desc || !counts || !mono_idx) { set_error("null argument"); return PLVI_ERR_INVALID; }
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  const uint8_t* dIn = nullptr;
  int pitch = 0, n = 0, w = 0, hh = 0;
  size_t fs = 0;
  int rc = plvi_line_share_input(src, h->stream, &dIn, &pitch, &fs, &n, &w, &hh);
  if (rc) return rc;
  if ((rc = check_batch(h, dIn, n, w, hh, pitch))) return rc;
  if ((rc = ensure_geom(h, w, hh))) return rc;
  OrbPtrs p;
  fill_ptrs(h, dIn, pitch, fs, p);
  h->lastPtrs = p;
  h->lastN = n;
  plvi_keypoint* dk; uint8_t* dd; int* dc; int* dm;
  if ((rc = orb_result_set(h, &dk, &dd, &dc, &dm))) return rc;
  rc = run_orb_pipeline(h, p, n, lap0, lap1, dk, dd, dc, dm);
  if (rc) return rc;
  if ((rc = plvi_line_share_done(src, h->stream))) return rc;
  return orb_copy_back(h, n, dk, dd, dc, dm, kps, desc, counts, mono_idx);
}

int plvi_orb_device_results(plvi_orb* h, plvi_keypoint** d_kps, uint8_t** d_desc, int** d_counts, int** d_mono_idx) {
  if (!h) return PLVI_ERR_INVALID;
  const bool b = h->aout.last == 1;
  if (d_kps) *d_kps = b ? h->dKpsB : h->dKps;
  if (d_desc) *d_desc = b ? h->dDescB : h->dDesc;
  if (d_counts) *d_counts = b ? h->dCountsB : h->dCounts;
  if (d_mono_idx) *d_mono_idx = b ? h->dMonoB : h->dMono;
  return PLVI_OK;
}

int plvi_orb_sync(plvi_orb* h) {
  if (!h) return PLVI_ERR_INVALID;
  PLVI_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return h->aout.sync();
}

void* plvi_orb_results_event(plvi_orb* h) { return h ? (void*)h->aout.last_done() : nullptr; }

int plvi_event_synchronize(void* cuda_event) {
  if (!cuda_event) return PLVI_OK;
  PLVI_CUDA_TRY(cudaEventSynchronize((cudaEvent_t)cuda_event));
  return PLVI_OK;
}
int plvi_stream_wait_event(void* stream, void* cuda_event) {
  if (!cuda_event) return PLVI_OK;
  PLVI_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)cuda_event, 0));
  return PLVI_OK;
}

int plvi_orb_set_profile(plvi_orb* h, int on) {
  if (!h) return PLVI_ERR_INVALID;
  h->prof.on = on != 0;
  return PLVI_OK;
}

const char* plvi_orb_profile(plvi_orb* h) {
  if (!h) return "";
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->profText = h->prof.report();
  return h->profText.c_str();
}

int plvi_orb_extract_batch(plvi_orb* h, const uint8_t* imgs, int n, int w, int hh, int stride,
                           size_t frame_stride, int lap0, int lap1, plvi_keypoint* kps,
                           uint8_t* desc, int* counts, int* mono_idx) {
  int rc = plvi_orb_extract_batch_async(h, imgs, n, w, hh, stride, frame_stride, lap0, lap1, kps,
                                        desc, counts, mono_idx);
  if (rc) return rc;
  return plvi_orb_sync(h);
}

int plvi_orb_read_level(plvi_orb* h, int frame, int level, int blurred, uint8_t* out) {
  if (!h || !out || level < 0 || level >= h->nlevels || frame < 0 || frame >= h->lastN || h->curW < 0)
    return PLVI_ERR_INVALID;
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  PLVI_CUDA_TRY(cudaStreamSynchronize(h->stream));
  const OrbLevel& L = h->geom.lv[level];
  const u8* src;
  int pitch;
  if (blurred) {
    src = h->lastPtrs.blur[level] + (size_t)frame * h->lastPtrs.bfs[level];
    pitch = L.pitch;
  } else {
    src = h->lastPtrs.img[level] + (size_t)frame * h->lastPtrs.ifs[level];
    pitch = h->lastPtrs.ipitch[level];
  }
  PLVI_CUDA_TRY(cudaMemcpy2D(out, L.w, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost));
  return PLVI_OK;
}

int plvi_orb_read_candidates(plvi_orb* h, int frame, int level, uint32_t* out, int cap, int* count) {
  if (!h || !out || !count || level < 0 || level >= h->nlevels || frame < 0 || frame >= h->lastN)
    return PLVI_ERR_INVALID;
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  PLVI_CUDA_TRY(cudaStreamSynchronize(h->stream));
  int n = 0;
  PLVI_CUDA_TRY(cudaMemcpy(&n, h->scr.candCount + (size_t)frame * h->nlevels + level, sizeof(int),
                           cudaMemcpyDeviceToHost));
  *count = n;
  const OrbLevel& L = h->geom.lv[level];
  const int m = std::min(std::min(n, cap), L.candCap);
  PLVI_CUDA_TRY(cudaMemcpy(out, h->scr.cand + (size_t)frame * h->geom.candTotal + L.candOff,
                           sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
  return PLVI_OK;
}

}  // extern "C"
