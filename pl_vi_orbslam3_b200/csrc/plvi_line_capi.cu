// C ABI for the line extractor (include/plvi.h): Lineextractor ctor / operator()
// (include/LineExtractor.h:55-61, src/LineExtractor.cc:39-117), geometry planning for
// LSDDetectorC::ComputePyramid + LineSegmentDetectorImpl::flsd + BinaryDescriptor.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "line_internal.cuh"
#include "lsd_gauss_table.h"

using namespace plvi;

struct plvi_line {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool ownStream = false;
  int nfeat = 0, nlevels = 0, refine = 0, brMaxWanted = -1;
  float lsdScale = 0.8f, scale = 2.f;
  int maxW = 0, maxH = 0, maxBatch = 0;
  int curW = -1, curH = -1;
  LineGeom geom, capGeom;
  LineBufs buf = {};
  u8* dImg[2] = {nullptr, nullptr};
  AsyncInput ain;   // input staging of the host-buffer entry points
  LineTab* dTabs = nullptr;
  int2* dRsTab = nullptr;
  double* dLbdG = nullptr;
  double* dLbdL = nullptr;
  double2* dTrig = nullptr;
  size_t tabCap = 0, rsCap = 0;
  plvi_keyline* dKl = nullptr;      // result set 0 (device copies of the host-buffer entry points' results)
  uint8_t* dDesc = nullptr;
  double* dEq = nullptr;
  int* dCounts = nullptr;
  plvi_keyline* dKlB = nullptr;     // result set 1, allocated by the second host-buffer call
  uint8_t* dDescB = nullptr;
  double* dEqB = nullptr;
  int* dCountsB = nullptr;
  AsyncOutput aout;
  int lastN = 0, lastLaunches = 0;
  StageProf prof;
  GraphCache graphs;
  std::string profText;
  bool debug = false;
  LineAux aux = {nullptr, nullptr, nullptr};
  LinePtrs lastPtrs = {};
};

namespace {

inline int round_even_f(float v) { return (int)nearbyintf(v); }
inline int round_even_d(double v) { return (int)nearbyint(v); }

void linear_rows_u8(int ssize, int dsize, std::vector<int2>& out) {
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

void linear_rows_f64(int ssize, int dsize, double inv_scale, std::vector<LineTab>& out) {
  const double scale = 1.0 / inv_scale;
  for (int d = 0; d < dsize; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    out.push_back(LineTab{s, 1.f - f, f});
  }
}

// first source index of scaled pixel d (the `ofs` of linear_rows_f64)
static int lsd_src_ofs(int d, int ssize, double inv_scale) {
  const double scale = 1.0 / inv_scale;
  const float f = (float)((d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  if (s < 0) s = 0;
  if (s >= ssize - 1) s = ssize - 1;
  return s;
}

// largest source window (columns or rows) of a k_lsd_pre tile of `tile` scaled pixels (+1 halo pixel)
static int lsd_pre_window(int ssize, int dsize, double inv_scale, int tile) {
  int m = 0;
  for (int d0 = 0; d0 < dsize; d0 += tile) {
    const int a = lsd_src_ofs(d0, ssize, inv_scale);
    const int b = std::min(lsd_src_ofs(std::min(d0 + tile, dsize - 1), ssize, inv_scale) + 1, ssize - 1);
    m = std::max(m, b - a + 1);
  }
  return m;
}

// cv::getGaussianKernel(n, sigma, CV_64F) for LSD's sigma = 0.6 / (double)lsd_scale.  OpenCV 4.x builds the kernel with
// its own soft-float exp, whose last bits differ from libm's: for the common lsd_scale settings (0.3 ... 0.95 in steps
// of 0.05, incl. the shipped 0.8 and the 0.5 / 0.6 the reference's yaml comments recommend) the values probed from
// OpenCV are used verbatim (lsd_gauss_table.h, tools/gen_lsd_gauss.py); any other scale gets the defining formula
// evaluated with libm (taps within a few ulp of OpenCV's: documented in DESIGN.md).
void lsd_gaussian_kernel(float lsdScale, int n, double sigma, double* k) {
  uint32_t bits;
  memcpy(&bits, &lsdScale, 4);
  for (const LsdGaussEntry& e : kLsdGaussTable)
    if (e.scale_bits == bits && e.n == n) { memcpy(k, e.k, sizeof(double) * n); return; }
  const double s2 = -0.5 / (sigma * sigma);
  double sum = 0;
  for (int i = 0; i < n; i++) { const double x = i - (n - 1) * 0.5; k[i] = exp(s2 * x * x); sum += k[i]; }
  sum = 1. / sum;
  for (int i = 0; i < n; i++) k[i] *= sum;
}

int make_geom(const plvi_line* h, int w, int hh, LineGeom& g, std::vector<LineTab>* tabs, std::vector<int2>* rs) {
  memset(&g, 0, sizeof(g));
  g.noct = h->nlevels;
  g.nfeat = h->nfeat;
  g.refine = h->refine;
  g.keepCap = h->nfeat > 0 ? h->nfeat : 4096;
  g.lineScale = h->scale;
  g.lsdScale = (double)h->lsdScale;
  g.prec = M_PI * 22.5 / 180;
  g.rho = 2.0 / sin(g.prec);
  // fastAtan2 deviates from atan2 by <= 0.0092 deg (1.6e-4 rad); 5e-4 rad leaves a 3x margin (the exact test runs only inside that band)
  g.alignHi2 = (float)(cos(g.prec - 5e-4) * cos(g.prec - 5e-4));
  g.alignLo2 = (float)(cos(g.prec + 5e-4) * cos(g.prec + 5e-4));
  g.minLength = 0.025 * std::min(w, hh);
  // flsd (src/LSD/lsd.cpp:446-462): SCALE == 1 works on the image itself; else GaussianBlur(sigma = 0.6 / SCALE, ksize
  // 1 + 2 h) and resize by SCALE.  The range is (0, 1] by the reference's own yaml comment; h <= 8 <=> scale >= 0.279.
  g.hk = 0;
  if (g.lsdScale != 1.0) {
    const double sigma = 0.6 / g.lsdScale;
    const unsigned hk = (unsigned)ceil(sigma * sqrt(2 * 3.0 * log(10.0)));
    if (!(g.lsdScale > 0) || g.lsdScale > 1.0 || hk < 3 || hk > 8) {
      set_error("lsd_scale outside the supported range [0.28, 1]");
      return PLVI_ERR_INVALID;
    }
    g.hk = (int)hk;
    lsd_gaussian_kernel(h->lsdScale, 1 + 2 * g.hk, sigma, g.kern);
  }
  // k_lsd_pre tiles: 32 x preTH scaled pixels; the taller the tile the less of the Gaussian's row halo is recomputed,
  // the smaller lsd_scale the larger its source window in shared memory
  g.preTH = g.lsdScale >= 0.7 ? 32 : (g.lsdScale >= 0.45 ? 16 : 8);
  float sf = 1.f;
  size_t px = 0, raw = 0, lbd = 0, reg = 0, sbm = 0, srec = 0, brBm = 0, brRec = 0, brList = 0;
  int bm = 0, seg = 0, tabOff = 0, task = 0, brBand = 0;
  for (int o = 0; o < g.noct; o++) {
    LineOct& O = g.o[o];
    if (o > 0) sf = sf * h->scale;
    const float isf = 1.0f / sf;
    O.w = round_even_f((float)w * isf);
    O.h = round_even_f((float)hh * isf);
    O.pitch = (O.w + 63) & ~63;
    O.sw = round_even_d(O.w * g.lsdScale);
    O.sh = round_even_d(O.h * g.lsdScale);
    if (O.sw < 8 || O.sh < 8 || O.sw > 65535 || O.sh > 65535) { set_error("image size unsupported for LSD"); return PLVI_ERR_INVALID; }
    O.wpr = O.sw / 32 + 1;   // at least one padding bit per row: x = -1 and x = W read as "not available"
    if (g.hk > 0 && (O.w < 2 * g.hk + 2 || O.h < 2 * g.hk + 2)) { set_error("image smaller than the LSD Gaussian"); return PLVI_ERR_INVALID; }
    if (g.lsdScale == 0.5 && (2 * O.sw > O.w || 2 * O.sh > O.h)) {
      // cv::resize runs its 2x2 area path here; the border handling of blocks that leave the image is not modelled
      set_error("lsd_scale 0.5 needs octave sizes whose halves round down (w, h not 3 mod 4)");
      return PLVI_ERR_INVALID;
    }
    O.preSW = O.preRH = 0;
    if (g.hk > 0) {
      O.preSW = (lsd_pre_window(O.w, O.sw, g.lsdScale, 32) + 3) & ~3;
      O.preRH = lsd_pre_window(O.h, O.sh, g.lsdScale, g.preTH);
    }
    const double LOG_NT = 5 * (log10((double)O.sw) + log10((double)O.sh)) / 2 + log10(11.0);
    O.minRegSize = (int)(-LOG_NT / log10(22.5 / 180));
    O.pxOff = px; px += ((size_t)O.sw * O.sh + 3) & ~(size_t)3;
    O.rawOff = raw; raw += (size_t)O.w * O.h;
    O.bmOff = bm; bm += O.wpr * O.sh;
    O.segOff = seg;
    O.segCap = std::min((O.sw * O.sh) / std::max(O.minRegSize, 1), std::max(1024, 8192 >> (2 * o)));
    seg += O.segCap;
    {  // speculation bands: equal pixel counts per band across octaves (16 bands on octave 0, 4 on octave 1)
      static const int nb0 = [] { const char* ev = getenv("PLVI_LSD_BANDS"); return ev ? std::max(1, atoi(ev)) : 16; }();
      const int nbT = std::max(1, nb0 >> (2 * o));
      O.bandRows = (O.sh + nbT - 1) / nbT;
      O.nbands = (O.sh + O.bandRows - 1) / O.bandRows;
      O.bandPxCap = 2 * O.bandRows * O.sw;
      O.bandRecCap = std::min(std::max(O.bandRows * O.sw / 2, 64), 4096);
      O.regOff = reg; reg += ((size_t)O.sw * O.sh + (size_t)nbT * O.bandPxCap + 3) & ~(size_t)3;
      O.specBmOff = sbm; sbm += (size_t)nbT * O.wpr * O.sh;
      O.specRecOff = srec; srec += (size_t)nbT * O.bandRecCap;
      O.taskOff = task; task += nbT;
    }
    {  // band-run (small batches): bands of equal pixel counts across octaves, one warp each
      static const int rows0 = [] { const char* ev = getenv("PLVI_LSD_BR_ROWS"); return ev ? std::max(2, atoi(ev)) : 6; }();
      O.brRows = std::min(rows0 << o, std::max(O.sh, 2));
      O.brBands = (O.sh + O.brRows - 1) / O.brRows;
      O.brPxCap = 4 * O.brRows * O.sw + 64;
      O.brRecCap = std::min(std::max(O.brRows * O.sw / 2, 64), 4096);
      O.brBandOff = brBand; brBand += O.brBands;
      O.brBmOff = brBm; brBm += (size_t)O.brBands * O.wpr * O.sh;
      O.brRecOff = brRec; brRec += (size_t)O.brBands * O.brRecCap;
      O.brListOff = brList; brList += (size_t)O.brBands * O.brPxCap;
    }
    O.xtabOff = tabOff; tabOff += O.sw;
    O.ytabOff = tabOff; tabOff += O.sh;
    if (tabs) {
      linear_rows_f64(O.w, O.sw, g.lsdScale, *tabs);
      linear_rows_f64(O.h, O.sh, g.lsdScale, *tabs);
    }
    O.lw = o == 0 ? w : g.o[o - 1].lw / 2;
    O.lh = o == 0 ? hh : g.o[o - 1].lh / 2;
    O.lpitch = (O.lw + 63) & ~63;
    O.lbdOff = lbd; lbd += (size_t)O.lw * O.lh;
    // (for odd sizes the LBD octave (w/2, h/2) can be one pixel smaller than the LSD octave
    //  cvRound(w/2): computeLBD clamps its sample coordinates, exactly like the reference)
  }
  if (rs && g.noct > 1) {
    linear_rows_u8(g.o[0].w, g.o[1].w, *rs);
    linear_rows_u8(g.o[0].h, g.o[1].h, *rs);
  }
  g.pxTotal = px; g.rawTotal = raw; g.lbdTotal = lbd; g.bmTotal = bm; g.segTotal = seg;
  g.regTotal = reg; g.specBmTotal = sbm; g.specRecTotal = srec; g.tasksPerFrame = task;
  g.brBandsPerFrame = brBand; g.brBmTotal = brBm; g.brRecTotal = brRec; g.brListTotal = brList;
  return PLVI_OK;
}

int ensure_geom(plvi_line* h, int w, int hh) {
  if (w == h->curW && hh == h->curH) return PLVI_OK;
  if (w > h->maxW || hh > h->maxH) { set_error("image larger than the handle's max_width/max_height"); return PLVI_ERR_CAPACITY; }
  std::vector<LineTab> tabs;
  std::vector<int2> rs;
  LineGeom g;
  int rc = make_geom(h, w, hh, g, &tabs, &rs);
  if (rc) return rc;
  const LineGeom& c = h->capGeom;
  if (tabs.size() > h->tabCap || rs.size() > h->rsCap || g.pxTotal > c.pxTotal || g.rawTotal > c.rawTotal ||
      g.lbdTotal > c.lbdTotal || g.bmTotal > c.bmTotal || g.segTotal > c.segTotal || g.regTotal > c.regTotal ||
      g.specBmTotal > c.specBmTotal || g.specRecTotal > c.specRecTotal || g.tasksPerFrame != c.tasksPerFrame ||
      g.brBmTotal > c.brBmTotal || g.brRecTotal > c.brRecTotal || g.brListTotal > c.brListTotal ||
      g.brBandsPerFrame > c.brBandsPerFrame) {
    set_error("internal: line geometry exceeds allocated capacity");
    return PLVI_ERR_CAPACITY;
  }
  PLVI_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->graphs.clear();   // captured graphs hold the old geometry
  PLVI_CUDA_TRY(cudaMemcpy(h->dTabs, tabs.data(), tabs.size() * sizeof(LineTab), cudaMemcpyHostToDevice));
  if (!rs.empty()) PLVI_CUDA_TRY(cudaMemcpy(h->dRsTab, rs.data(), rs.size() * sizeof(int2), cudaMemcpyHostToDevice));
  // keep the allocated per-frame strides
  g.pxTotal = c.pxTotal; g.rawTotal = c.rawTotal; g.lbdTotal = c.lbdTotal; g.bmTotal = c.bmTotal; g.segTotal = c.segTotal;
  g.regTotal = c.regTotal; g.specBmTotal = c.specBmTotal; g.specRecTotal = c.specRecTotal;
  // band-run: strides of the allocation; the band offsets inside a frame are those of the current geometry (they
  // fit: every per-frame total is bounded by the capacity geometry's)
  g.brBmTotal = c.brBmTotal; g.brRecTotal = c.brRecTotal; g.brListTotal = c.brListTotal;
  const int curBands = g.brBandsPerFrame;
  g.brBandsPerFrame = c.brBandsPerFrame;
  (void)curBands;
  for (int o = 0; o < g.noct; o++) {
    g.o[o].regOff = c.o[o].regOff; g.o[o].specBmOff = c.o[o].specBmOff; g.o[o].specRecOff = c.o[o].specRecOff;
    g.o[o].pxOff = c.o[o].pxOff; g.o[o].rawOff = c.o[o].rawOff; g.o[o].bmOff = c.o[o].bmOff;
    g.o[o].segOff = c.o[o].segOff; g.o[o].segCap = std::min(g.o[o].segCap, c.o[o].segCap);
    g.o[o].lbdOff = c.o[o].lbdOff;
  }
  h->geom = g;
  rc = line_kernel_attrs(h->geom);
  if (rc) return rc;
  h->curW = w;
  h->curH = hh;
  return PLVI_OK;
}

void fill_ptrs(plvi_line* h, const u8* l0, int l0pitch, size_t l0fs, LinePtrs& p) {
  for (int o = 0; o < 2; o++) {
    p.img[o] = h->dImg[o];
    p.ipitch[o] = h->geom.o[o].pitch;
    p.ifs[o] = (size_t)h->geom.o[o].pitch * h->geom.o[o].h;
  }
  if (l0) { p.img[0] = l0; p.ipitch[0] = l0pitch; p.ifs[0] = l0fs; }
}

int check_batch(plvi_line* h, const void* imgs, int n, int w, int hh, int stride) {
  if (!h) { set_error("null handle"); return PLVI_ERR_INVALID; }
  if (!imgs || w <= 0 || hh <= 0 || n <= 0) { set_error("empty image batch"); return PLVI_ERR_EMPTY; }
  if (stride < w) { set_error("stride < width"); return PLVI_ERR_INVALID; }
  if (n > h->maxBatch) { set_error("batch larger than max_batch"); return PLVI_ERR_CAPACITY; }
  return PLVI_OK;
}

}  // namespace

extern "C" {

int plvi_line_create(plvi_line** out, int lsd_nfeatures, int lsd_refine, float lsd_scale, int nlevels, float scale,
                     int extractor, int max_width, int max_height, int max_batch, int device, void* stream) {
  return plvi_line_create_ex(out, lsd_nfeatures, lsd_refine, lsd_scale, nlevels, scale, extractor, max_width, max_height, max_batch,
                             device, stream, -1);
}

int plvi_line_create_ex(plvi_line** out, int lsd_nfeatures, int lsd_refine, float lsd_scale, int nlevels, float scale,
                        int extractor, int max_width, int max_height, int max_batch, int device, void* stream, int band_run_max) {
  if (!out || lsd_nfeatures < 0 || max_batch < 1 || max_width < 1 || max_height < 1 || !(lsd_scale > 0.f)) {
    set_error("plvi_line_create: invalid argument");
    return PLVI_ERR_INVALID;
  }
  if (extractor != 0) { set_error("extractor=1 (EDLines) is outside the hot path: only the LSD branch is implemented"); return PLVI_ERR_INVALID; }
  if (lsd_refine < 0 || lsd_refine > 2) { set_error("lsd_refine must be 0 (none), 1 (standard) or 2 (advanced)"); return PLVI_ERR_INVALID; }
  if (nlevels < 1 || nlevels > 2) { set_error("levels must be 1 or 2 (as in the reference's yaml contract)"); return PLVI_ERR_INVALID; }
  PLVI_CUDA_TRY(cudaSetDevice(device));
  plvi_line* h = new plvi_line();
  h->device = device;
  h->nfeat = lsd_nfeatures;
  h->lsdScale = lsd_scale;
  h->refine = lsd_refine;
  h->brMaxWanted = band_run_max;
  h->nlevels = nlevels;
  h->scale = scale;
  h->maxW = max_width;
  h->maxH = max_height;
  h->maxBatch = max_batch;
  std::vector<LineTab> tabs;
  std::vector<int2> rs;
  int rc = make_geom(h, max_width, max_height, h->capGeom, &tabs, &rs);
  if (rc) { delete h; return rc; }
  if (stream) h->stream = (cudaStream_t)stream;
  else {
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); delete h; return PLVI_ERR_CUDA; }
    h->ownStream = true;
  }
  if (cudaStreamCreateWithFlags(&h->aux.stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->aux.fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->aux.join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->aux.stage, cudaEventDisableTiming) != cudaSuccess) {
    set_error("cannot create the auxiliary stream");
    plvi_line_destroy(h);
    return PLVI_ERR_CUDA;
  }
  const LineGeom& c = h->capGeom;
  const size_t B = max_batch;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes + 256); };
  for (int o = 0; o < nlevels; o++) A((void**)&h->dImg[o], B * c.o[o].pitch * c.o[o].h);
  A((void**)&h->buf.ang, B * c.pxTotal * sizeof(float));
  A((void**)&h->buf.cs, B * c.pxTotal * sizeof(float2));
  A((void**)&h->buf.seed, B * c.pxTotal * sizeof(float2));
  A((void**)&h->buf.mod, B * c.pxTotal * sizeof(double));
  A((void**)&h->buf.bitmap, B * c.bmTotal * sizeof(unsigned));
  A((void**)&h->buf.reg, B * c.regTotal * sizeof(unsigned));
  A((void**)&h->buf.specBm, B * c.specBmTotal * sizeof(unsigned));
  A((void**)&h->buf.specRec, B * c.specRecTotal * sizeof(SpecRec));
  A((void**)&h->buf.specCnt, B * c.tasksPerFrame * sizeof(int));
  A((void**)&h->buf.specStart, 64);
  A((void**)&h->buf.bandRow, B * (c.tasksPerFrame + 2) * sizeof(int));
  A((void**)&h->buf.phantom, B * c.bmTotal * sizeof(unsigned));
  {  // band-run buffers for batches of up to PLVI_LSD_BR_MAX frames (0 switches the path off).  Default 384: measured on
    // B200 the band-run rounds beat band speculation + serial commit up to ~512 frames per batch (32 frames: 5.1x, 128:
    // 2.8x, 256: 1.7x, 512: 1.0x, profiles/r02_notes.md); about 10 MB of scratch per frame of that capacity.
    const char* ev = getenv("PLVI_LSD_BR_MAX");
    const int brMax = std::min(h->brMaxWanted >= 0 ? h->brMaxWanted : (ev ? std::max(0, atoi(ev)) : 384), max_batch);
    const char* er = getenv("PLVI_LSD_BR_ROUNDS");
    // 28 rounds: of 3 x 512 (frame, octave) problems measured (640x480, 752x480, 1280x720) the slowest needed 16; rounds
    // after the fixed point only cost their launches, a problem that has not converged costs a full serial chain
    h->buf.brRounds = std::min(std::max(er ? atoi(er) : 28, 1), BR_FLAGS - 4);
    h->buf.brMax = brMax;
    h->buf.brUse = brMax;
    if (brMax > 0) {
      const size_t S = brMax;
      A((void**)&h->buf.brIn, S * c.brBmTotal * sizeof(unsigned));
      A((void**)&h->buf.brWk, S * c.brBmTotal * sizeof(unsigned));
      A((void**)&h->buf.brPh, S * c.brBmTotal * sizeof(unsigned));
      A((void**)&h->buf.brRec, S * 2 * c.brRecTotal * sizeof(uint4));
      A((void**)&h->buf.brList, S * 2 * c.brListTotal * sizeof(unsigned));
      A((void**)&h->buf.brState, S * c.brBandsPerFrame * 8 * sizeof(int));
    }
    A((void**)&h->buf.brFlags, (size_t)std::max(brMax, 1) * 2 * BR_FLAGS * sizeof(int));
  }
  A((void**)&h->buf.regTab, B * c.segTotal * sizeof(LineRegion));
  A((void**)&h->buf.regCount, B * 2 * sizeof(int));
  A((void**)&h->buf.segs, B * c.segTotal * sizeof(float4));
  A((void**)&h->buf.tmpResp, B * c.segTotal * sizeof(float));
  A((void**)&h->buf.tmpCls, B * c.segTotal * sizeof(int));
  A((void**)&h->buf.lbdImg0, B * c.o[0].lpitch * c.o[0].lh);
  if (nlevels > 1) A((void**)&h->buf.lbdImg1, B * c.o[1].lpitch * c.o[1].lh);
  A((void**)&h->buf.grad, B * c.lbdTotal * sizeof(short2));
  A((void**)&h->buf.lbdRows, B * c.keepCap * 512 * sizeof(float));
  h->tabCap = tabs.size() + 64;
  h->rsCap = rs.size() + 64;
  A((void**)&h->dTabs, h->tabCap * sizeof(LineTab));
  A((void**)&h->dRsTab, h->rsCap * sizeof(int2));
  A((void**)&h->dLbdG, 63 * sizeof(double));
  A((void**)&h->dLbdL, 21 * sizeof(double));
  A((void**)&h->dTrig, 1024 * sizeof(double2));
  A((void**)&h->dKl, B * c.keepCap * sizeof(plvi_keyline));
  A((void**)&h->dDesc, B * c.keepCap * 32);
  A((void**)&h->dEq, B * c.keepCap * 3 * sizeof(double));
  A((void**)&h->dCounts, B * sizeof(int));
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
    plvi_line_destroy(h);
    return PLVI_ERR_CUDA;
  }
  {  // BinaryDescriptor ctor weights (binary_descriptor_custom.cpp:219-261; integer divisions kept)
    double L[21], G[63];
    double u = (7 * 3 - 1) / 2, sigma = (7 * 2 + 1) / 2, inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < 21; i++) { const double d = i - u; L[i] = exp(d * d * inv); }
    u = (9 * 7 - 1) / 2; sigma = u; inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < 63; i++) { const double d = i - u; G[i] = exp(d * d * inv); }
    cudaMemcpy(h->dLbdG, G, sizeof(G), cudaMemcpyHostToDevice);
    cudaMemcpy(h->dLbdL, L, sizeof(L), cudaMemcpyHostToDevice);
    std::vector<double2> trig(1024);
    for (int k = 0; k < 1024; k++) trig[k] = make_double2(cos(k * (2.0 * M_PI / 1024)), sin(k * (2.0 * M_PI / 1024)));
    cudaMemcpy(h->dTrig, trig.data(), trig.size() * sizeof(double2), cudaMemcpyHostToDevice);
  }
  h->buf.tabs = h->dTabs;
  h->buf.rsTab = h->dRsTab;
  h->buf.lbdG = h->dLbdG;
  h->buf.lbdL = h->dLbdL;
  h->buf.trig = h->dTrig;
  {  // PLVI_LSD_SPEC=0 selects the serial region growing (A/B measurements)
    const char* ev = getenv("PLVI_LSD_SPEC");
    h->buf.useSpec = !(ev && ev[0] == '0');
    const char* eq = getenv("PLVI_LSD_EQLOAD");
    h->buf.eqLoad = eq ? (eq[0] == '0' ? 0 : 2) : 1;   // 1 (default): by batch size, 2 / 0: forced on / off
  }
  *out = h;
  return PLVI_OK;
}

void plvi_line_destroy(plvi_line* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->ain.destroy();
  if (h->aux.stream) { cudaStreamSynchronize(h->aux.stream); cudaStreamDestroy(h->aux.stream); }
  if (h->aux.fork) cudaEventDestroy(h->aux.fork);
  if (h->aux.join) cudaEventDestroy(h->aux.join);
  if (h->aux.stage) cudaEventDestroy(h->aux.stage);
  cudaFree(h->dImg[0]); cudaFree(h->dImg[1]);
  cudaFree(h->buf.ang); cudaFree(h->buf.cs); cudaFree(h->buf.seed); cudaFree(h->buf.mod); cudaFree(h->buf.bitmap);
  cudaFree(h->buf.specBm); cudaFree(h->buf.specRec); cudaFree(h->buf.specCnt); cudaFree(h->buf.bandRow); cudaFree(h->buf.specStart); cudaFree(h->buf.phantom);
  cudaFree(h->buf.brIn); cudaFree(h->buf.brWk); cudaFree(h->buf.brPh); cudaFree(h->buf.brRec); cudaFree(h->buf.brList);
  cudaFree(h->buf.brState); cudaFree(h->buf.brFlags);
  cudaFree(h->buf.reg); cudaFree(h->buf.regTab); cudaFree(h->buf.regCount); cudaFree(h->buf.segs);
  cudaFree(h->buf.tmpResp); cudaFree(h->buf.tmpCls); cudaFree(h->buf.lbdImg0); cudaFree(h->buf.lbdImg1);
  cudaFree(h->buf.grad); cudaFree(h->buf.lbdRows); cudaFree(h->buf.scaledDbg);
  cudaFree(h->dTabs); cudaFree(h->dRsTab); cudaFree(h->dLbdG); cudaFree(h->dLbdL); cudaFree(h->dTrig);
  h->aout.destroy();
  cudaFree(h->dKl); cudaFree(h->dDesc); cudaFree(h->dEq); cudaFree(h->dCounts);
  cudaFree(h->dKlB); cudaFree(h->dDescB); cudaFree(h->dEqB); cudaFree(h->dCountsB);
  if (h->ownStream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int plvi_line_capacity(const plvi_line* h) { return h ? h->capGeom.keepCap : PLVI_ERR_INVALID; }
void* plvi_line_stream(const plvi_line* h) { return h ? (void*)h->stream : nullptr; }
int plvi_line_last_launches(const plvi_line* h) { return h ? h->lastLaunches : PLVI_ERR_INVALID; }
void* plvi_line_stage_event(plvi_line* h) { return h ? (void*)h->aux.stage : nullptr; }
const int* plvi_line_stage_counter(plvi_line* h, int* target) {
  if (!h) return nullptr;
  if (target) {
    // blocks of k_lsd_spec in the last batch, capped at what is resident at once beside a few other blocks (20 of the
    // 24 warps of 80 registers per SM)
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int blocks = h->geom.tasksPerFrame * ((h->lastN + 32 * SPEC_WPB - 1) / (32 * SPEC_WPB));
    *target = std::min(blocks, sms * (20 / SPEC_WPB));
  }
  return h->buf.specStart;
}
int plvi_line_graph_stats(const plvi_line* h, int* captures) {
  if (!h) return PLVI_ERR_INVALID;
  if (captures) *captures = (int)h->graphs.captures;
  return (int)h->graphs.replays;
}
int plvi_line_levels(const plvi_line* h) { return h ? h->nlevels : PLVI_ERR_INVALID; }

int plvi_line_scale_factors(const plvi_line* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
  if (!h) return PLVI_ERR_INVALID;
  float sf = 1.f;
  for (int i = 0; i < h->nlevels; i++) {
    if (i > 0) sf = sf * h->scale;
    if (scale) scale[i] = sf;
    if (inv_scale) inv_scale[i] = 1.0f / sf;
    const float s2 = i > 0 ? sf * sf : 1.0f;
    if (sigma2) sigma2[i] = s2;
    if (inv_sigma2) inv_sigma2[i] = 1.0f / s2;
  }
  return PLVI_OK;
}

int plvi_line_octave_sizes(const plvi_line* h, int w, int hh, int* ow, int* oh, int* sw, int* sh) {
  if (!h) return PLVI_ERR_INVALID;
  LineGeom g;
  int rc = make_geom(h, w, hh, g, nullptr, nullptr);
  if (rc) return rc;
  for (int o = 0; o < g.noct; o++) {
    if (ow) ow[o] = g.o[o].w;
    if (oh) oh[o] = g.o[o].h;
    if (sw) sw[o] = g.o[o].sw;
    if (sh) sh[o] = g.o[o].sh;
  }
  return PLVI_OK;
}

int plvi_line_set_debug(plvi_line* h, int on) {
  if (!h) return PLVI_ERR_INVALID;
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  PLVI_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->graphs.clear();   // the debug buffer pointer is baked into captured graphs
  if (on && !h->buf.scaledDbg)
    PLVI_CUDA_TRY(cudaMalloc((void**)&h->buf.scaledDbg, (size_t)h->maxBatch * h->capGeom.pxTotal * sizeof(double)));
  if (!on && h->buf.scaledDbg) { cudaFree(h->buf.scaledDbg); h->buf.scaledDbg = nullptr; }
  h->debug = on != 0;
  return PLVI_OK;
}

// The per-batch launch sequence, replayed from a captured CUDA graph when possible (GraphCache).
static int run_line_pipeline(plvi_line* h, const LinePtrs& p, int n, plvi_keyline* d_kl, uint8_t* d_desc, double* d_eq,
                             int* d_counts) {
  auto record = [&](int* launches) {
    return launch_line_pipeline(h->geom, p, h->buf, n, d_kl, d_desc, d_eq, d_counts, h->stream, h->aux, launches, &h->prof);
  };
  if (h->prof.on || !h->graphs.on()) return record(&h->lastLaunches);
  std::vector<uint64_t> key = {(uint64_t)n, (uint64_t)h->curW, (uint64_t)h->curH, (uint64_t)(uintptr_t)p.img[0],
                               (uint64_t)p.ipitch[0], (uint64_t)p.ifs[0], (uint64_t)(uintptr_t)d_kl,
                               (uint64_t)(uintptr_t)d_desc, (uint64_t)(uintptr_t)d_eq, (uint64_t)(uintptr_t)d_counts,
                               (uint64_t)h->buf.brUse};
  return h->graphs.run(h->stream, key, &h->lastLaunches, record);
}

int plvi_line_extract_batch_device(plvi_line* h, const uint8_t* d_imgs, int n, int w, int hh, int stride,
                                   size_t frame_stride, plvi_keyline* d_kl, uint8_t* d_desc, double* d_eq,
                                   int* d_counts) {
  int rc = check_batch(h, d_imgs, n, w, hh, stride);
  if (rc) return rc;
  if (!d_kl || !d_desc || !d_eq || !d_counts) { set_error("null output"); return PLVI_ERR_INVALID; }
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  if ((rc = ensure_geom(h, w, hh))) return rc;
  LinePtrs p;
  fill_ptrs(h, d_imgs, stride, frame_stride, p);
  h->lastPtrs = p;
  h->lastN = n;
  return run_line_pipeline(h, p, n, d_kl, d_desc, d_eq, d_counts);
}

int plvi_line_extract_batch_async(plvi_line* h, const uint8_t* imgs, int n, int w, int hh, int stride,
                                  size_t frame_stride, plvi_keyline* kl, uint8_t* desc, double* line_eq,
                                  int* counts) {
  int rc = check_batch(h, imgs, n, w, hh, stride);
  if (rc) return rc;
  if (!kl || !desc || !line_eq || !counts) { set_error("null output"); return PLVI_ERR_INVALID; }
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  if ((rc = ensure_geom(h, w, hh))) return rc;
  const LineOct& O0 = h->geom.o[0];
  u8* dIn = nullptr;
  if ((rc = h->ain.begin(h->dImg[0], (size_t)h->maxBatch * h->capGeom.o[0].pitch * h->capGeom.o[0].h, &dIn))) return rc;
  int inPitch = 0;
  size_t inFs = 0;
  if ((rc = h->ain.upload(dIn, O0.pitch, O0.h, imgs, n, w, hh, stride, frame_stride, &inPitch, &inFs))) return rc;
  if ((rc = h->ain.uploaded(h->stream))) return rc;
  LinePtrs p;
  fill_ptrs(h, dIn, inPitch, inFs, p);
  h->lastPtrs = p;
  h->lastN = n;
  // two device result sets: the copy of this call's results (device-to-host stream) overlaps the next call's kernels
  if ((rc = h->aout.begin(h->stream))) return rc;
  if (h->aout.sel == 1 && !h->dKlB) {
    const size_t B = h->maxBatch, kc = h->capGeom.keepCap;
    PLVI_CUDA_TRY(cudaMalloc(&h->dKlB, B * kc * sizeof(plvi_keyline)));
    PLVI_CUDA_TRY(cudaMalloc(&h->dDescB, B * kc * 32));
    PLVI_CUDA_TRY(cudaMalloc(&h->dEqB, B * kc * 3 * sizeof(double)));
    PLVI_CUDA_TRY(cudaMalloc(&h->dCountsB, B * sizeof(int)));
  }
  const bool sb = h->aout.sel == 1;
  plvi_keyline* dk = sb ? h->dKlB : h->dKl;
  uint8_t* dd = sb ? h->dDescB : h->dDesc;
  double* de = sb ? h->dEqB : h->dEq;
  int* dc = sb ? h->dCountsB : h->dCounts;
  rc = run_line_pipeline(h, p, n, dk, dd, de, dc);
  if (rc) return rc;
  if ((rc = h->ain.finish(h->stream))) return rc;
  if ((rc = h->aout.start_copy(h->stream))) return rc;
  const size_t rows = (size_t)n * h->geom.keepCap;
  cudaStream_t s = h->aout.d2h;
  PLVI_CUDA_TRY(cudaMemcpyAsync(counts, dc, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  PLVI_CUDA_TRY(cudaMemcpyAsync(kl, dk, rows * sizeof(plvi_keyline), cudaMemcpyDeviceToHost, s));
  PLVI_CUDA_TRY(cudaMemcpyAsync(desc, dd, rows * 32, cudaMemcpyDeviceToHost, s));
  PLVI_CUDA_TRY(cudaMemcpyAsync(line_eq, de, rows * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  return h->aout.end_copy();
}

// the frames of the last plvi_line_extract_batch_async call, for a second reader (plvi_orb_extract_batch_async_from_line)
int plvi_line_share_input(plvi_line* h, void* reader_stream, const uint8_t** d_img, int* pitch, size_t* frame_stride, int* n, int* w,
                          int* hh) {
  if (!h || !d_img || !pitch || !frame_stride) return PLVI_ERR_INVALID;
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  if (h->ain.share_last((cudaStream_t)reader_stream, d_img, pitch, frame_stride)) { set_error("no host-buffer call has been made on the line handle"); return PLVI_ERR_INVALID; }
  if (n) *n = h->lastN;
  if (w) *w = h->curW;
  if (hh) *hh = h->curH;
  return PLVI_OK;
}
int plvi_line_share_done(plvi_line* h, void* reader_stream) {
  if (!h) return PLVI_ERR_INVALID;
  return h->ain.share_done((cudaStream_t)reader_stream);
}

int plvi_line_device_results(plvi_line* h, plvi_keyline** d_keylines, uint8_t** d_desc, double** d_line_eq, int** d_counts) {
  if (!h) return PLVI_ERR_INVALID;
  const bool sb = h->aout.last == 1;
  if (d_keylines) *d_keylines = sb ? h->dKlB : h->dKl;
  if (d_desc) *d_desc = sb ? h->dDescB : h->dDesc;
  if (d_line_eq) *d_line_eq = sb ? h->dEqB : h->dEq;
  if (d_counts) *d_counts = sb ? h->dCountsB : h->dCounts;
  return PLVI_OK;
}

int plvi_line_sync(plvi_line* h) {
  if (!h) return PLVI_ERR_INVALID;
  PLVI_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return h->aout.sync();
}

void* plvi_line_results_event(plvi_line* h) { return h ? (void*)h->aout.last_done() : nullptr; }

int plvi_line_set_band_run_max(plvi_line* h, int max_frames) {
  if (!h || max_frames < 0) return PLVI_ERR_INVALID;
  h->buf.brUse = std::min(max_frames, h->buf.brMax);
  return h->buf.brUse;
}

int plvi_line_set_profile(plvi_line* h, int on) {
  if (!h) return PLVI_ERR_INVALID;
  h->prof.on = on != 0;
  return PLVI_OK;
}

const char* plvi_line_profile(plvi_line* h) {
  if (!h) return "";
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->profText = h->prof.report();
  return h->profText.c_str();
}

int plvi_line_extract_batch(plvi_line* h, const uint8_t* imgs, int n, int w, int hh, int stride, size_t frame_stride,
                            plvi_keyline* kl, uint8_t* desc, double* line_eq, int* counts) {
  int rc = plvi_line_extract_batch_async(h, imgs, n, w, hh, stride, frame_stride, kl, desc, line_eq, counts);
  if (rc) return rc;
  return plvi_line_sync(h);
}

// what: 0 scaled f64 image (needs set_debug), 1 angle degrees f32 (-1024 = NOTDEF), 2 gradient magnitude f64,
// 3 raw segments float4 (out holds cap entries; *count receives the number), 4 pyramid octave u8 (dense w x h),
// 5 LBD octave image u8 (dense lw x lh), 6 LBD Sobel gradients s16 (lw x lh x {dx, dy})
int plvi_line_read_lsd(plvi_line* h, int frame, int octave, int what, void* out, int cap, int* count) {
  if (!h || !out || frame < 0 || frame >= h->lastN || octave < 0 || octave >= h->nlevels || h->curW < 0) return PLVI_ERR_INVALID;
  PLVI_CUDA_TRY(cudaSetDevice(h->device));
  PLVI_CUDA_TRY(cudaStreamSynchronize(h->stream));
  const LineGeom& g = h->geom;
  const LineOct& O = g.o[octave];
  const size_t npx = (size_t)O.sw * O.sh, pb = (size_t)frame * g.pxTotal + O.pxOff;
  switch (what) {
    case 0:
      if (!h->buf.scaledDbg) { set_error("debug not enabled"); return PLVI_ERR_INVALID; }
      PLVI_CUDA_TRY(cudaMemcpy(out, h->buf.scaledDbg + pb, npx * sizeof(double), cudaMemcpyDeviceToHost));
      break;
    case 1:
      PLVI_CUDA_TRY(cudaMemcpy(out, h->buf.ang + pb, npx * sizeof(float), cudaMemcpyDeviceToHost));
      break;
    case 2: PLVI_CUDA_TRY(cudaMemcpy(out, h->buf.mod + pb, npx * sizeof(double), cudaMemcpyDeviceToHost)); break;
    case 3: {
      int n = 0;
      PLVI_CUDA_TRY(cudaMemcpy(&n, h->buf.regCount + frame * 2 + octave, sizeof(int), cudaMemcpyDeviceToHost));
      if (count) *count = n;
      const int m = std::max(0, std::min(n, cap));
      PLVI_CUDA_TRY(cudaMemcpy(out, h->buf.segs + (size_t)frame * g.segTotal + O.segOff, m * sizeof(float4), cudaMemcpyDeviceToHost));
      break;
    }
    case 4:
      PLVI_CUDA_TRY(cudaMemcpy2D(out, O.w, h->lastPtrs.img[octave] + (size_t)frame * h->lastPtrs.ifs[octave],
                                 h->lastPtrs.ipitch[octave], O.w, O.h, cudaMemcpyDeviceToHost));
      break;
    case 5: {  // LBD pyramid octave (Gaussian 5x5 of the frame / its pyrDown), dense lw x lh u8
      const u8* img = octave == 0 ? h->buf.lbdImg0 : h->buf.lbdImg1;
      PLVI_CUDA_TRY(cudaMemcpy2D(out, O.lw, img + (size_t)frame * O.lpitch * O.lh, O.lpitch, O.lw, O.lh, cudaMemcpyDeviceToHost));
      break;
    }
    case 6:    // Sobel (dx, dy) of the LBD octave, lw x lh x 2 s16
      PLVI_CUDA_TRY(cudaMemcpy(out, h->buf.grad + (size_t)frame * g.lbdTotal + O.lbdOff, (size_t)O.lw * O.lh * sizeof(short2),
                               cudaMemcpyDeviceToHost));
      break;
    case 7:    // band-run diagnostics of (frame, octave): BR_FLAGS ints = fallback, converged, bands dirty in round r at [2 + r]
      if (!h->buf.brMax || frame >= h->buf.brMax || cap < BR_FLAGS) { set_error("band-run path not active for this frame"); return PLVI_ERR_INVALID; }
      PLVI_CUDA_TRY(cudaMemcpy(out, h->buf.brFlags + ((size_t)frame * 2 + octave) * BR_FLAGS, BR_FLAGS * sizeof(int), cudaMemcpyDeviceToHost));
      if (count) *count = h->buf.brRounds;
      break;
    case 8: {  // band-run per-band state of (frame, octave): 8 ints per band (nrec[2], cur, dirty, hasPrev, pixels, runs, cycles of the last run)
      if (!h->buf.brMax || frame >= h->buf.brMax || cap < O.brBands * 8) { set_error("band-run path not active for this frame"); return PLVI_ERR_INVALID; }
      PLVI_CUDA_TRY(cudaMemcpy(out, h->buf.brState + ((size_t)frame * g.brBandsPerFrame + O.brBandOff) * 8, (size_t)O.brBands * 8 * sizeof(int), cudaMemcpyDeviceToHost));
      if (count) *count = O.brBands;
      break;
    }
    default: return PLVI_ERR_INVALID;
  }
  return PLVI_OK;
}

}  // extern "C"
