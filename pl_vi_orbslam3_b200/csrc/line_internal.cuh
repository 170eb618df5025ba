// Geometry / buffer tables of the line pipeline (shared by line_kernels.cu and plvi_line_capi.cu).
#pragma once
#include "plvi_internal.cuh"

namespace plvi {

struct LineOct {
  int w, h, pitch;     // LSD pyramid octave image (u8)
  int sw, sh;          // LSD working size after lsd_scale (flsd: resize(gaussian_img, scaled_image, Size(), SCALE, SCALE))
  int wpr;             // "available" bitmap words per row
  int preSW, preRH;    // k_lsd_pre: f64 row stride / max source rows of a tile's source window in shared memory
  int minRegSize;      // int(-LOG_NT / log10(p)), src/LSD/lsd.cpp:466-467
  size_t pxOff;        // offset of this octave in per-frame scaled-pixel arrays
  size_t rawOff;       // offset in the per-frame row-filtered f64 array
  int bmOff;           // offset (words) in the per-frame bitmap array
  int segOff, segCap;  // raw segment slots of this octave per frame
  int xtabOff, ytabOff;
  int lw, lh, lpitch;  // LBD octave image (w >> o, h >> o)
  size_t lbdOff;       // offset in the per-frame gradient array
  // band speculation of the region growing (k_lsd_spec / k_lsd_commit)
  int nbands, bandRows;       // bands of bandRows working rows; band j covers rows [j * bandRows, ...)
  int bandPxCap, bandRecCap;  // pixel-list / record slots of one band
  size_t regOff;              // offset in the per-frame region pixel array: serial area (sw * sh) then the band areas
  size_t specBmOff;           // offset (words) in the per-frame private-bitmap array; band j holds rows [j * bandRows, sh)
  size_t specRecOff;          // offset in the per-frame speculative record array
  int taskOff;                // first speculation task of this octave within a frame
  // band-run region growing for small batches (k_lsd_band_*): brBands bands of brRows rows, one warp each
  int brBands, brRows, brPxCap, brRecCap;
  int brBandOff;              // first band of this octave among the frame's bands
  size_t brBmOff;             // words: bitmap of band j at brBmOff + j * wpr * sh
  size_t brRecOff;            // records (per buffer): band j at brRecOff + j * brRecCap
  size_t brListOff;           // pixels (per buffer): band j at brListOff + j * brPxCap
};

struct LineGeom {
  int noct;
  LineOct o[2];
  size_t pxTotal, rawTotal, lbdTotal;
  size_t regTotal, specBmTotal, specRecTotal;
  int tasksPerFrame;
  int brBandsPerFrame;                       // band-run: bands of both octaves
  size_t brBmTotal, brRecTotal, brListTotal; // per frame (and per buffer)
  int bmTotal, segTotal;
  double rho, prec, lsdScale, minLength;
  float alignHi2, alignLo2;   // cos^2(prec -/+ margin): bounds of the cheap alignment test in k_lsd_grow
  int hk;                     // half width of the LSD Gaussian: h = ceil(sigma * sqrt(2 * 3 * ln 10)), src/LSD/lsd.cpp:452; 0: lsd_scale == 1
  int preTH;                  // k_lsd_pre: scaled rows per tile (tiles are 32 scaled columns wide = one bitmap word)
  double kern[17];            // cv::getGaussianKernel(1 + 2 hk, sigma, CV_64F)
  float lineScale;
  int nfeat, keepCap;
  int refine;                 // lsd_refine: 0 none, 1 standard (refine), 2 advanced (+ NFA), src/LSD/lsd.cpp:493-504
};

struct LineTab { int ofs; float a0, a1; };       // f64 bilinear taps (float32 weights)
struct LineRegion { int start, size; double angle; };
struct SpecRec { unsigned seed, start, size; float angDeg; };   // speculative region: seed x | y << 16, pixel list slice, region angle

struct LinePtrs {
  const u8* img[2];
  int ipitch[2];
  size_t ifs[2];
};

struct LineBufs {
  float* ang;            // [B][pxTotal]   level-line angle in degrees (-1024 = NOTDEF)
  float2* cs;            // [B][pxTotal]   cos, sin of float(angle): what a region accumulates (lsd.cpp:673-675)
  float2* seed;          // [B][pxTotal]   cos, sin of the f64 angle (region seed: float(std::cos(reg_angle)))
  double* mod;           // [B][pxTotal]   gradient magnitude
  unsigned* bitmap;      // [B][bmTotal]   angle defined & not used
  unsigned* reg;         // [B][regTotal]  region pixel lists (x | y << 16)
  unsigned* specBm;      // [B][specBmTotal] private availability bitmaps of the speculation bands
  SpecRec* specRec;      // [B][specRecTotal]
  int* specCnt;          // [B][tasksPerFrame] speculative regions per band
  int* specStart;        // [1] blocks of k_lsd_spec that have started in the current batch (plvi_line_stage_counter)
  int* bandRow;          // [B][tasksPerFrame + 2] first row of every band, octave o from taskOff + o (nbands + 1 entries): equal LOAD per band
  unsigned* phantom;     // [B][bmTotal]   pixels a discarded speculative region had consumed
  int useSpec;           // 0: serial k_lsd_grow only
  int eqLoad;            // speculation bands of equal load (k_lsd_band_split): 1 by batch size, 2 always, 0 never (PLVI_LSD_EQLOAD)
  // band-run (small batches): per band the last input bitmap, the working / output bitmap, the initial phantom
  // bitmap of the next run; two generations of region records + pixel lists; per-band state; per-octave flags
  unsigned* brIn; unsigned* brWk; unsigned* brPh;   // [brMax][brBmTotal]
  uint4* brRec;          // [brMax][2][brRecTotal]  {seed x | y << 16, list start, size, angle (float bits)}
  unsigned* brList;      // [brMax][2][brListTotal]
  int* brState;          // [brMax][brBandsPerFrame][8]  nrec[0], nrec[1], cur, dirty, hasPrev
  int* brFlags;          // [brMax][2][BR_FLAGS]  [0] fallback to the serial kernel, [1] converged, [2 + r] bands dirty in round r
  int brMax;             // frames the band-run buffers hold (0: off)
  int brUse;             // batches of up to brUse (<= brMax) frames take the band-run path (plvi_line_set_band_run_max)
  int brRounds;          // rounds launched per batch
  LineRegion* regTab;    // [B][segTotal]
  int* regCount;         // [B][2]  (-1: segment table overflow)
  float4* segs;          // [B][segTotal]
  float* tmpResp;        // [B][segTotal]
  int* tmpCls;           // [B][segTotal]
  u8* lbdImg0; u8* lbdImg1;
  short2* grad;          // [B][lbdTotal] Sobel (dx, dy)
  float* lbdRows;        // [B][keepCap][8][64] weighted row sums of the LBD support region
  const LineTab* tabs;
  const int2* rsTab;     // u8 bilinear table for pyramid level 1 (x rows then y rows)
  const double* lbdG;    // [63] gaussCoefG_
  const double* lbdL;    // [21] gaussCoefL_
  const double2* trig;   // [1024] {cos, sin}(k * 2 pi / 1024), host libm values
  double* scaledDbg;     // [B][pxTotal] or nullptr
};

#define BR_FLAGS 40
// warps per block of k_lsd_spec: 640 blocks of 4 warps put 5 blocks on 48 of the 148 SMs and 4 on the others at 4096
// frames; blocks of 2 warps spread evenly (54.2 -> 50.4 ms)
#define SPEC_WPB 2

struct LineAux { cudaStream_t stream; cudaEvent_t fork, join; cudaEvent_t stage; };   // stage: recorded when the streaming kernels are done and region growing starts   // side stream for the LBD pre-processing

int launch_line_pipeline(const LineGeom& g, const LinePtrs& p, const LineBufs& b, int n, plvi_keyline* dKl,
                         uint8_t* dDesc, double* dEq, int* dCounts, cudaStream_t st, LineAux aux, int* launches,
                         StageProf* prof);
int line_kernel_attrs(const LineGeom& g);
void launch_resize_u8(const u8* src, int spitch, size_t sfs, int sw, int sh, u8* dst, int dpitch, size_t dfs, int dw,
                      int dh, const int2* xtab, const int2* ytab, int n, cudaStream_t st);

}  // namespace plvi
