// Internal declarations shared by the CUDA translation units of libplvi_cuda.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>

#include <string>
#include <utility>
#include <vector>

#include "../../include/plvi.h"

namespace plvi {

typedef unsigned char u8;

// ---- error plumbing --------------------------------------------------------------
void set_error(const std::string& msg);
#define PLVI_CUDA_TRY(expr)                                                              \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      ::plvi::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));            \
      return PLVI_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)

// ---- optional per-kernel timing (CUDA events on the launching stream) ----------------
// ---------------------------------------------------------------------------------------------------
// The host libm functions the reference calls on float arguments (cos / sin / atan2 with <cmath>'s overloads in
// scope: src/ORBextractor.cc:110-111, src/LSD/lsd.cpp:678-679, binary_descriptor_custom.cpp:1131-1132,
// LSDDetector_custom.cpp:336), restated operation by operation so that the results are the ones of glibc
// 2.28 .. 2.40 on an x86-64 CPU with FMA3 (see oracle/oracle_common.h, namespace glibcm, for the provenance and the
// exhaustive checks against the build image's libm).  Valid for |x| < 120.
// ---------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// sinf and cosf of the same argument share the reduction; each needs one of the two polynomials (which one is
// decided by the quadrant n), so both are evaluated once for all lanes and swapped by n & 1: no divergence.
__device__ __forceinline__ void glibc_sincosf(float y, float& s, float& c) {
  const double x = (double)y;
  const unsigned top = (__float_as_uint(y) >> 20) & 0x7ffu;
  // quadrant reduction (for |y| < pi/4 it yields n = 0 and xr = x, i.e. glibc's short path)
  const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
  const int n = (__double2int_rz(r) + 0x800000) >> 24;
  const double xr = __fma_rn(-(double)n, 0x1.921FB54442D18p0, x);
  const double sgn = ((n + 1) & 2) ? -1.0 : 1.0;          // sign table {1, -1, -1, 1}[n & 3]
  const double sg = (n & 2) ? -1.0 : 1.0;                 // second coefficient table = negated cosine coefficients
  const double xs = __dmul_rn(xr, sgn), x2 = __dmul_rn(xr, xr);
  // sine polynomial of xs
  const double x3 = __dmul_rn(xs, x2), s1 = __fma_rn(x2, -0x1.994eb3774cf24p-13, 0x1.1107605230bc4p-7);
  const double x7 = __dmul_rn(x3, x2), sp = __fma_rn(x3, -0x1.555545995a603p-3, xs);
  const float fs = (float)__fma_rn(x7, s1, sp);
  // cosine polynomial
  const double x4 = __dmul_rn(x2, x2);
  const double c2 = __fma_rn(x2, sg * 0x1.99343027bf8c3p-16, sg * -0x1.6c087e89a359dp-10);
  const double c1 = __fma_rn(x2, sg * -0x1.ffffffd0c621cp-2, sg);
  const double x6 = __dmul_rn(x4, x2);
  const float fc = (float)__fma_rn(x6, c2, __fma_rn(x4, sg * 0x1.55553e1068f19p-5, c1));
  s = (n & 1) ? fc : fs;
  c = (n & 1) ? fs : fc;
  if (top < 0x398u) { s = y; c = 1.0f; }   // |y| < 2^-12
}
__device__ __forceinline__ float glibc_atanf(float x) {
  const int hx = __float_as_int(x), ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) {
    if (ix > 0x7f800000) return __fadd_rn(x, x);
    return hx > 0 ? __fadd_rn(1.5707962513e+00f, 7.5497894159e-08f) : __fsub_rn(-1.5707962513e+00f, 7.5497894159e-08f);
  }
  if (ix < 0x3ee00000) {
    if (ix < 0x31000000) return x;
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) { id = 0; x = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, x), 1.0f), __fadd_rn(2.0f, x)); }
      else { id = 1; x = __fdiv_rn(__fsub_rn(x, 1.0f), __fadd_rn(x, 1.0f)); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = __fdiv_rn(__fsub_rn(x, 1.5f), __fadd_rn(1.0f, __fmul_rn(1.5f, x))); }
      else { id = 3; x = __fdiv_rn(-1.0f, x); }
    }
  }
  const float z = __fmul_rn(x, x), w = __fmul_rn(z, z);
  float s1 = __fadd_rn(4.9768779427e-02f, __fmul_rn(w, 1.6285819933e-02f));
  s1 = __fadd_rn(6.6610731184e-02f, __fmul_rn(w, s1));
  s1 = __fadd_rn(9.0908870101e-02f, __fmul_rn(w, s1));
  s1 = __fadd_rn(1.4285714924e-01f, __fmul_rn(w, s1));
  s1 = __fadd_rn(3.3333334327e-01f, __fmul_rn(w, s1));
  s1 = __fmul_rn(z, s1);
  float s2 = __fadd_rn(-5.8335702866e-02f, __fmul_rn(w, -3.6531571299e-02f));
  s2 = __fadd_rn(-7.6918758452e-02f, __fmul_rn(w, s2));
  s2 = __fadd_rn(-1.1111110449e-01f, __fmul_rn(w, s2));
  s2 = __fadd_rn(-2.0000000298e-01f, __fmul_rn(w, s2));
  s2 = __fmul_rn(w, s2);
  const float xs = __fmul_rn(x, __fadd_rn(s1, s2));
  if (id < 0) return __fsub_rn(x, xs);
  const float hi = id == 0 ? 4.6364760399e-01f : id == 1 ? 7.8539812565e-01f : id == 2 ? 9.8279368877e-01f : 1.5707962513e+00f;
  const float lo = id == 0 ? 5.0121582440e-09f : id == 1 ? 3.7748947079e-08f : id == 2 ? 3.4473217170e-08f : 7.5497894159e-08f;
  const float r = __fsub_rn(hi, __fsub_rn(__fsub_rn(xs, lo), x));
  return hx < 0 ? -r : r;
}
__device__ __forceinline__ float glibc_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  const int hx = __float_as_int(x), hy = __float_as_int(y);
  const int ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return __fadd_rn(x, y);
  if (hx == 0x3f800000) return glibc_atanf(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) return m < 2 ? y : (m == 2 ? __fadd_rn(pi, tiny) : __fsub_rn(-pi, tiny));
  if (ix == 0) return hy < 0 ? __fsub_rn(-pi_o_2, tiny) : __fadd_rn(pi_o_2, tiny);
  const int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = __fadd_rn(pi_o_2, __fmul_rn(0.5f, pi_lo));
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = glibc_atanf(fabsf(__fdiv_rn(y, x)));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return __fsub_rn(pi, __fsub_rn(z, pi_lo));
    default: return __fsub_rn(__fsub_rn(z, pi_lo), pi);
  }
}
#endif

struct StageProf {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> names;
  int n = 0;
  void begin(cudaStream_t st) { n = 0; names.clear(); if (on) { if (timeline()) origin(st); mark(nullptr, st); } }
  void mark(const char* name, cudaStream_t st) {
    if (!on) return;
    if ((int)ev.size() <= n) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
    cudaEventRecord(ev[n], st);
    if (name) names.push_back(name);
    n++;
  }
  // PLVI_TIMELINE=1 (diagnostic, tools/timeline.py): the stages keep their streams (no serialisation by the caller) and
  // report() gives "name=start,end;" in ms since one process-wide origin event instead of durations
  static bool timeline() { static int t = -1; if (t < 0) { const char* e = getenv("PLVI_TIMELINE"); t = (e && e[0] == '1') ? 1 : 0; } return t == 1; }
  static cudaEvent_t origin(cudaStream_t st) {
    static cudaEvent_t o = nullptr;
    if (!o) { cudaEventCreate(&o); cudaEventRecord(o, st); }
    return o;
  }
  // after the stream has been synchronised: "name=ms;name=ms;..." (same-named stages summed)
  std::string report() {
    std::string out;
    if (timeline()) {
      for (int i = 1; i < n; i++) {
        float t0 = 0.f, t1 = 0.f;
        if (cudaEventElapsedTime(&t0, origin(nullptr), ev[i - 1]) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&t1, origin(nullptr), ev[i]) != cudaSuccess) continue;
        out += std::string(names[i - 1]) + "=" + std::to_string(t0) + "," + std::to_string(t1) + ";";
      }
      return out;
    }
    std::vector<std::pair<std::string, float>> acc;
    for (int i = 1; i < n; i++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev[i - 1], ev[i]) != cudaSuccess) continue;
      bool found = false;
      for (auto& a : acc) if (a.first == names[i - 1]) { a.second += ms; found = true; }
      if (!found) acc.push_back({names[i - 1], ms});
    }
    for (auto& a : acc) out += a.first + "=" + std::to_string(a.second) + ";";
    return out;
  }
  ~StageProf() { for (auto e : ev) cudaEventDestroy(e); }
};

// ---------------------------------------------------------------------------------------------------
// CUDA-graph cache of a handle's per-batch launch sequence.  The sequence is static for a given (batch size,
// geometry, pointer set): it is captured once from the launching stream (the side stream of the line pipeline
// joins the capture through its fork / join events) and replayed with one cudaGraphLaunch afterwards.  Keys are
// the values baked into the kernel parameters; a handle that alternates between a few buffer sets (the batched
// front-end does) keeps one executable graph per set.  Disabled while profiling (events between the launches),
// by PLVI_GRAPHS=0, and cleared whenever the handle's device tables change.
// ---------------------------------------------------------------------------------------------------
struct GraphCache {
  struct Entry { std::vector<uint64_t> key; cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr; int launches = 0; };
  std::vector<Entry> entries;
  int enabled = -1;
  long replays = 0, captures = 0;
  static const size_t kMaxEntries = 8;
  bool on() {
    if (enabled < 0) { const char* e = getenv("PLVI_GRAPHS"); enabled = (e && e[0] == '0') ? 0 : 1; }
    return enabled == 1;
  }
  void clear() {
    for (auto& e : entries) { cudaGraphExecDestroy(e.exec); cudaGraphDestroy(e.graph); }
    entries.clear();
  }
  ~GraphCache() { clear(); }
  // record(launches*) issues the launch sequence on `st`; returns a PLVI status
  template <class F> int run(cudaStream_t st, const std::vector<uint64_t>& key, int* launches, F&& record) {
    // a caller that is itself capturing this stream (e.g. into a larger graph) gets plain launches: they become
    // nodes of the caller's graph
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) return record(launches);
    for (auto& e : entries)
      if (e.key == key) {
        PLVI_CUDA_TRY(cudaGraphLaunch(e.exec, st));
        if (launches) *launches = e.launches;
        replays++;
        return PLVI_OK;
      }
    Entry e;
    e.key = key;
    e.launches = 0;
    PLVI_CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = record(&e.launches);
    cudaGraph_t gph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &gph);
    if (rc != PLVI_OK) { if (gph) cudaGraphDestroy(gph); return rc; }
    PLVI_CUDA_TRY(ce);
    e.graph = gph;
    if (cudaGraphInstantiate(&e.exec, e.graph, 0) != cudaSuccess) {
      // no executable graph: give the captured one back and run the sequence as plain launches
      cudaGetLastError();
      cudaGraphDestroy(gph);
      return record(launches);
    }
    if (entries.size() >= kMaxEntries) {
      cudaGraphExecDestroy(entries.front().exec);
      cudaGraphDestroy(entries.front().graph);
      entries.erase(entries.begin());
    }
    entries.push_back(e);
    captures++;
    PLVI_CUDA_TRY(cudaGraphLaunch(e.exec, st));
    if (launches) *launches = e.launches;
    return PLVI_OK;
  }
};

// ---------------------------------------------------------------------------------------------------
// Input staging of the host-buffer ("async") entry points: two device buffers filled by a copy stream, so that the
// upload of call i + 1 overlaps the kernels of call i (the caller issues the next call without waiting; plvi_*_sync
// waits for everything).  Buffer 0 is the handle's own level-0 image, buffer 1 is allocated on first use.
// ---------------------------------------------------------------------------------------------------
struct AsyncInput {
  u8* buf[2] = {nullptr, nullptr};
  int sel = 0;
  cudaStream_t copy = nullptr;
  cudaEvent_t h2d[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, shared[2] = {nullptr, nullptr};
  bool doneValid[2] = {false, false}, sharedValid[2] = {false, false};
  // the last upload (what another handle may read instead of uploading the same frames again, see share_last)
  int lastSel = -1, lastPitch = 0;
  size_t lastFs = 0;
  // the buffer to upload into (on `copy`), after the call that last read it has finished
  int begin(u8* first, size_t bytes, u8** dst) {
    if (!copy) {
      PLVI_CUDA_TRY(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
      for (int k = 0; k < 2; k++) {
        PLVI_CUDA_TRY(cudaEventCreateWithFlags(&h2d[k], cudaEventDisableTiming));
        PLVI_CUDA_TRY(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
        PLVI_CUDA_TRY(cudaEventCreateWithFlags(&shared[k], cudaEventDisableTiming));
      }
      buf[0] = first;
    }
    if (sel == 1 && !buf[1]) PLVI_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&buf[1]), bytes + 256));
    if (doneValid[sel]) PLVI_CUDA_TRY(cudaStreamWaitEvent(copy, done[sel], 0));
    if (sharedValid[sel]) PLVI_CUDA_TRY(cudaStreamWaitEvent(copy, shared[sel], 0));   // a second reader of that buffer
    *dst = buf[sel];
    return PLVI_OK;
  }
  int uploaded(cudaStream_t compute) {   // the kernels wait for the upload
    PLVI_CUDA_TRY(cudaEventRecord(h2d[sel], copy));
    PLVI_CUDA_TRY(cudaStreamWaitEvent(compute, h2d[sel], 0));
    return PLVI_OK;
  }
  int finish(cudaStream_t compute) {     // everything that reads the buffer has been enqueued
    PLVI_CUDA_TRY(cudaEventRecord(done[sel], compute));
    doneValid[sel] = true;
    lastSel = sel;
    sel ^= 1;
    return PLVI_OK;
  }
  // Another handle reads the frames of the last upload on its own stream `reader`: it waits for the upload, and the
  // buffer is not overwritten before share_done() has been recorded on that stream.
  int share_last(cudaStream_t reader, const u8** img, int* pitch, size_t* fs) {
    if (lastSel < 0) return PLVI_ERR_INVALID;
    PLVI_CUDA_TRY(cudaStreamWaitEvent(reader, h2d[lastSel], 0));
    *img = buf[lastSel]; *pitch = lastPitch; *fs = lastFs;
    return PLVI_OK;
  }
  int share_done(cudaStream_t reader) {
    PLVI_CUDA_TRY(cudaEventRecord(shared[lastSel], reader));
    sharedValid[lastSel] = true;
    return PLVI_OK;
  }
  // Uploads n host frames (w x hh, row stride `stride`, frame stride `frame_stride`) into dst on the copy stream and
  // reports the device layout.  2-D copies of 752-byte rows reach a fraction of the PCIe rate (measured: 6.4 GB/s
  // against > 40 GB/s for linear copies), so host frames whose rows fit the device pitch are uploaded with their own
  // row stride as plain linear copies (one copy when the frames are contiguous too); the kernels take any pitch.
  int upload(u8* dst, int devPitch, int devRows, const uint8_t* imgs, int n, int w, int hh, int stride, size_t frame_stride,
             int* pitch, size_t* fs) {
    if (stride <= devPitch) {
      const size_t fbytes = (size_t)stride * hh;
      *pitch = stride;
      *fs = fbytes;
      lastPitch = stride; lastFs = fbytes;
      if (frame_stride == fbytes) {
        PLVI_CUDA_TRY(cudaMemcpyAsync(dst, imgs, fbytes * n, cudaMemcpyHostToDevice, copy));
      } else {
        for (int i = 0; i < n; i++)
          PLVI_CUDA_TRY(cudaMemcpyAsync(dst + (size_t)i * fbytes, imgs + (size_t)i * frame_stride, fbytes, cudaMemcpyHostToDevice, copy));
      }
      return PLVI_OK;
    }
    *pitch = devPitch;
    *fs = (size_t)devPitch * devRows;
    lastPitch = devPitch; lastFs = *fs;
    for (int i = 0; i < n; i++)
      PLVI_CUDA_TRY(cudaMemcpy2DAsync(dst + (size_t)i * *fs, devPitch, imgs + (size_t)i * frame_stride, stride, w, hh,
                                      cudaMemcpyHostToDevice, copy));
    return PLVI_OK;
  }
  void destroy() {
    if (copy) { cudaStreamSynchronize(copy); cudaStreamDestroy(copy); copy = nullptr; }
    for (int k = 0; k < 2; k++) {
      if (h2d[k]) cudaEventDestroy(h2d[k]);
      if (done[k]) cudaEventDestroy(done[k]);
      if (shared[k]) cudaEventDestroy(shared[k]);
      h2d[k] = done[k] = shared[k] = nullptr;
    }
    cudaFree(buf[1]);
    buf[1] = nullptr;
  }
};

// ---------------------------------------------------------------------------------------------------
// Result copies of the host-buffer ("async") entry points: two device result sets and a device-to-host stream, so that
// the copy of call i's results to the caller's buffers overlaps the kernels of call i + 1 (which write the other set)
// instead of sitting between them on the compute stream.  Call i + 2 reuses set i after its copy has finished.
// ---------------------------------------------------------------------------------------------------
struct AsyncOutput {
  cudaStream_t d2h = nullptr;
  cudaEvent_t ready = nullptr, done[2] = {nullptr, nullptr};
  bool valid[2] = {false, false};
  int sel = 0, last = 0;
  int begin(cudaStream_t compute) {      // the result set this call writes is free again
    if (!d2h) {
      PLVI_CUDA_TRY(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
      PLVI_CUDA_TRY(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
      for (int k = 0; k < 2; k++) PLVI_CUDA_TRY(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
    }
    if (valid[sel]) PLVI_CUDA_TRY(cudaStreamWaitEvent(compute, done[sel], 0));
    return PLVI_OK;
  }
  int start_copy(cudaStream_t compute) {   // the copies (issued on d2h by the caller) wait for the kernels
    PLVI_CUDA_TRY(cudaEventRecord(ready, compute));
    PLVI_CUDA_TRY(cudaStreamWaitEvent(d2h, ready, 0));
    return PLVI_OK;
  }
  int end_copy() {
    PLVI_CUDA_TRY(cudaEventRecord(done[sel], d2h));
    valid[sel] = true;
    last = sel;
    sel ^= 1;
    return PLVI_OK;
  }
  cudaEvent_t last_done() const { return valid[last] ? done[last] : nullptr; }
  int sync() {
    if (d2h) PLVI_CUDA_TRY(cudaStreamSynchronize(d2h));
    return PLVI_OK;
  }
  void destroy() {
    if (d2h) { cudaStreamSynchronize(d2h); cudaStreamDestroy(d2h); d2h = nullptr; }
    if (ready) cudaEventDestroy(ready);
    for (int k = 0; k < 2; k++) if (done[k]) cudaEventDestroy(done[k]);
    ready = done[0] = done[1] = nullptr;
  }
};

// ---- ORB geometry (ORBextractor ctor + ComputeKeyPointsOctTree grid) ----------------
static const int kEdge = 16;      // minBorder = EDGE_THRESHOLD - 3, src/ORBextractor.cc:771
static const int kMaxNodesMin = 8;

struct OrbLevel {
  int w, h;
  int pitch;                 // row pitch (bytes) of the handle-owned level image + blurred copy
  int nCols, nRows, wCell, hCell;  // FAST cell grid, src/ORBextractor.cc:779-785
  int quota;                 // mnFeaturesPerLevel
  int kpOff, kpCap;          // slot range of this level in the per-frame staging array
  int candOff, candCap;      // range of this level in the per-frame candidate array
  int nIni;                  // DistributeOctTree root nodes, src/ORBextractor.cc:541
  float hX;
  float scale;               // mvScaleFactor
  int sizeField;             // (int)(PATCH_SIZE * scale), src/ORBextractor.cc:862
  int rsOff;                 // offset of this level's resize coefficient rows (x then y)
};

struct OrbGeom {
  int nlevels, iniTh, minTh;
  int candTotal, kpTotal;
  int maxNodes;              // octree node capacity (max over levels of max(quota+3, 8) + slack)
  OrbLevel lv[PLVI_MAX_LEVELS];
};

// Per-call pointer table.  img[0] may alias caller-owned device memory.
struct OrbPtrs {
  const u8* img[PLVI_MAX_LEVELS];
  u8* blur[PLVI_MAX_LEVELS];
  int ipitch[PLVI_MAX_LEVELS];
  size_t ifs[PLVI_MAX_LEVELS];   // frame stride of img[l]
  size_t bfs[PLVI_MAX_LEVELS];   // frame stride of blur[l] (pitch = lv[l].pitch)
};

struct FastTile {   // one CTA of k_fast: `ncells` horizontally adjacent cells of one cell row
  unsigned short level, cellRow, cellCol0, ncells;
};
struct BlurTile {
  unsigned short level, tx, ty, pad;
};

// candidate / staged keypoint packing: x | y << 12 | score << 24
__host__ __device__ inline uint32_t pack_xys(int x, int y, int s) {
  return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24);
}
__host__ __device__ inline int unpack_x(uint32_t p) { return p & 0xFFF; }
__host__ __device__ inline int unpack_y(uint32_t p) { return (p >> 12) & 0xFFF; }
__host__ __device__ inline int unpack_s(uint32_t p) { return p >> 24; }

// ---- kernel launchers (orb_kernels.cu) ----------------------------------------------
struct OrbScratch {
  uint32_t* cand;      // [B][candTotal]
  int* candCount;      // [B][nlevels]
  uint16_t* knode;     // [B][candTotal]
  uint32_t* lvlKp;     // [B][kpTotal]  staged keypoints per level (absolute level coords)
  int* lvlCount;       // [B][nlevels]
  int* slot;           // [B][kpTotal]  output row of each staged keypoint
  const int2* rsTab;   // resize coefficient table
  const FastTile* fastTiles; int nFastTiles;
  const BlurTile* blurTiles; int nBlurTiles;
  int fastSmem, octSmem;
};

int launch_stereo(const OrbGeom& g, const OrbPtrs& L, const OrbPtrs& R, const float* scale, const float* invScale, int n,
                  const plvi_keypoint* kl, const uint8_t* dl, const int* nl, const plvi_keypoint* kr, const uint8_t* dr,
                  const int* nr, int stride, float mb, float mbf, float* uRight, float* depth, int* sad, int* nstereo,
                  cudaStream_t st);
int launch_orb_pipeline(const OrbGeom& g, const OrbPtrs& p, const OrbScratch& s, int n, int lap0,
                        int lap1, plvi_keypoint* d_kps, uint8_t* d_desc, int* d_counts,
                        int* d_mono, int cap, cudaStream_t st, int* launches, StageProf* prof, cudaEvent_t waitAfterPyramid = nullptr, int stages = 3);
int orb_kernel_attrs(const OrbGeom& g, int* fastSmem, int* octSmem);

}  // namespace plvi
