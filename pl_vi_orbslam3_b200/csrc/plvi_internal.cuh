// Internal declarations shared by the CUDA translation units of libplvi_cuda.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
struct StageProf {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> names;
  int n = 0;
  void begin(cudaStream_t st) { n = 0; names.clear(); if (on) mark(nullptr, st); }
  void mark(const char* name, cudaStream_t st) {
    if (!on) return;
    if ((int)ev.size() <= n) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
    cudaEventRecord(ev[n], st);
    if (name) names.push_back(name);
    n++;
  }
  // after the stream has been synchronised: "name=ms;name=ms;..." (same-named stages summed)
  std::string report() {
    std::string out;
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

int launch_orb_pipeline(const OrbGeom& g, const OrbPtrs& p, const OrbScratch& s, int n, int lap0,
                        int lap1, plvi_keypoint* d_kps, uint8_t* d_desc, int* d_counts,
                        int* d_mono, int cap, cudaStream_t st, int* launches, StageProf* prof);
int orb_kernel_attrs(const OrbGeom& g, int* fastSmem, int* octSmem);

}  // namespace plvi
