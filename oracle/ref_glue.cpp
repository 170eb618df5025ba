// TEST INFRASTRUCTURE -- C entry points over the reference's own classes, compiled together with the
// reference's UNMODIFIED sources (oracle/Makefile.ref) into oracle/_ref/libplvi_ref.so.  Used only by
// tests/ (to pin the oracle restatement against the code its authors wrote) and by bench.py's CPU arms.
// The OpenCV / Eigen layer underneath is the stand-in of oracle/cvmini/ (see cvmini.hpp).
#include <cstring>
#include <vector>

#include <Eigen/Core>          // cvmini/ stand-in
#include "DBoW2/FORB.h"                  // /root/reference/Thirdparty/DBoW2
#include "DBoW2/TemplatedVocabulary.h"
#include "ORBextractor.h"      // /root/reference/include
#include "LineExtractor.h"     // /root/reference/include

using cv::line_descriptor::KeyLine;

// ------------------------------------------------------------------------------------------------
// Monotone allocator.  DistributeOctTree sorts pair<int, ExtractorNode*> (src/ORBextractor.cc:682): nodes
// holding the same number of keys are ordered by their HEAP ADDRESS, which with a general-purpose malloc
// depends on the whole allocation history of the process.  The oracle (and the CUDA path) define that tie
// as "node created earlier first".  While a plviref_* call runs, operator new inside this library (it is
// linked with -Bsymbolic-functions, so only this library is affected) hands out strictly increasing
// addresses and never reuses one, which makes the reference's own code realise exactly that definition.
// ------------------------------------------------------------------------------------------------
#include <sys/mman.h>
#include <cstdlib>
#include <new>
#ifdef PLVI_DROPIN
// drop-in build (Makefile.ref, "DROP-IN PROOF"): ORBextractor.h / LineExtractor.h are the product's headers, the work
// happens in libplvi_cuda.so; only the two extractor entry points below are compiled
namespace { struct ArenaScope {}; }
#else
namespace {
// one lazily committed virtual range per call: addresses only ever increase inside it
const size_t kArenaBytes = (size_t)16 << 30;
thread_local bool g_arena_on = false;
thread_local char* g_base = nullptr;
thread_local char* g_cur = nullptr;

void* arena_alloc(size_t n) {
  n = (n + 15) & ~(size_t)15;
  if ((size_t)(g_base + kArenaBytes - g_cur) < n) { fprintf(stderr, "libplvi_ref: arena exhausted\n"); abort(); }
  void* p = g_cur;
  g_cur += n;
  return p;
}
inline bool arena_owns(void* p) { return g_base && (char*)p >= g_base && (char*)p < g_base + kArenaBytes; }
bool g_monotone = true;   // plviref_set_monotone(0): plain malloc (timing runs; ties then follow malloc's addresses)
struct ArenaScope {   // one per plviref_* call: everything allocated inside is released at the end
  ArenaScope() {
    if (!g_monotone) return;
    void* m = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (m == MAP_FAILED) { fprintf(stderr, "libplvi_ref: cannot reserve the arena\n"); abort(); }
    g_base = g_cur = (char*)m;
    g_arena_on = true;
  }
  ~ArenaScope() {
    if (!g_base) return;
    g_arena_on = false;
    munmap(g_base, kArenaBytes);
    g_base = g_cur = nullptr;
  }
};
}  // namespace

void* operator new(size_t n) {
  if (g_arena_on) return arena_alloc(n);
  void* p = malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { if (p && !arena_owns(p)) free(p); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }
#endif  // PLVI_DROPIN

static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint POD layout");
static_assert(sizeof(KeyLine) == 68, "KeyLine POD layout");

extern "C" {

#ifndef PLVI_DROPIN
void plviref_set_monotone(int on) { g_monotone = on != 0; }
#else
void plviref_set_monotone(int) {}
#endif

// ORB_SLAM3::ORBextractor::operator() (src/ORBextractor.cc:1068-1150).  Returns the number of keypoints
// (-1: empty image, -2: capacity); *mono_index = the operator's return value.  Optional: pyr_out receives
// the dense level images (mvImagePyramid), level 0 first.
int plviref_orb_extract(const uchar* img, int w, int h, int stride, int nfeatures, float sf, int nlevels, int iniTh,
                        int minTh, int lap0, int lap1, cv::KeyPoint* kps, uchar* desc, int cap, int* mono_index,
                        uchar* pyr_out) {
  ArenaScope scope;
  ORB_SLAM3::ORBextractor ex(nfeatures, sf, nlevels, iniTh, minTh);
  cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)stride);
  std::vector<cv::KeyPoint> keys;
  cv::Mat d;
  std::vector<int> lap = {lap0, lap1};
  const int mono = ex(image, cv::Mat(), keys, d, lap);
  if (mono_index) *mono_index = mono;
  if (mono < 0) return -1;
  if (pyr_out) {
    size_t o = 0;
    for (int l = 0; l < nlevels; l++) {
      const cv::Mat& m = ex.mvImagePyramid[l];
      for (int y = 0; y < m.rows; y++) { memcpy(pyr_out + o, m.ptr(y), m.cols); o += m.cols; }
    }
  }
  const int n = (int)keys.size();
  if (n > cap) return -2;
  if (n) memcpy(kps, keys.data(), sizeof(cv::KeyPoint) * n);
  for (int i = 0; i < n; i++) memcpy(desc + 32 * i, d.ptr(i), 32);
  return n;
}

#ifndef PLVI_DROPIN
// cv::createLineSegmentDetector(...)->detect on one u8 image (src/LSD/lsd.cpp:412-534).  segs: x1,y1,x2,y2.
int plviref_lsd(const uchar* img, int stride, int w, int h, int refine, float lsd_scale, float* segs, int cap) {
  ArenaScope scope;
  cv::Ptr<cv::LineSegmentDetector> ls = cv::createLineSegmentDetector(refine, lsd_scale, 0.6, 2.0, 22.5, 1.0, 0.6, 1024);
  cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)stride);
  std::vector<cv::Vec4f> lines;
  ls->detect(image, lines);
  const int n = (int)lines.size();
  for (int i = 0; i < std::min(n, cap); i++) memcpy(segs + 4 * i, lines[i].val, 16);
  return n;
}

#endif  // PLVI_DROPIN

// ORB_SLAM3::Lineextractor::operator() (src/LineExtractor.cc:45-117), extractor 0 (LSD + LBD).
int plviref_line_extract(const uchar* img, int w, int h, int stride, int lsd_nfeatures, int lsd_refine, float lsd_scale,
                         int nlevels, float scale, KeyLine* keylines, uchar* desc, double* lineeq, int cap) {
  ArenaScope scope;
  ORB_SLAM3::Lineextractor ex(lsd_nfeatures, lsd_refine, lsd_scale, nlevels, scale, 0);
  cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)stride);
  std::vector<KeyLine> kls;
  cv::Mat d;
  std::vector<Eigen::Vector3d> eq;
  ex(image, cv::Mat(), kls, d, eq);
  const int n = (int)kls.size();
  if (n > cap) return -2;
  if (n) memcpy(keylines, kls.data(), sizeof(KeyLine) * n);
  if (!d.empty())
    for (int i = 0; i < n; i++) memcpy(desc + 32 * i, d.ptr(i), 32);
  for (int i = 0; i < (int)eq.size() && i < n; i++) { lineeq[3 * i] = eq[i](0); lineeq[3 * i + 1] = eq[i](1); lineeq[3 * i + 2] = eq[i](2); }
  return n;
}

}  // extern "C"

#ifndef PLVI_DROPIN
// Frame::ComputeBoW (src/Frame.cc:1115-1122): ORBVocabulary (include/ORBVocabulary.h) = DBoW2's
// TemplatedVocabulary<FORB::TDescriptor, FORB>, loaded with the reference's own loadFromTextFile (the ORBvoc.txt
// format) and applied with transform(features, BowVector, FeatureVector, levelsup).  The two maps are returned in key
// order: BowVector as (word, value) pairs, FeatureVector as CSR (node, start, feature indices).
extern "C" int plviref_bow_transform(const char* voc_text_path, const uchar* desc, int n, int levelsup, int* bow_count, int* bow_words,
                                     double* bow_values, int* fv_count, int* fv_nodes, int* fv_start, int* fv_features) {
  typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;
  ORBVocabulary voc;
  if (!voc.loadFromTextFile(voc_text_path)) return -1;
  std::vector<cv::Mat> vDesc;   // Converter::toDescriptorVector: one 1 x 32 row per feature
  vDesc.reserve(n);
  for (int i = 0; i < n; i++) {
    cv::Mat row(1, 32, CV_8UC1);
    memcpy(row.data, desc + 32 * (size_t)i, 32);
    vDesc.push_back(row);
  }
  DBoW2::BowVector bv;
  DBoW2::FeatureVector fv;
  voc.transform(vDesc, bv, fv, levelsup);
  int k = 0;
  for (auto it = bv.begin(); it != bv.end(); ++it, ++k) { bow_words[k] = (int)it->first; bow_values[k] = it->second; }
  *bow_count = k;
  int f = 0, pos = 0;
  for (auto it = fv.begin(); it != fv.end(); ++it, ++f) {
    fv_nodes[f] = (int)it->first;
    fv_start[f] = pos;
    for (unsigned int idx : it->second) fv_features[pos++] = (int)idx;
  }
  fv_start[f] = pos;
  *fv_count = f;
  return (int)voc.size();
}

// LineMatcher (src/LineMatcher.cpp, compiled unmodified with cvmini/slam_mock.h force-included in place of the
// Frame / KeyFrame / MapLine headers): the descriptor matchers.  plviref_line_mock.cpp holds the calls (it needs the
// stand-in classes, which cannot share a translation unit with the real ORBextractor.h / LineExtractor.h).
// The EDLines detector (extractor: 1, not the shipped configuration) lives in ED_Lib, which is not compiled
// here; LSDDetector_custom.cpp only references these three entry points.  They abort if ever reached.
#define PLVIREF_STUB(fn, sym)                                                         \
  extern "C" void fn() __asm__(sym);                                                  \
  void fn() { fprintf(stderr, "libplvi_ref: EDLines is not part of this build\n"); abort(); }
PLVIREF_STUB(plviref_stub_edlines_ctor0, "_ZN7EDLinesC1Ev")
PLVIREF_STUB(plviref_stub_edlines_ctor1, "_ZN7EDLinesC1EN2cv3MatEdidd")
PLVIREF_STUB(plviref_stub_edlines_getlines, "_ZN7EDLines8getLinesEv")
#endif  // PLVI_DROPIN
