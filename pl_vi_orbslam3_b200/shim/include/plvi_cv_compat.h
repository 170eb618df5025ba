// Types the drop-in headers share.  With -DPLVI_HAVE_OPENCV (a SLAM build: the reference's CMake adds this
// directory in front of its own include/) the real <opencv2/core.hpp>, Eigen and the reference's
// line_descriptor headers are used and the classes are drop-ins for include/ORBextractor.h,
// include/LineExtractor.h, include/ORBmatcher.h, include/LineMatcher.h of the reference.  Without it (this
// image has no OpenCV C++) layout-identical POD stand-ins let the descriptor-level API be built and smoke
// tested on its own (shim_smoke.cpp).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/plvi.h"

#ifdef PLVI_HAVE_OPENCV
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
#include <line_descriptor_custom.hpp>
#include <line_descriptor/descriptor_custom.hpp>
#ifndef PLVI_HAVE_EIGEN
#define PLVI_HAVE_EIGEN
#endif
#else
#define CV_8UC1 0
namespace cv {
struct Point2f { float x = 0, y = 0; };
struct KeyPoint {  // same 28-byte layout as cv::KeyPoint
  Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
};
// 8-bit single-channel (or N x 32 descriptor) matrix view/owner
struct Mat {
  int rows = 0, cols = 0; size_t step = 0; uint8_t* data = nullptr; std::vector<uint8_t> own;
  Mat() {}
  Mat(int r, int c, uint8_t* d, size_t s = 0) : rows(r), cols(c), step(s ? s : (size_t)c), data(d) {}
  void create(int r, int c, int /*type*/ = CV_8UC1) { rows = r; cols = c; step = (size_t)c; own.assign((size_t)r * c, 0); data = own.data(); }
  void release() { rows = cols = 0; step = 0; data = nullptr; own.clear(); }
  bool empty() const { return !data || rows == 0 || cols == 0; }
  int type() const { return CV_8UC1; }
  bool isContinuous() const { return rows <= 1 || step == (size_t)cols; }
  uint8_t* ptr(int r = 0) { return data + (size_t)r * step; }
  const uint8_t* ptr(int r = 0) const { return data + (size_t)r * step; }
  Mat row(int r) const { return Mat(1, cols, const_cast<uint8_t*>(ptr(r)), step); }
  // the two proxy calls the extractors make on their arguments
  Mat getMat() const { return Mat(rows, cols, data, step); }
  Mat& getMatRef() const { return const_cast<Mat&>(*this); }
};
typedef const Mat& InputArray;
typedef const Mat& OutputArray;   // the stand-in writes through getMatRef()
namespace line_descriptor {
struct KeyLine {  // 68-byte layout of descriptor_custom.hpp:107-146
  float angle; int class_id; int octave; Point2f pt; float response; float size;
  float startPointX, startPointY, endPointX, endPointY;
  float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
  float lineLength; int numOfPixels;
};
}  // namespace line_descriptor
}  // namespace cv
#endif

#ifndef PLVI_HAVE_EIGEN
namespace Eigen {
struct Vector3d { double v[3]; double& operator()(int i) { return v[i]; } double operator()(int i) const { return v[i]; } };
}  // namespace Eigen
#else
#include <Eigen/Core>
#endif

static_assert(sizeof(cv::KeyPoint) == sizeof(plvi_keypoint), "cv::KeyPoint layout");
static_assert(sizeof(cv::line_descriptor::KeyLine) == sizeof(plvi_keyline), "KeyLine layout");

namespace plvi_shim {
inline void check(int rc, const char* what) {
  if (rc < 0) throw std::runtime_error(std::string(what) + ": " + plvi_last_error());
}

// rows x 32 descriptor matrix -> contiguous bytes (a cv::Mat of descriptors may be a non-continuous ROI)
inline const uint8_t* packed_rows(const cv::Mat& d, int rows, std::vector<uint8_t>& tmp) {
  if (rows <= 0) return nullptr;
  if (d.isContinuous()) return d.ptr(0);
  tmp.resize((size_t)rows * 32);
  for (int i = 0; i < rows; i++) std::memcpy(&tmp[(size_t)i * 32], d.ptr(i), 32);
  return tmp.data();
}

// One matcher handle per host thread: Tracking, LocalMapping and LoopClosing run their searches concurrently and a
// handle (one CUDA stream + staging) is not re-entrant.  Sized for 16 k features per frame; staging is allocated on
// first use.
class MatcherHandle {
 public:
  static plvi_matcher* get() {
    thread_local MatcherHandle inst;
    return inst.h_;
  }
 private:
  MatcherHandle() { check(plvi_matcher_create(&h_, 1, 16384, 16384, 0, nullptr), "plvi_matcher_create"); }
  ~MatcherHandle() { plvi_matcher_destroy(h_); }
  plvi_matcher* h_ = nullptr;
};
}  // namespace plvi_shim
