// Minimal stand-ins for the OpenCV / Eigen types that appear in the reference signatures,
// used only when the real headers are not available (this image has no OpenCV C++).  With
// -DPLVI_HAVE_OPENCV the real <opencv2/core.hpp> types are used instead and the shim is a
// drop-in for include/ORBextractor.h, include/LineExtractor.h, include/ORBmatcher.h,
// include/LineMatcher.h of the reference.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/plvi.h"

#ifdef PLVI_HAVE_OPENCV
#include <opencv2/core.hpp>
#else
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
  void create(int r, int c) { rows = r; cols = c; step = (size_t)c; own.assign((size_t)r * c, 0); data = own.data(); }
  void release() { rows = cols = 0; step = 0; data = nullptr; own.clear(); }
  bool empty() const { return !data || rows == 0 || cols == 0; }
  uint8_t* ptr(int r) { return data + (size_t)r * step; }
  const uint8_t* ptr(int r) const { return data + (size_t)r * step; }
  Mat row(int r) const { return Mat(1, cols, const_cast<uint8_t*>(ptr(r)), step); }
};
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
}  // namespace plvi_shim
