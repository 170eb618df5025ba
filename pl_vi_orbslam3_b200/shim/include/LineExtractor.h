// Drop-in for the reference's include/LineExtractor.h:52-92 (class Lineextractor): same constructor, operator() and
// public members; LSD + LBD run in libplvi_cuda.so.
#ifndef LINEEXTRACTOR_H
#define LINEEXTRACTOR_H
#include <cstdlib>
#include <list>
#include <vector>

#include "plvi_cv_compat.h"

#ifdef PLVI_HAVE_OPENCV
// the reference header exports these two namespaces to every file that includes it (include/LineExtractor.h:31-32);
// Frame.cc / Tracking.cc name KeyLine, Mat, DMatch unqualified and rely on it
using namespace cv;
using namespace line_descriptor;
#endif

namespace ORB_SLAM3 {

// comparison functors the reference header declares next to the class (include/LineExtractor.h:37-51)
struct sort_lines_by_response {
  inline bool operator()(const cv::line_descriptor::KeyLine& a, const cv::line_descriptor::KeyLine& b) { return a.response > b.response; }
};
struct sort_lines_by_length {
  inline bool operator()(const cv::line_descriptor::KeyLine& a, const cv::line_descriptor::KeyLine& b) { return a.lineLength > b.lineLength; }
};

class Lineextractor {
 public:
  typedef cv::line_descriptor::KeyLine KeyLine;

  // Lineextractor(int lsd_nfeatures, int lsd_refine, float lsd_scale, int nlevels, float scale, int extractor)
  // (include/LineExtractor.h:55, src/LineExtractor.cc:39-43)
  Lineextractor(int lsd_nfeatures, int lsd_refine, float lsd_scale, int nlevels, float scale, int extractor,
                int max_width = 0, int max_height = 0, int device = 0)
      : nlevels_l(nlevels), nfeatures_(lsd_nfeatures), refine_(lsd_refine), extractor_(extractor), device_(device),
        lsdScale_(lsd_scale), scale_(scale) {
    if (max_width <= 0) max_width = env_int("PLVI_MAX_WIDTH", 1280);
    if (max_height <= 0) max_height = env_int("PLVI_MAX_HEIGHT", 1024);
    create(max_width, max_height);
  }
  ~Lineextractor() { plvi_line_destroy(h_); }
  Lineextractor(const Lineextractor&) = delete;
  Lineextractor& operator=(const Lineextractor&) = delete;

  // void operator()(const cv::Mat& image, const cv::Mat& mask, std::vector<KeyLine>& keylines, cv::Mat& descriptors_line,
  //                 std::vector<Eigen::Vector3d>& keylineFunction)   (include/LineExtractor.h:59-61)
  // clears keylines, APPENDS to keylineFunction, leaves descriptors untouched when no line is found
  // (binary_descriptor_custom.cpp:557-561), throws std::runtime_error on a mask of the wrong size
  // (LSDDetector_custom.cpp:256-257,274)  (src/LineExtractor.cc:45-117).
  void operator()(const cv::Mat& image, const cv::Mat& mask, std::vector<KeyLine>& keylines, cv::Mat& descriptors_line,
                  std::vector<Eigen::Vector3d>& keylineFunction) {
    if (mask.data != nullptr && (mask.rows != image.rows || mask.cols != image.cols || mask.type() != CV_8UC1))
      throw std::runtime_error("Mask error while detecting lines: please check its dimensions and that data type is CV_8UC1");
    if (image.type() != CV_8UC1) throw std::runtime_error("Lineextractor: image must be CV_8UC1");
    keylines.clear();
    if (image.cols > maxW_ || image.rows > maxH_) {
      plvi_line_destroy(h_);
      h_ = nullptr;
      create(image.cols > maxW_ ? image.cols : maxW_, image.rows > maxH_ ? image.rows : maxH_);
    }
    kl_.resize(cap_); desc_.resize((size_t)cap_ * 32); eq_.resize((size_t)cap_ * 3);
    int count = 0;
    plvi_shim::check(plvi_line_extract_batch(h_, image.ptr(0), 1, image.cols, image.rows, (int)image.step,
                                             (size_t)image.step * image.rows, kl_.data(), desc_.data(), eq_.data(), &count),
                     "Lineextractor::operator()");
    plvi_shim::check(count, "Lineextractor::operator() (segment table overflow)");
    lastW_ = image.cols; lastH_ = image.rows;
    if (pyramidReadback_) FetchPyramid();
    if (count == 0) return;   // "Error: keypoint list is empty": descriptors untouched
    keylines.resize(count);
    std::memcpy(static_cast<void*>(keylines.data()), kl_.data(), (size_t)count * sizeof(plvi_keyline));
    descriptors_line.create(count, 32, CV_8UC1);
    for (int i = 0; i < count; i++) std::memcpy(descriptors_line.ptr(i), desc_.data() + (size_t)i * 32, 32);
    for (int i = 0; i < count; i++) {
      Eigen::Vector3d l;
      l(0) = eq_[3 * i]; l(1) = eq_[3 * i + 1]; l(2) = eq_[3 * i + 2];
      keylineFunction.push_back(l);
    }
  }

  // Public members read by Frame (src/Frame.cc:569-574).  The reference push_back()s nlevels entries into the scale
  // vectors and mvImagePyramid_l on EVERY call without clearing (src/LineExtractor.cc:90-101: they grow by nlevels per
  // frame, only the first nlevels entries are ever indexed); here they always hold exactly nlevels entries.
  // mvImagePyramid_l (LSDDetectorC::gaussianPyrs) has no reader in the reference, so its device-to-host copy is off
  // unless SetPyramidReadback(true).
  std::vector<cv::Mat> mvImagePyramid_l;
  std::vector<float> mvScaleFactor_l, mvInvScaleFactor_l, mvLevelSigma2_l, mvInvLevelSigma2_l;
  int nlevels_l;
  void SetPyramidReadback(bool on) { pyramidReadback_ = on; }
  void FetchPyramid() {
    std::vector<int> ow(nlevels_l), oh(nlevels_l), sw(nlevels_l), sh(nlevels_l);
    plvi_line_octave_sizes(h_, lastW_, lastH_, ow.data(), oh.data(), sw.data(), sh.data());
    mvImagePyramid_l.resize(nlevels_l);
    for (int o = 0; o < nlevels_l; o++) {
      mvImagePyramid_l[o].create(oh[o], ow[o], CV_8UC1);
      int cnt = 0;
      plvi_shim::check(plvi_line_read_lsd(h_, 0, o, 4, mvImagePyramid_l[o].ptr(0), ow[o] * oh[o], &cnt), "Lineextractor::mvImagePyramid_l");
    }
  }

  plvi_line* handle() const { return h_; }

 protected:
  static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && std::atoi(v) > 0) ? std::atoi(v) : dflt;
  }
  void create(int maxW, int maxH) {
    plvi_shim::check(plvi_line_create(&h_, nfeatures_, refine_, lsdScale_, nlevels_l, scale_, extractor_, maxW, maxH, 1, device_,
                                      nullptr), "Lineextractor");
    maxW_ = maxW; maxH_ = maxH;
    cap_ = plvi_line_capacity(h_);
    mvScaleFactor_l.resize(nlevels_l); mvInvScaleFactor_l.resize(nlevels_l); mvLevelSigma2_l.resize(nlevels_l); mvInvLevelSigma2_l.resize(nlevels_l);
    plvi_line_scale_factors(h_, mvScaleFactor_l.data(), mvInvScaleFactor_l.data(), mvLevelSigma2_l.data(), mvInvLevelSigma2_l.data());
  }

  plvi_line* h_ = nullptr;
  int nfeatures_, refine_, extractor_, device_;
  int cap_ = 0, lastW_ = 0, lastH_ = 0, maxW_ = 0, maxH_ = 0;
  float lsdScale_, scale_;
  bool pyramidReadback_ = false;
  std::vector<plvi_keyline> kl_;
  std::vector<uint8_t> desc_;
  std::vector<double> eq_;
};

}  // namespace ORB_SLAM3

#endif  // LINEEXTRACTOR_H
