// Drop-in for the reference's include/LineExtractor.h:52-92 (class Lineextractor).
#pragma once
#include "plvi_cv_compat.h"

namespace ORB_SLAM3 {

class Lineextractor {
 public:
  typedef cv::line_descriptor::KeyLine KeyLine;
  Lineextractor(int lsd_nfeatures, int lsd_refine, float lsd_scale, int nlevels, float scale, int extractor,
                int max_width = 1280, int max_height = 1024, int device = 0)
      : nlevels_l(nlevels) {
    plvi_shim::check(plvi_line_create(&h_, lsd_nfeatures, lsd_refine, lsd_scale, nlevels, scale, extractor, max_width,
                                      max_height, 1, device, nullptr), "Lineextractor");
    cap_ = plvi_line_capacity(h_);
    mvScaleFactor_l.resize(nlevels); mvInvScaleFactor_l.resize(nlevels); mvLevelSigma2_l.resize(nlevels); mvInvLevelSigma2_l.resize(nlevels);
    plvi_line_scale_factors(h_, mvScaleFactor_l.data(), mvInvScaleFactor_l.data(), mvLevelSigma2_l.data(), mvInvLevelSigma2_l.data());
  }
  ~Lineextractor() { plvi_line_destroy(h_); }
  Lineextractor(const Lineextractor&) = delete;
  Lineextractor& operator=(const Lineextractor&) = delete;

  // void operator()(const Mat& image, const Mat& mask, vector<KeyLine>&, Mat& descriptors_line, vector<Vector3d>& keylineFunction)
  // clears keylines, APPENDS to keylineFunction, leaves descriptors untouched when no line is found,
  // throws std::runtime_error on a mask of the wrong size (src/LineExtractor.cc:45-117).
  void operator()(const cv::Mat& image, const cv::Mat& mask, std::vector<KeyLine>& keylines, cv::Mat& descriptors_line,
                  std::vector<Eigen::Vector3d>& keylineFunction) {
    if (mask.data != nullptr && (mask.rows != image.rows || mask.cols != image.cols))
      throw std::runtime_error("Mask error while detecting lines: please check its dimensions and that data type is CV_8UC1");
    keylines.clear();
    kl_.resize(cap_); desc_.resize((size_t)cap_ * 32); eq_.resize((size_t)cap_ * 3);
    int count = 0;
    plvi_shim::check(plvi_line_extract_batch(h_, image.data, 1, image.cols, image.rows, (int)image.step, image.step * image.rows,
                                             kl_.data(), desc_.data(), eq_.data(), &count), "Lineextractor::operator()");
    plvi_shim::check(count, "Lineextractor::operator() (segment table overflow)");
    if (count == 0) return;   // "Error: keypoint list is empty": descriptors untouched
    keylines.resize(count);
    std::memcpy(static_cast<void*>(keylines.data()), kl_.data(), (size_t)count * sizeof(plvi_keyline));
    descriptors_line.create(count, 32);
    for (int i = 0; i < count; i++) std::memcpy(descriptors_line.ptr(i), desc_.data() + (size_t)i * 32, 32);
    for (int i = 0; i < count; i++) {
      Eigen::Vector3d l;
      l(0) = eq_[3 * i]; l(1) = eq_[3 * i + 1]; l(2) = eq_[3 * i + 2];
      keylineFunction.push_back(l);
    }
  }

  // public members read by Frame (src/Frame.cc:569-574).  The reference push_back()s into
  // these on every call without clearing (a leak); here they are filled once.
  std::vector<float> mvScaleFactor_l, mvInvScaleFactor_l, mvLevelSigma2_l, mvInvLevelSigma2_l;
  int nlevels_l;

 protected:
  plvi_line* h_ = nullptr;
  int cap_ = 0;
  std::vector<plvi_keyline> kl_;
  std::vector<uint8_t> desc_;
  std::vector<double> eq_;
};

}  // namespace ORB_SLAM3
