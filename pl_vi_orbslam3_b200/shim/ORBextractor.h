// Drop-in for the reference's include/ORBextractor.h:45-110: same class name, constructor,
// operator() and getters; the work is forwarded to libplvi_cuda.so through include/plvi.h.
#pragma once
#include "plvi_cv_compat.h"

namespace ORB_SLAM3 {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
               int max_width = 1280, int max_height = 1024, int device = 0)
      : nlevels_(nlevels), scaleFactor_(scaleFactor) {
    plvi_shim::check(plvi_orb_create(&h_, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_width, max_height, 1,
                                     device, nullptr), "ORBextractor");
    cap_ = plvi_orb_capacity(h_);
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    plvi_orb_scale_factors(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data());
    mvImagePyramid.resize(nlevels);
  }
  ~ORBextractor() { plvi_orb_destroy(h_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // int operator()(InputArray image, InputArray mask, vector<KeyPoint>&, OutputArray descriptors, vector<int>& vLappingArea)
  // Returns -1 on an empty image, else monoIndex (src/ORBextractor.cc:1068-1150).  mask is ignored, as in the reference.
  int operator()(const cv::Mat& image, const cv::Mat& /*mask*/, std::vector<cv::KeyPoint>& keypoints,
                 cv::Mat& descriptors, std::vector<int>& vLappingArea) {
    if (image.empty()) return -1;
    kps_.resize(cap_);
    desc_.resize((size_t)cap_ * 32);
    int count = 0, mono = 0;
    plvi_shim::check(plvi_orb_extract_batch(h_, image.data, 1, image.cols, image.rows, (int)image.step, image.step * image.rows,
                                            vLappingArea[0], vLappingArea[1], kps_.data(), desc_.data(), &count, &mono),
                     "ORBextractor::operator()");
    lastW_ = image.cols; lastH_ = image.rows;
    keypoints.resize(count);
    std::memcpy(static_cast<void*>(keypoints.data()), kps_.data(), (size_t)count * sizeof(plvi_keypoint));
    if (count == 0) descriptors.release();
    else {
      descriptors.create(count, 32);
      for (int i = 0; i < count; i++) std::memcpy(descriptors.ptr(i), desc_.data() + (size_t)i * 32, 32);
    }
    return mono;
  }

  int inline GetLevels() { return nlevels_; }
  float inline GetScaleFactor() { return scaleFactor_; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  plvi_orb* handle() const { return h_; }

  // void Frame::ComputeStereoMatches() (src/Frame.cc:1228-1406) for the frame whose left / right images this
  // extractor and `right` have just processed: the SAD refinement reads both pyramids on the device, so the host
  // copies of mvImagePyramid are not needed.  mvuRight / mvDepth are resized to keysLeft.size() (-1 = no match).
  int ComputeStereoMatches(ORBextractor& right, const std::vector<cv::KeyPoint>& keysLeft, const cv::Mat& descLeft,
                           const std::vector<cv::KeyPoint>& keysRight, const cv::Mat& descRight, float mb, float mbf,
                           std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(plvi_keypoint), "cv::KeyPoint layout");
    mvuRight.assign(keysLeft.size(), -1.0f);
    mvDepth.assign(keysLeft.size(), -1.0f);
    int n = 0;
    plvi_shim::check(plvi_orb_stereo_matches_host(h_, right.h_, reinterpret_cast<const plvi_keypoint*>(keysLeft.data()), descLeft.data,
                                                  (int)keysLeft.size(), reinterpret_cast<const plvi_keypoint*>(keysRight.data()),
                                                  descRight.data, (int)keysRight.size(), mb, mbf, mvuRight.data(), mvDepth.data(), &n),
                     "ComputeStereoMatches");
    return n;
  }

  // The reference exposes mvImagePyramid as a public member that only the stereo paths read
  // (src/Frame.cc:1235,1325-1344).  Here it is filled lazily: call FetchPyramid() after
  // operator() when a consumer needs the host copy.
  std::vector<cv::Mat> mvImagePyramid;
  void FetchPyramid() {
    std::vector<int> lw(nlevels_), lh(nlevels_);
    plvi_orb_level_sizes(h_, lastW_, lastH_, lw.data(), lh.data());
    for (int l = 0; l < nlevels_; l++) {
      mvImagePyramid[l].create(lh[l], lw[l]);
      plvi_shim::check(plvi_orb_read_level(h_, 0, l, 0, mvImagePyramid[l].data), "FetchPyramid");
    }
  }

 protected:
  plvi_orb* h_ = nullptr;
  int nlevels_, cap_ = 0, lastW_ = 0, lastH_ = 0;
  float scaleFactor_;
  std::vector<plvi_keypoint> kps_;
  std::vector<uint8_t> desc_;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

}  // namespace ORB_SLAM3
