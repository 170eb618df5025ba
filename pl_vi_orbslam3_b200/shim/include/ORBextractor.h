// Drop-in for the reference's include/ORBextractor.h:45-110: same class name, constructor, operator(), getters and
// public mvImagePyramid; the work is forwarded to libplvi_cuda.so through include/plvi.h.  A SLAM build puts this
// directory in front of the reference's include/ (INTEGRATION.md section 2); src/Frame.cc, src/Tracking.cc compile
// against it unmodified (tests/test_dropin_gpu.py builds the reference's Frame.cc this way).
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H
#include <cstdlib>
#include <list>
#include <vector>

#include "plvi_cv_compat.h"

namespace ORB_SLAM3 {

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  // ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
  // (include/ORBextractor.h:50-51, src/ORBextractor.cc:408-468).  The extra, defaulted arguments size the device
  // buffers; a larger image re-creates them on the fly.
  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int max_width = 0,
               int max_height = 0, int device = 0)
      : nfeatures_(nfeatures), nlevels_(nlevels), iniThFAST_(iniThFAST), minThFAST_(minThFAST), device_(device),
        scaleFactor_(scaleFactor) {
    if (max_width <= 0) max_width = env_int("PLVI_MAX_WIDTH", 1280);
    if (max_height <= 0) max_height = env_int("PLVI_MAX_HEIGHT", 1024);
    create(max_width, max_height);
    mvImagePyramid.resize(nlevels);
  }
  ~ORBextractor() { plvi_orb_destroy(h_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // int operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
  //                cv::OutputArray descriptors, std::vector<int>& vLappingArea)   (include/ORBextractor.h:58-60)
  // Returns -1 on an empty image, else monoIndex (src/ORBextractor.cc:1068-1150).  mask is ignored, as in the reference;
  // descriptors are released when no keypoint is found (:1091-1092), else create()d as N x 32 CV_8U.
  int operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& keypoints,
                 cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
    if (_image.empty()) return -1;
    cv::Mat image = _image.getMat();
    if (image.type() != CV_8UC1) throw std::runtime_error("ORBextractor: image must be CV_8UC1");   // assert(), :1076
    if (image.cols > maxW_ || image.rows > maxH_) {
      plvi_orb_destroy(h_);
      h_ = nullptr;
      create(image.cols > maxW_ ? image.cols : maxW_, image.rows > maxH_ ? image.rows : maxH_);
    }
    kps_.resize(cap_);
    desc_.resize((size_t)cap_ * 32);
    int count = 0, mono = 0;
    plvi_shim::check(plvi_orb_extract_batch(h_, image.ptr(0), 1, image.cols, image.rows, (int)image.step,
                                            (size_t)image.step * image.rows, vLappingArea[0], vLappingArea[1], kps_.data(),
                                            desc_.data(), &count, &mono),
                     "ORBextractor::operator()");
    lastW_ = image.cols; lastH_ = image.rows;
    keypoints.resize(count);
    if (count) std::memcpy(static_cast<void*>(keypoints.data()), kps_.data(), (size_t)count * sizeof(plvi_keypoint));
    if (count == 0) {
      _descriptors.getMatRef().release();
    } else {
      cv::Mat& descriptors = _descriptors.getMatRef();
      descriptors.create(count, 32, CV_8UC1);
      for (int i = 0; i < count; i++) std::memcpy(descriptors.ptr(i), desc_.data() + (size_t)i * 32, 32);
    }
    if (pyramidReadback_) FetchPyramid();
    return mono;
  }

  int inline GetLevels() { return nlevels_; }
  float inline GetScaleFactor() { return scaleFactor_; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  // include/ORBextractor.h:84.  The reference leaves the level images of the last frame here; only the stereo paths
  // read them (src/Frame.cc:1235,1325-1344: inside the image, never the EDGE_THRESHOLD border), so every level is a
  // dense w_l x h_l CV_8U image.  Filled after every operator() (one 1.1 MB device-to-host copy for 752x480, ~50 us);
  // a monocular build that never reads it can switch the copy off.
  std::vector<cv::Mat> mvImagePyramid;
  void SetPyramidReadback(bool on) { pyramidReadback_ = on; }
  void FetchPyramid() {
    std::vector<int> lw(nlevels_), lh(nlevels_);
    plvi_orb_level_sizes(h_, lastW_, lastH_, lw.data(), lh.data());
    for (int l = 0; l < nlevels_; l++) {
      mvImagePyramid[l].create(lh[l], lw[l], CV_8UC1);
      plvi_shim::check(plvi_orb_read_level(h_, 0, l, 0, mvImagePyramid[l].ptr(0)), "ORBextractor::mvImagePyramid");
    }
  }

  plvi_orb* handle() const { return h_; }

  // void Frame::ComputeStereoMatches() (src/Frame.cc:1228-1406) for the frame whose left / right images this
  // extractor and `right` have just processed, with both pyramids read on the device (an OPTIONAL faster path: the
  // unmodified Frame.cc keeps working through mvImagePyramid).  mvuRight / mvDepth are resized to keysLeft.size().
  int ComputeStereoMatches(ORBextractor& right, const std::vector<cv::KeyPoint>& keysLeft, const cv::Mat& descLeft,
                           const std::vector<cv::KeyPoint>& keysRight, const cv::Mat& descRight, float mb, float mbf,
                           std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
    mvuRight.assign(keysLeft.size(), -1.0f);
    mvDepth.assign(keysLeft.size(), -1.0f);
    int n = 0;
    std::vector<uint8_t> tl, tr;
    plvi_shim::check(plvi_orb_stereo_matches_host(h_, right.h_, reinterpret_cast<const plvi_keypoint*>(keysLeft.data()),
                                                  plvi_shim::packed_rows(descLeft, (int)keysLeft.size(), tl), (int)keysLeft.size(),
                                                  reinterpret_cast<const plvi_keypoint*>(keysRight.data()),
                                                  plvi_shim::packed_rows(descRight, (int)keysRight.size(), tr), (int)keysRight.size(),
                                                  mb, mbf, mvuRight.data(), mvDepth.data(), &n),
                     "ComputeStereoMatches");
    return n;
  }

 protected:
  static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && std::atoi(v) > 0) ? std::atoi(v) : dflt;
  }
  void create(int maxW, int maxH) {
    plvi_shim::check(plvi_orb_create(&h_, nfeatures_, scaleFactor_, nlevels_, iniThFAST_, minThFAST_, maxW, maxH, 1, device_,
                                     nullptr), "ORBextractor");
    maxW_ = maxW; maxH_ = maxH;
    cap_ = plvi_orb_capacity(h_);
    mvScaleFactor.resize(nlevels_); mvInvScaleFactor.resize(nlevels_); mvLevelSigma2.resize(nlevels_); mvInvLevelSigma2.resize(nlevels_);
    plvi_orb_scale_factors(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data());
  }

  plvi_orb* h_ = nullptr;
  int nfeatures_, nlevels_, iniThFAST_, minThFAST_, device_;
  int cap_ = 0, lastW_ = 0, lastH_ = 0, maxW_ = 0, maxH_ = 0;
  float scaleFactor_;
  bool pyramidReadback_ = true;
  std::vector<plvi_keypoint> kps_;
  std::vector<uint8_t> desc_;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

}  // namespace ORB_SLAM3

#endif  // ORBEXTRACTOR_H
