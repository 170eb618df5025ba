// TEST INFRASTRUCTURE -- C entry point over the reference's Pinhole camera (src/CameraModels/Pinhole.cpp + Pinhole.h +
// GeometricCamera.h compiled unmodified; TwoViewReconstruction is a stand-in, cvmini/slam_mock_pinhole.h).
// oracle/Makefile.ref builds it into oracle/_ref/libplvi_ref_pinhole.so.
#include <vector>
#include "Pinhole.h"   // /root/reference/include/CameraModels

using namespace ORB_SLAM3;

// Pinhole::epipolarConstrain(pCamera2, kp1, kp2, R12 = I, t12, sigmaLevel, unc) (src/CameraModels/Pinhole.cpp:135-157) for n
// keypoint pairs, both cameras with intrinsics K = (fx, fy, cx, cy).  With unit intrinsics F12 = [t12]x exactly; with other
// intrinsics F12 carries the rounding of the stand-in's matrix inverse and product (not OpenCV's).
extern "C" void plviref_pinhole_epipolar_constrain(const cv::KeyPoint* kp1, const cv::KeyPoint* kp2, int n, const float* K,
                                                   const float* t12, const float* unc, unsigned char* ok) {
  std::vector<float> p(K, K + 4);
  Pinhole c1(p), c2(p);
  cv::Mat R = cv::Mat::eye(3, 3, CV_32F), t(3, 1, CV_32F);
  for (int i = 0; i < 3; i++) t.at<float>(i) = t12[i];
  for (int i = 0; i < n; i++) ok[i] = c1.epipolarConstrain(&c2, kp1[i], kp2[i], R, t, 1.0f, unc[i]) ? 1 : 0;
}

// Pinhole::project(cv::Point3f) and toK (src/CameraModels/Pinhole.cpp:27-39, 129-133): the stand-in cameras of
// slam_mock_orb.h restate these two one-liners.
extern "C" void plviref_pinhole_project(const float* K, const float* xyz, int n, float* uv, float* Kout) {
  std::vector<float> p(K, K + 4);
  Pinhole c(p);
  for (int i = 0; i < n; i++) {
    const cv::Point2f q = c.project(cv::Point3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    uv[2 * i] = q.x; uv[2 * i + 1] = q.y;
  }
  const cv::Mat Km = c.toK();
  for (int i = 0; i < 9; i++) Kout[i] = Km.at<float>(i / 3, i % 3);
}
