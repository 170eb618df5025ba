// TEST INFRASTRUCTURE -- C entry point over the reference's MapPoint (src/MapPoint.cc + include/MapPoint.h compiled
// unmodified; KeyFrame / Frame / Map are the stand-ins of cvmini/slam_mock_orb.h in its SLAM_MOCK_REAL_MAPPOINT mode;
// ORBmatcher.cc is compiled a second time against this class set for DescriptorDistance).  oracle/Makefile.ref builds
// it into a library of its own (oracle/_ref/libplvi_ref_mappoint.so).
#include <cstring>
#include <vector>
#include "MapPoint.h"   // /root/reference/include
#include "MapLine.h"    // /root/reference/include

using namespace ORB_SLAM3;

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:330-402): observation i = keyframe i (the observations map
// is ordered by KeyFrame address: the keyframes live in one array, so that order is the index order), feature 0 of a
// monocular keyframe holding desc[i]; bad[i] != 0: pKF->isBad().  Writes the chosen descriptor; returns 1, or 0 when
// the reference returns early (no usable observation).
extern "C" int plviref_distinctive_descriptor(const unsigned char* desc, const unsigned char* bad, int n, unsigned char* out) {
  Map map;
  std::vector<KeyFrame> kfs(n > 0 ? n : 1);
  for (int i = 0; i < n; i++) {
    kfs[i].mnId = i;
    kfs[i].N = 1;
    kfs[i].mDescriptors = cv::Mat(1, 32, CV_8UC1);
    memcpy(kfs[i].mDescriptors.data, desc + 32 * (size_t)i, 32);
    kfs[i].mvuRight.assign(1, -1.0f);
    kfs[i].mBad = bad && bad[i];
  }
  cv::Mat pos = cv::Mat::zeros(3, 1, CV_32F);
  MapPoint mp(pos, &kfs[0], &map);
  for (int i = 0; i < n; i++) mp.AddObservation(&kfs[i], 0);
  mp.ComputeDistinctiveDescriptors();
  cv::Mat d = mp.GetDescriptor();
  if (d.empty()) return 0;
  memcpy(out, d.data, 32);
  return 1;
}

// MapLine::ComputeDistinctiveDescriptors (src/MapLine.cc:264-329), same construction: observation i = keyframe i holding
// the LBD descriptor desc[i] as row 0 of mDescriptors_l (a monocular line: mvDepth_l = (-1, -1)).  The distance it uses
// is ORBmatcher::DescriptorDistance (src/MapLine.cc:305).
extern "C" int plviref_mapline_distinctive_descriptor(const unsigned char* desc, const unsigned char* bad, int n, unsigned char* out) {
  Map map;
  std::vector<KeyFrame> kfs(n > 0 ? n : 1);
  for (int i = 0; i < n; i++) {
    kfs[i].mnId = i;
    kfs[i].mDescriptors_l = cv::Mat(1, 32, CV_8UC1);
    memcpy(kfs[i].mDescriptors_l.data, desc + 32 * (size_t)i, 32);
    kfs[i].mvDepth_l.assign(1, std::make_pair(-1.0f, -1.0f));
    kfs[i].mBad = bad && bad[i];
  }
  MapLine ml(Eigen::Vector3d(0, 0, 1), Eigen::Vector3d(1, 0, 1), &kfs[0], &map);
  for (int i = 0; i < n; i++) ml.AddObservation(&kfs[i], 0);
  ml.ComputeDistinctiveDescriptors();
  cv::Mat d = ml.GetDescriptor();
  if (d.empty()) return 0;
  memcpy(out, d.data, 32);
  return 1;
}
