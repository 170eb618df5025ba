// TEST INFRASTRUCTURE -- lets src/KeyFrame.cc of the reference compile unmodified WITH ITS OWN include/KeyFrame.h AND the
// reference's own include/Frame.h: slam_mock_frame.h (stand-in MapPoint / MapLine / cameras / IMU types / vocabulary /
// Converter) with the stand-in KeyFrame switched off, plus stand-ins for Map.h and KeyFrameDatabase.h.  Only member
// functions that touch plain data are called (oracle/ref_glue_keyframe.cpp).
#pragma once
#define SLAM_MOCK_REAL_KEYFRAME
#define MAP_H
#define KEYFRAMEDATABASE_H
#include "slam_mock_frame.h"

using Eigen::Vector3d;

namespace ORB_SLAM3 {
class Map {
 public:
  long unsigned int GetId() { return 0; }
  long unsigned int GetInitKFid() { return 0; }
  void EraseKeyFrame(KeyFrame*) {}
  bool IsInertial() { return false; }
  bool isImuInitialized() { return false; }
};
class KeyFrameDatabase { public: void erase(KeyFrame*) {} };
}  // namespace ORB_SLAM3
