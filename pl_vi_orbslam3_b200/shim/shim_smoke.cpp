// Compiles the shim headers against libplvi_cuda.so and, on a GPU box, runs one frame
// through the reference-shaped C++ API:  g++ shim_smoke.cpp -L.. -lplvi_cuda
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "LineExtractor.h"
#include "LineMatcher.h"
#include "ORBextractor.h"
#include "ORBmatcher.h"

int main(int argc, char** argv) {
  const int w = 752, h = 480;
  std::vector<uint8_t> img((size_t)w * h);
  unsigned s = 12345;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      s = s * 1664525u + 1013904223u;
      const int block = (((x / 24) * 7 + (y / 20) * 13) % 11) * 20 + 20;
      img[(size_t)y * w + x] = (uint8_t)(block + (s >> 29));
    }
  cv::Mat im(h, w, img.data()), mask;
  try {
    ORB_SLAM3::ORBextractor orb(1000, 1.2f, 8, 20, 7, w, h);
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc;
    std::vector<int> lap = {0, 0};
    const int mono = orb(im, mask, kps, desc, lap);
    ORB_SLAM3::Lineextractor line(200, 0, 0.8f, 2, 2.0f, 0, w, h);
    std::vector<cv::line_descriptor::KeyLine> kls;
    cv::Mat ldesc;
    std::vector<Eigen::Vector3d> eq;
    line(im, mask, kls, ldesc, eq);
    std::vector<int> m12;
    const int nm = ORB_SLAM3::LineMatcher::match(ldesc, ldesc, 0.9f, m12);
    const int d0 = kps.size() > 1 ? ORB_SLAM3::ORBmatcher::DescriptorDistance(desc.row(0), desc.row(1)) : -1;
    // stereo: a right image = the left one shifted by 10 px; Frame::ComputeStereoMatches through the two extractors
    std::vector<uint8_t> imgR((size_t)w * h);
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) imgR[(size_t)y * w + x] = img[(size_t)y * w + std::min(x + 10, w - 1)];
    cv::Mat imR(h, w, imgR.data());
    ORB_SLAM3::ORBextractor orbR(1000, 1.2f, 8, 20, 7, w, h);
    std::vector<cv::KeyPoint> kpsR;
    cv::Mat descR;
    orbR(imR, mask, kpsR, descR, lap);
    std::vector<float> uRight, depth;
    const int nst = orb.ComputeStereoMatches(orbR, kps, desc, kpsR, descR, 47.9f / 435.2f, 47.9f, uRight, depth);
    int near10 = 0;
    for (size_t i = 0; i < kps.size(); i++)
      if (uRight[i] >= 0 && std::fabs((kps[i].pt.x - uRight[i]) - 10.f) < 0.5f) near10++;
    // SearchByProjection(Frame, Frame) of the frame onto itself (identity pose: every key projects onto its own position)
    // with the first 10 keys blocked: every other key finds itself at distance 0, or a twin with the same descriptor
    ORB_SLAM3::ORBmatcher matcher(0.9f, true);
    std::vector<plvi_query> qs(kps.size());
    for (size_t i = 0; i < kps.size(); i++) {
      qs[i].u = kps[i].pt.x; qs[i].v = kps[i].pt.y;
      qs[i].radius = 15.0f * std::pow(1.2f, (float)kps[i].octave);
      qs[i].min_level = kps[i].octave - 1; qs[i].max_level = kps[i].octave + 1;
      qs[i].angle = kps[i].angle; qs[i].flags = 0;
    }
    plvi_grid grid = {0.f, 0.f, 64.0f / (float)w, 48.0f / (float)h};
    std::vector<uint8_t> blocked(kps.size(), 0);
    for (size_t i = 0; i < blocked.size() && i < 10; i++) blocked[i] = 1;
    std::vector<int> ofKey;
    const int nproj = matcher.SearchByProjectionRaw(kps, desc, grid, qs, desc, ofKey, PLVI_SEARCH_FRAME, &blocked);
    int blockedHit = 0;
    for (size_t i = 0; i < ofKey.size() && i < 10; i++) blockedHit += ofKey[i] >= 0;
    if (blockedHit != 0 || nproj < (int)kps.size() / 2) { std::printf("shim error: SearchByProjection %d matches, %d on blocked keys\n", nproj, blockedHit); return 1; }
    // stereo lines: the right image's keylines through LineMatcher::matchGrid (Frame::ComputeStereoMatches_Lines' search)
    ORB_SLAM3::Lineextractor lineR(200, 0, 0.8f, 2, 2.0f, 0, w, h);
    std::vector<cv::line_descriptor::KeyLine> klsR;
    cv::Mat ldescR;
    std::vector<Eigen::Vector3d> eqR;
    lineR(imR, mask, klsR, ldescR, eqR);
    std::vector<int> mgrid;
    const int ngrid = ORB_SLAM3::LineMatcher::matchStereoLines(kls, ldesc, klsR, ldescR, 64.0 / w, 48.0 / h, mgrid);
    int gridCount = 0;
    for (int v : mgrid) gridCount += v >= 0;
    if (ngrid != gridCount || (int)mgrid.size() != (int)kls.size()) { std::printf("shim error: matchGrid %d vs %d\n", ngrid, gridCount); return 1; }
    std::printf("shim: matchGrid %d stereo line matches of %zu / %zu lines\n", ngrid, kls.size(), klsR.size());
    std::printf("shim ok: monoIndex=%d keypoints=%zu lines=%zu self-matches=%d dist01=%d stereo=%d (disparity 10: %d)\n", mono, kps.size(),
                kls.size(), nm, d0, nst, near10);
    return (mono == (int)kps.size() && nm <= (int)kls.size() && nst > 0 && near10 * 4 > nst) ? 0 : 1;   // the block pattern repeats, so part of the matches sit on another period
  } catch (const std::exception& e) {
    std::printf("shim error: %s\n", e.what());
    return argc > 1 ? 0 : 2;   // with an argument: tolerate "no CUDA device" (CPU box link check)
  }
}
