"""CPU: the oracle's restatement of cv::undistortPoints (as called by Frame::UndistortKeyPoints,
src/Frame.cc:1124-1159) against cv2 and the committed golden vectors, and the grid assignment."""
import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import synth
from pl_vi_orbslam3_b200.matchers import frame_grid

GOLD = np.load(__import__("pathlib").Path(__file__).parent / "golden" / "undistort_euroc.npz")


def test_undistort_matches_golden():
    und = oracle.undistort_points(GOLD["points"])
    assert np.array_equal(und, GOLD["undistorted"])


def test_undistort_matches_cv2_if_present():
    cv2 = pytest.importorskip("cv2")
    cam = oracle.EUROC_CAMERA
    K = np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], np.float32)
    D = np.array(cam["dist"], np.float32)
    rng = np.random.RandomState(0)
    pts = (rng.rand(5000, 2) * [752, 480]).astype(np.float32)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    assert np.array_equal(oracle.undistort_points(pts), ref)
    # five-coefficient model (k3) and a different new camera matrix
    cam5 = dict(cam, dist=(-0.2834, 0.0739, 0.0002, 1.8e-05, -0.011), new_fx=400.0, new_fy=401.0, new_cx=376.0, new_cy=240.0)
    D5 = np.array(cam5["dist"], np.float32)
    P = np.array([[400.0, 0, 376.0], [0, 401.0, 240.0], [0, 0, 1]])
    ref5 = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D5, None, P).reshape(-1, 2)
    assert np.array_equal(oracle.undistort_points(pts, cam5), ref5)


def test_undistort_zero_distortion_is_a_copy():
    pts = np.array([[10.5, 20.25], [700.0, 400.0]], np.float32)
    cam = dict(oracle.EUROC_CAMERA, dist=(0.0, 0.1, 0.0, 0.0))    # mDistCoef[0] == 0 -> mvKeysUn = mvKeys
    assert np.array_equal(oracle.undistort_points(pts, cam), pts)


def test_assign_grid_orders_and_bounds():
    rng = np.random.RandomState(3)
    xy = (rng.rand(2000, 2) * [760, 490] - [4, 5]).astype(np.float32)     # some points outside the grid
    grid = frame_grid(0, 752, 0, 480)
    start, items = oracle.assign_grid(xy, grid)
    assert start[0] == 0 and (np.diff(start) >= 0).all() and start[-1] == len(items)
    px = np.round((xy[:, 0] - grid["min_x"]) * grid["inv_w"])
    py = np.round((xy[:, 1] - grid["min_y"]) * grid["inv_h"])
    inside = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    assert len(items) == inside.sum()
    for c in rng.randint(0, 64 * 48, 200):
        cell = items[start[c]:start[c + 1]]
        assert (np.diff(cell) > 0).all()                                  # insertion = index order
        assert all(int(px[i]) * 48 + int(py[i]) == c for i in cell)


def test_stereo_oracle_recovers_known_disparity():
    """Frame::ComputeStereoMatches restatement on a synthetic pair with a constant 12 px disparity: most keypoints get
    a stereo match, the sub-pixel disparities sit at 12 px, depth = bf / disparity, SAD outliers are removed."""
    left = synth.frame_euroc(5)
    d = 12
    right = np.empty_like(left)
    right[:, :-d] = left[:, d:]
    right[:, -d:] = left[:, -1:]
    a, b = oracle.orb_extract(left, debug=True), oracle.orb_extract(right, debug=True)
    mbf, mb = 47.906, np.float32(47.906) / np.float32(435.2)
    ur, dp, n = oracle.stereo_matches(a["keypoints"], a["descriptors"], b["keypoints"], b["descriptors"], a["pyramid"],
                                      b["pyramid"], a["plan"]["scale"], mb, mbf)
    ok = ur >= 0
    assert n == ok.sum() and n > 0.4 * len(ur)
    disp = a["keypoints"]["x"][ok] - ur[ok]
    assert abs(np.median(disp) - d) < 0.05 and np.percentile(np.abs(disp - d), 75) < 0.5
    assert np.allclose(dp[ok], np.float32(mbf) / disp, rtol=1e-6)
    assert (dp[~ok] == -1).all()
