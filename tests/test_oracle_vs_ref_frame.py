"""CPU: the oracle's Frame steps against THE REFERENCE'S OWN src/Frame.cc + include/Frame.h.

oracle/_ref/libplvi_ref_frame.so = Frame.cc compiled unmodified, with its own class definition, over stand-in MapPoint /
KeyFrame / MapLine / camera / IMU / vocabulary types (oracle/cvmini/slam_mock_frame.h) and the reference's own
ORBextractor.cc / ORBmatcher.cc / gridStructure.cpp.  Frames are default-constructed and their public members filled in.
Called: AssignFeaturesToGrid (+ PosInGrid), GetFeaturesInArea, lineDescriptorMAD, UndistortKeyPoints, UndistortKeyLines,
ComputeStereoMatches; from KeyFrame.cc + KeyFrame.h (same library): GetFeaturesInArea, GetLinesInArea, lineDescriptorMAD on a
keyframe built by the reference's own KeyFrame(Frame&, ...) constructor.  cv::undistortPoints underneath is the oracle's restatement of OpenCV's (pinned against cv2).

Bar: bit-exact.  Committed outputs (tests/golden/ref_outputs.npz: frame/*) run everywhere, the live tests where the
library exists.
"""
from pathlib import Path

import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import synth
from pl_vi_orbslam3_b200.matchers import frame_grid

GOLD = Path(__file__).resolve().parent / "golden"
R = np.load(GOLD / "ref_outputs.npz")
needs_ref = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
MBF = 47.906
MB = float(np.float32(47.906) / np.float32(435.2))


def undistorted_bounds(w, h):
    """Frame::ComputeImageBounds (src/Frame.cc:1199-1226) for the EuRoC camera: the undistorted image corners."""
    c = oracle.undistort_points(np.array([[0, 0], [w, 0], [0, h], [w, h]], np.float32))
    return (min(c[0, 0], c[2, 0]), max(c[1, 0], c[3, 0]), min(c[0, 1], c[1, 1]), max(c[2, 1], c[3, 1]))


def area_queries(keys, seed, nq=400):
    rng = np.random.RandomState(seed)
    src = rng.randint(0, len(keys), nq)
    xyr = np.stack([keys["x"][src] + rng.uniform(-20, 20, nq), keys["y"][src] + rng.uniform(-20, 20, nq),
                    rng.choice([3.0, 7.2, 15.0, 40.0, 100.0], nq)], 1).astype(np.float32)
    xyr[::25, 0] += 800          # far outside the image
    xyr[1::25, 1] -= 500
    lv = np.stack([rng.randint(-1, 7, nq), rng.randint(-1, 8, nq)], 1).astype(np.int32)
    return xyr, lv


def stereo_pair(seed, d=12):
    left = synth.frame_euroc(seed)
    right = np.empty_like(left)
    right[:, :-d] = left[:, d:]
    right[:, -d:] = left[:, -1:]
    rng = np.random.RandomState(seed)
    right = np.clip(right.astype(np.int32) + rng.randint(-3, 4, right.shape), 0, 255).astype(np.uint8)
    return oracle.orb_extract(left, debug=True), oracle.orb_extract(right, debug=True)


@needs_ref
@pytest.mark.parametrize("seed,bounds", [(0, (0.0, 752.0, 0.0, 480.0)), (1, None), (2, (-31.5, 790.25, -20.0, 505.5))])
def test_live_reference_grid_and_features_in_area(seed, bounds):
    keys = oracle.orb_extract(synth.frame_euroc(seed))["keypoints"]
    if bounds is None:
        bounds = undistorted_bounds(752, 480)
        xy = oracle.undistort_points(np.stack([keys["x"], keys["y"]], 1))
        keys = keys.copy()
        keys["x"], keys["y"] = xy[:, 0], xy[:, 1]
    grid = frame_grid(bounds[0], bounds[1], bounds[2], bounds[3])
    rs, ri = oracle.ref_assign_grid(keys, bounds)
    os_, oi = oracle.assign_grid(np.stack([keys["x"], keys["y"]], 1), grid)
    assert np.array_equal(rs, os_) and np.array_equal(ri, oi)
    xyr, lv = area_queries(keys, seed)
    a = oracle.ref_features_in_area(keys, bounds, xyr, lv)
    b = oracle.features_in_area(keys, grid, xyr, lv)
    assert sum(len(x) for x in a) > 2000
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@needs_ref
def test_live_reference_line_descriptor_mad():
    rng = np.random.RandomState(0)
    for n in (1, 2, 3, 10, 57, 200):
        for spread in (3, 40):
            d0 = rng.randint(0, spread, n).astype(np.int32)
            d1 = d0 + rng.randint(0, spread, n).astype(np.int32)
            assert oracle.ref_line_descriptor_mad(d0, d1) == oracle.line_descriptor_mad(d0, d1)


@needs_ref
@pytest.mark.parametrize("seed", [0, 3])
def test_live_reference_undistort(seed):
    img = synth.frame_euroc(seed)
    keys = oracle.orb_extract(img)["keypoints"]
    r = oracle.ref_undistort_keypoints(keys)
    xy = oracle.undistort_points(np.stack([keys["x"], keys["y"]], 1))
    assert np.array_equal(r["x"], xy[:, 0]) and np.array_equal(r["y"], xy[:, 1])
    for f in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(r[f], keys[f])
    kl = oracle.line_extract(img)["keylines"]
    rl = oracle.ref_undistort_keylines(kl)
    s = oracle.undistort_points(np.stack([kl["startPointX"], kl["startPointY"]], 1))
    e = oracle.undistort_points(np.stack([kl["endPointX"], kl["endPointY"]], 1))
    assert np.array_equal(rl["startPointX"], s[:, 0]) and np.array_equal(rl["startPointY"], s[:, 1])
    assert np.array_equal(rl["endPointX"], e[:, 0]) and np.array_equal(rl["endPointY"], e[:, 1])
    # no distortion: the reference copies the keys
    cam0 = dict(oracle.EUROC_CAMERA, dist=(0.0, 0.0, 0.0, 0.0))
    assert oracle.ref_undistort_keypoints(keys, cam0).tobytes() == keys.tobytes()


@needs_ref
@pytest.mark.parametrize("seed,d", [(5, 12), (6, 30), (7, 3)])
def test_live_reference_compute_stereo_matches(seed, d):
    a, b = stereo_pair(seed, d)
    args = (a["keypoints"], a["descriptors"], b["keypoints"], b["descriptors"], a["pyramid"], b["pyramid"], a["plan"]["scale"], MB, MBF)
    rur, rdp, rn = oracle.ref_stereo_matches(*args)
    our, odp, on = oracle.stereo_matches(*args)
    assert rn == on and np.array_equal(rur, our) and np.array_equal(rdp, odp)
    assert rn > 0.3 * len(rur)


# ---- KeyFrame.cc + KeyFrame.h of the reference, same library: the keyframe is built by the reference's own
# KeyFrame(Frame&, Map*, KeyFrameDatabase*) from a frame filled as above
@needs_ref
@pytest.mark.parametrize("seed", [0, 1])
def test_live_reference_keyframe_features_in_area(seed):
    keys = oracle.orb_extract(synth.frame_euroc(seed))["keypoints"]
    bounds = (0.0, 752.0, 0.0, 480.0)
    xyr, _ = area_queries(keys, seed + 10)
    a = oracle.ref_keyframe_features_in_area(keys, bounds, xyr)
    b = oracle.features_in_area(keys, frame_grid(*bounds), xyr, np.full((len(xyr), 2), -1, np.int32))
    assert sum(len(x) for x in a) > 5000
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_live_reference_keyframe_lines_in_area(seed):
    """GetLinesInArea as written in the reference: midpoint distance in mixed float / double, then `slope - angle` with
    its division by zero for vertical projections."""
    from test_oracle_vs_ref import line_fuse_case
    kl, _, _, q, _, _ = line_fuse_case(seed, [3.0, 20.0, 60.0][seed])
    q5 = q[:, :5]
    a = oracle.ref_keyframe_lines_in_area(kl, (0.0, 752.0, 0.0, 480.0), q5)
    b = oracle.lines_in_area(kl, q5)
    assert sum(len(x) for x in a) > 100
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@needs_ref
def test_live_reference_keyframe_line_descriptor_mad():
    rng = np.random.RandomState(1)
    for n in (1, 2, 9, 64, 200):
        d0 = rng.randint(0, 30, n).astype(np.int32)
        d1 = d0 + rng.randint(0, 30, n).astype(np.int32)
        assert oracle.ref_keyframe_line_descriptor_mad(d0, d1) == oracle.line_descriptor_mad(d0, d1) == oracle.ref_line_descriptor_mad(d0, d1)


# ---- Pinhole.cpp + Pinhole.h + GeometricCamera.h of the reference (library of its own)
@needs_ref
@pytest.mark.parametrize("t12", [(1.0, 0.0, 0.0), (0.37, -0.05, 0.02), (-2.5, 0.4, 1.25), (0.0, 0.0, 0.0)])
def test_live_reference_epipolar_constrain(t12):
    """Pinhole::epipolarConstrain with unit intrinsics (F12 = [t12]x exactly) against the oracle's point-to-line test
    for that F12, on keypoint pairs around and off the epipolar lines, with the level sigmas of the pyramid."""
    rng = np.random.RandomState(3)
    a = oracle.orb_extract(synth.frame_euroc(9))["keypoints"][:600].copy()
    b = a.copy()
    b["x"] += rng.normal(0, 6, len(b)).astype(np.float32)
    b["y"] += rng.normal(0, 1.5, len(b)).astype(np.float32)
    unc = (np.float32(1.2) ** (2 * b["octave"])).astype(np.float32)
    tx, ty, tz = (np.float32(v) for v in t12)
    F12 = np.array([[0, -tz, ty], [tz, 0, -tx], [-ty, tx, 0]], np.float32)
    r = oracle.ref_epipolar_constrain(a, b, t12, unc)
    o = oracle.epipolar_constrain(a, b, F12, unc)
    assert np.array_equal(r, o)
    if t12[0] == 1.0:
        assert 50 < r.sum() < len(r) - 50
    if not any(t12):
        assert not r.any()          # den == 0 -> false


@needs_ref
def test_live_reference_pinhole_project_equals_stand_in_camera():
    """Pinhole::project / toK: the two one-liners the stand-in cameras of the matcher libraries restate."""
    rng = np.random.RandomState(0)
    K = (458.654, 457.296, 367.215, 248.375)
    xyz = rng.uniform(-3, 3, (1000, 3)).astype(np.float32)
    xyz[:, 2] = np.abs(xyz[:, 2]) + np.float32(0.1)
    uv, Km = oracle.ref_pinhole_project(K, xyz)
    Kf = np.array(K, np.float32)
    assert np.array_equal(uv[:, 0], Kf[0] * xyz[:, 0] / xyz[:, 2] + Kf[2]) and np.array_equal(uv[:, 1], Kf[1] * xyz[:, 1] / xyz[:, 2] + Kf[3])
    assert np.array_equal(Km, np.array([[Kf[0], 0, Kf[2]], [0, Kf[1], Kf[3]], [0, 0, 1]], np.float32))


def test_oracle_equals_reference_frame_outputs():
    """Committed outputs of the reference's Frame.cc (tools/gen_golden_ref.py)."""
    a, b = stereo_pair(5, 12)
    ur, dp, n = oracle.stereo_matches(a["keypoints"], a["descriptors"], b["keypoints"], b["descriptors"], a["pyramid"], b["pyramid"],
                                      a["plan"]["scale"], MB, MBF)
    assert n == int(R["frame/stereo_n"]) and np.array_equal(ur, R["frame/stereo_ur"]) and np.array_equal(dp, R["frame/stereo_depth"])
    keys = oracle.orb_extract(synth.frame_euroc(0))["keypoints"]
    grid = frame_grid(0, 752, 0, 480)
    xyr, lv = area_queries(keys, 0)
    lists = oracle.features_in_area(keys, grid, xyr, lv)
    assert np.array_equal(np.cumsum([0] + [len(x) for x in lists]), R["frame/area_start"])
    assert np.array_equal(np.concatenate(lists), R["frame/area_items"])
    gs, gi = oracle.assign_grid(np.stack([keys["x"], keys["y"]], 1), grid)
    assert np.array_equal(gs, R["frame/grid_start"]) and np.array_equal(gi, R["frame/grid_items"])
    assert np.array_equal(oracle.undistort_points(np.stack([keys["x"], keys["y"]], 1)), R["frame/undistorted_xy"])
