"""GPU parity of the Frame steps after extraction (device-resident extractor outputs in, through
the C ABI): UndistortKeyPoints / UndistortKeyLines and AssignFeaturesToGrid against the oracle.
f64 arithmetic with float results: required to be bit-exact (tolerance 0)."""
import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import Lineextractor, ORBextractor, synth
from pl_vi_orbslam3_b200 import frame as fr
from pl_vi_orbslam3_b200.matchers import frame_grid

pytestmark = pytest.mark.gpu

CAM = oracle.EUROC_CAMERA


def _cam(c=CAM):
    new_k = None if "new_fx" not in c else (c["new_fx"], c["new_fy"], c["new_cx"], c["new_cy"])
    return fr.make_camera(c["fx"], c["fy"], c["cx"], c["cy"], c["dist"], new_k)


def test_undistort_keypoints_and_grid_after_extraction(gpu):
    import torch
    frames = np.stack([synth.frame_euroc(s) for s in (0, 1, 2)])
    e = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=3)
    try:
        d = torch.from_numpy(frames).cuda()
        kps, desc, counts, _ = e.extract_batch_device(d)
        un = fr.undistort_keypoints(kps, counts, _cam())
        # image bounds as Frame::ComputeImageBounds: undistorted corners
        corners = oracle.undistort_points(np.array([[0, 0], [752, 0], [0, 480], [752, 480]], np.float32))
        grid = frame_grid(min(corners[0, 0], corners[2, 0]), max(corners[1, 0], corners[3, 0]),
                          min(corners[0, 1], corners[1, 1]), max(corners[2, 1], corners[3, 1]))
        start, items = fr.assign_features_to_grid(un, counts, grid)
        torch.cuda.synchronize()
        k_in, k_un, cnt = kps.cpu().numpy(), un.cpu().numpy(), counts.cpu().numpy()
        start, items = start.cpu().numpy(), items.cpu().numpy()
        for f in range(3):
            n = cnt[f]
            ref = oracle.undistort_points(k_in[f, :n, :2])
            assert np.array_equal(k_un[f, :n, :2], ref)
            assert np.array_equal(k_un[f, :n, 2:].view(np.uint32), k_in[f, :n, 2:].view(np.uint32))   # other fields copied
            assert np.abs(ref - k_in[f, :n, :2]).max() > 1.0                                           # the model does move points
            rs, ri = oracle.assign_grid(ref, grid)
            assert np.array_equal(start[f], rs) and np.array_equal(items[f, :rs[-1]], ri)
            if oracle.ref_available():   # the reference's own Frame::UndistortKeyPoints / AssignFeaturesToGrid (Frame.cc compiled unmodified)
                keys = np.zeros(n, oracle.KEYPOINT_DTYPE)
                keys.view(np.float32).reshape(n, 7)[:] = k_in[f, :n].view(np.float32).reshape(n, 7)
                ru = oracle.ref_undistort_keypoints(keys)
                assert np.array_equal(k_un[f, :n, 0], ru["x"]) and np.array_equal(k_un[f, :n, 1], ru["y"])
                g = np.asarray(grid).reshape(-1)[0]
                bounds = (float(g["min_x"]), float(max(corners[1, 0], corners[3, 0])), float(g["min_y"]), float(max(corners[2, 1], corners[3, 1])))
                fs, fi = oracle.ref_assign_grid(ru, bounds)
                assert np.array_equal(start[f], fs) and np.array_equal(items[f, :fs[-1]], fi)
    finally:
        e.close()


def test_undistort_points_golden_and_variants(gpu):
    import torch
    from pathlib import Path
    gold = np.load(Path(__file__).parent / "golden" / "undistort_euroc.npz")
    pts = gold["points"]
    rec = np.zeros((1, len(pts), 7), np.float32)
    rec[0, :, :2] = pts
    cnt = torch.tensor([len(pts)], dtype=torch.int32).cuda()
    d = torch.from_numpy(rec).cuda()
    un = fr.undistort_keypoints(d, cnt, _cam()).cpu().numpy()
    assert np.array_equal(un[0, :, :2], gold["undistorted"])
    cam5 = dict(CAM, dist=(-0.2834, 0.0739, 0.0002, 1.8e-05, -0.011), new_fx=400.0, new_fy=401.0, new_cx=376.0, new_cy=240.0)
    un5 = fr.undistort_keypoints(d, cnt, _cam(cam5)).cpu().numpy()
    assert np.array_equal(un5[0, :, :2], oracle.undistort_points(pts, cam5))
    cam0 = dict(CAM, dist=(0.0, 0.1, 0.0, 0.0))
    un0 = fr.undistort_keypoints(d, cnt, _cam(cam0)).cpu().numpy()
    assert np.array_equal(un0[0, :, :2], pts)                               # mDistCoef[0] == 0: copy
    # in place
    fr.undistort_keypoints(d, cnt, _cam(), out=d)
    assert np.array_equal(d.cpu().numpy()[0, :, :2], gold["undistorted"])


def test_undistort_keylines(gpu):
    import torch
    frames = np.stack([synth.frame_euroc(s) for s in (4, 5)])
    l = Lineextractor(200, 0, 0.8, 2, 2.0, 0, max_batch=2)
    try:
        d = torch.from_numpy(frames).cuda()
        kl, ldesc, leq, lc = l.extract_batch_device(d)
        un = fr.undistort_keylines(kl, lc, _cam())
        torch.cuda.synchronize()
        a, b, cnt = kl.cpu().numpy(), un.cpu().numpy(), lc.cpu().numpy()
        for f in range(2):
            n = cnt[f]
            assert n > 50
            assert np.array_equal(b[f, :n, 7:9], oracle.undistort_points(a[f, :n, 7:9]))     # startPoint
            assert np.array_equal(b[f, :n, 9:11], oracle.undistort_points(a[f, :n, 9:11]))   # endPoint
            keep = [i for i in range(17) if i not in (7, 8, 9, 10)]
            assert np.array_equal(b[f, :n][:, keep].view(np.uint32), a[f, :n][:, keep].view(np.uint32))
    finally:
        l.close()


def test_frame_steps_reject_bad_arguments(gpu):
    import torch
    from pl_vi_orbslam3_b200.capi import PlviError
    d = torch.zeros((1, 8, 7), dtype=torch.float32).cuda()
    cnt = torch.zeros(1, dtype=torch.int32).cuda()
    with pytest.raises(PlviError):
        fr.undistort_keypoints(d, cnt, fr.make_camera(0.0, 1.0, 0.0, 0.0, (0.1,)))
