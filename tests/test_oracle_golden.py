"""CPU: the oracle against the committed golden vectors (tests/golden/, made by
tools/gen_golden.py).  cv2_* vectors are OpenCV 4.13 outputs on the reference's own
frames and on synthetic frames: they pin the oracle's restatement of the OpenCV primitives
the reference calls.  oracle_* vectors pin the reference-owned logic against regressions; the pin
against the reference itself (its sources compiled unmodified into oracle/_ref) lives in
tests/test_oracle_vs_ref*.py."""
import zlib
from pathlib import Path

import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import synth

GOLD = Path(__file__).resolve().parent / "golden"
G = np.load(GOLD / "golden.npz")
S = float(np.float32(0.8))
SIGMA = 0.6 / S


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def _frame(name):
    if name == "synth_0":
        return synth.frame_euroc(0)
    if name == "synth_7_640":
        return synth.frame_euroc(7, 640, 480)
    return np.load(GOLD / f"frame_{name}.npz")["img"]


NAMES = ["data2_1", "data2_3", "data_1_gray", "synth_0", "synth_7_640"]


@pytest.mark.parametrize("name", NAMES)
def test_pyramid_blur_fast_match_opencv(name):
    img = _frame(name)
    h, w = img.shape
    plan = oracle.orb_plan(w, h)
    cur = img
    for l in range(1, 8):
        cur = oracle.resize_linear(cur, int(plan["w"][l]), int(plan["h"][l]))
        assert crc(cur) == G[f"cv2_pyr_{name}_{l}"], f"resize level {l}"
        if l in (1, 4, 7):
            assert crc(oracle.gaussian_blur7(cur)) == G[f"cv2_blur7_{name}_{l}"]
            c = oracle.grid_fast(cur)
            ref = G[f"cv2_gridfast_{name}_{l}"]
            if ref.ndim == 2:
                assert np.array_equal(c, ref), f"grid FAST level {l}"
            else:
                assert (len(c), crc(c)) == (ref[0], np.uint32(ref[1]))
    assert crc(oracle.gaussian_blur7(img)) == G[f"cv2_blur7_{name}_0"]
    c0 = oracle.grid_fast(img)
    assert (len(c0), crc(c0)) == (G[f"cv2_gridfast_{name}_0"][0], np.uint32(G[f"cv2_gridfast_{name}_0"][1]))
    assert crc(oracle.resize_linear(img, w // 2, h // 2)) == G[f"cv2_half_{name}"]


@pytest.mark.parametrize("name", NAMES)
def test_lbd_and_lsd_preprocessing_match_opencv(name):
    img = _frame(name)
    b5 = oracle.gaussian_blur5(img)
    assert crc(b5) == G[f"cv2_blur5_{name}"]
    pd = oracle.pyr_down(b5)
    assert crc(pd) == G[f"cv2_pyrdown_{name}"]
    assert crc(oracle.sobel3(b5)[0]) == G[f"cv2_sobelx_{name}"]
    assert crc(oracle.sobel3(pd)[1]) == G[f"cv2_sobely_{name}"]
    k = oracle.gaussian_kernel_f64(7, SIGMA)
    assert np.array_equal(k, G["cv2_gauss_kernel7"])
    gb = oracle.gaussian_blur_f64(img.astype(np.float64), k)
    # OpenCV's own f64 blur is only reproducible to ~1e-13 (SIMD/FMA inside the library)
    assert np.abs(gb[::37, ::41] - G[f"cv2_blurf64_sample_{name}"]).max() <= 1e-12
    sh = tuple(G[f"cv2_scaled_shape_{name}"])
    sc = oracle.resize_linear_f64(gb, sh[1], sh[0], S, S)
    assert crc(sc) == G[f"cv2_resizef64_of_oracleblur_{name}"]


def test_fast_atan2_matches_opencv():
    yx = G["cv2_atan2_in"]
    got = np.array([oracle.fast_atan2(a, b) for a, b in yx], np.float32)
    assert np.array_equal(got, G["cv2_atan2_out"])


def test_knn2_tie_rule_matches_bfmatcher():
    d1, d2, ref = G["cv2_knn_d1"], G["cv2_knn_d2"], G["cv2_knn"]
    for i in range(len(d1)):
        dist = np.array([oracle.hamming256(d1[i], d2[j]) for j in range(len(d2))])
        order = np.lexsort((np.arange(len(d2)), dist))
        assert order[0] == int(ref[i, 0]) and dist[order[0]] == ref[i, 1]
        assert dist[order[1]] == ref[i, 3]
    n, m = oracle.match_nnr(d1, d2, 0.9)
    exp = np.where(ref[:, 1] < ref[:, 3] * np.float32(0.9), ref[:, 0], -1).astype(np.int32)
    assert np.array_equal(m, exp)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_orb_and_line_regression(name):
    img = _frame(name)
    r = oracle.orb_extract(img)
    ref = G[f"oracle_orb_{name}"]
    assert (len(r["keypoints"]), r["mono_index"]) == (ref[0], ref[1])
    assert crc(r["keypoints"]) == np.uint32(ref[2]) and crc(r["descriptors"]) == np.uint32(ref[3])
    assert np.array_equal(r["keypoints"][:16], G[f"oracle_orb_head_{name}"])
    lr = oracle.line_extract(img)
    ref = G[f"oracle_line_{name}"]
    assert (len(lr["keylines"]), lr["raw_counts"][0], lr["raw_counts"][1]) == (ref[0], ref[1], ref[2])
    assert crc(lr["keylines"]) == np.uint32(ref[3]) and crc(lr["descriptors"]) == np.uint32(ref[4])


def test_oracle_structural_properties():
    img = synth.frame_euroc(4)
    r = oracle.orb_extract(img, debug=True)
    k = r["keypoints"]
    quota = r["plan"]["quota"]
    assert (np.diff(k["octave"]) >= 0).all()                      # mono fill: level by level
    for l in range(8):
        n = (k["octave"] == l).sum()
        assert n <= max(quota[l] + 3, 8)
    lv = k["octave"]
    x = k["x"] / r["plan"]["scale"][lv]
    y = k["y"] / r["plan"]["scale"][lv]
    assert (x >= 18.99).all() and (y >= 18.99).all()              # EDGE_THRESHOLD
    assert (x <= r["plan"]["w"][lv] - 19.99 + 1e-3).all() and (y <= r["plan"]["h"][lv] - 19.99 + 1e-3).all()
    assert len({(a, b, c) for a, b, c in zip(k["x"], k["y"], k["octave"])}) == len(k)
    # lapping area: {0,1000} reverses the output (all keys are "stereo" candidates)
    r2 = oracle.orb_extract(img, lapping=(0, 1000))
    assert r2["mono_index"] == 0
    assert np.array_equal(r2["keypoints"][::-1], k) and np.array_equal(r2["descriptors"][::-1], r["descriptors"])
    # Hamming distance properties
    d = r["descriptors"]
    assert oracle.hamming256(d[0], d[0]) == 0
    assert oracle.hamming256(d[0], ~d[0]) == 256
    assert oracle.hamming256(d[0], d[1]) == int(np.unpackbits(d[0] ^ d[1]).sum())
    assert oracle.hamming256(d[0], d[1], shift25=True) == sum(
        int(np.unpackbits((d[0] ^ d[1])[4 * i:4 * i + 4]).sum()) // 2 for i in range(8))
    # line match is symmetric under swapping the sets (mutual check)
    la, lb = oracle.line_extract(img), oracle.line_extract(synth.frame_euroc(5))
    n12, m12 = oracle.line_match(la["descriptors"], lb["descriptors"], 0.9)
    n21, m21 = oracle.line_match(lb["descriptors"], la["descriptors"], 0.9)
    assert n12 == n21
    assert all(m21[j] == i for i, j in enumerate(m12) if j >= 0)


def test_oracle_edge_cases():
    flat = np.full((480, 752), 77, np.uint8)
    r = oracle.orb_extract(flat)
    assert len(r["keypoints"]) == 0 and r["mono_index"] == 0
    assert len(oracle.line_extract(flat)["keylines"]) == 0
    n, m = oracle.line_match(np.zeros((0, 32), np.uint8), np.zeros((5, 32), np.uint8), 0.9)
    assert n == 0 and len(m) == 0
    n, m = oracle.line_match(np.zeros((3, 32), np.uint8), np.zeros((1, 32), np.uint8), 0.9)
    assert n == 0 and (m == -1).all()
    # octree on a handful of candidates: every isolated candidate survives
    c = np.array([[10, 10, 30], [300, 40, 25], [600, 400, 21]], np.float32)
    sel = oracle.distribute_octree(c, 16, 736, 16, 464, 217)
    assert sorted(sel.tolist()) == [0, 1, 2]


def test_search_in_radius_oracle_matches_brute_force():
    """Independent model of the Fuse / SearchBySim3 per-point search: brute force over all keypoints in index order
    restricted to the reference's cell range equals the grid walk whenever distances are unique."""
    rng = np.random.RandomState(11)
    n, nq = 800, 300
    keys = np.zeros(n, oracle.KEYPOINT_DTYPE)
    keys["x"] = rng.uniform(20, 730, n).astype(np.float32)
    keys["y"] = rng.uniform(20, 460, n).astype(np.float32)
    keys["octave"] = rng.randint(0, 4, n)
    desc = rng.randint(0, 256, (n, 32)).astype(np.uint8)
    from pl_vi_orbslam3_b200 import frame_grid
    from pl_vi_orbslam3_b200.capi import QUERY_DTYPE
    grid = frame_grid(0, 752, 0, 480)
    q = np.zeros(nq, QUERY_DTYPE)
    q["u"] = rng.uniform(0, 752, nq).astype(np.float32)
    q["v"] = rng.uniform(0, 480, nq).astype(np.float32)
    q["radius"] = rng.choice([20.0, 45.0], nq).astype(np.float32)
    q["min_level"] = rng.randint(0, 3, nq)
    q["max_level"] = q["min_level"] + 1
    qd = rng.randint(0, 256, (nq, 32)).astype(np.uint8)
    s2 = (1.0 / (1.2 ** np.arange(8)) ** 2).astype(np.float32)
    for chi2, th in ((5.99, 256), (0.0, 256), (400.0, 120)):
        found, bi, bd = oracle.search_in_radius(keys, desc, grid, q, qd, s2, chi2, th)
        for i in range(nq):
            dx, dy = np.abs(keys["x"] - q["u"][i]), np.abs(keys["y"] - q["v"][i])
            ok = (dx < q["radius"][i]) & (dy < q["radius"][i]) & (keys["octave"] >= q["min_level"][i]) & (keys["octave"] <= q["max_level"][i])
            if chi2 > 0:
                ex, ey = (q["u"][i] - keys["x"]).astype(np.float32), (q["v"][i] - keys["y"]).astype(np.float32)
                e2 = (ex * ex + ey * ey).astype(np.float32)
                ok &= ~((e2 * s2[keys["octave"]]).astype(np.float32).astype(np.float64) > chi2)
            d = np.unpackbits(desc ^ qd[i], axis=1).sum(axis=1)
            d = np.where(ok, d, 999)
            best = int(d.min()) if ok.any() else 256
            assert bd[i] == best
            if best <= th and (d == best).sum() == 1:
                assert bi[i] == int(d.argmin())
            if best > th:
                assert bi[i] == -1


def test_line_fuse_search_oracle_uses_the_shift25_distance_and_first_minimum():
    """LineMatcher::Fuse measures with LineMatcher::DescriptorDistance (sum of floor(popcount32 / 2), the >> 25 bug of
    src/LineMatcher.cpp:487-499), not with the Hamming distance: a numpy model of GetLinesInArea + that distance."""
    rng = np.random.RandomState(4)
    r = oracle.line_extract(synth.frame_euroc(41))
    kl, desc = r["keylines"], r["descriptors"]
    nq = 120
    src = rng.randint(0, len(kl), nq)
    q = np.zeros((nq, 6), np.float32)
    q[:, 0], q[:, 1] = kl["startPointX"][src] + 1, kl["startPointY"][src] - 1
    q[:, 2], q[:, 3] = kl["endPointX"][src] - 1, kl["endPointY"][src] + 1
    q[:, 4] = 30.0
    q[:, 5] = kl["octave"][src]
    qd = desc[src] ^ (rng.rand(nq, 32) < 0.1).astype(np.uint8)
    found, bi, bd = oracle.line_fuse_search(kl, desc, q, qd, None, 50)
    w = lambda a: np.ascontiguousarray(a).view(np.uint32)
    for i in range(nq):
        mx, my = 0.5 * np.float64(q[i, 0] + q[i, 2]), 0.5 * np.float64(q[i, 1] + q[i, 3])
        dist = ((mx - kl["pt_x"]) ** 2 + (my - kl["pt_y"]) ** 2).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            slope = np.float32(q[i, 1] - q[i, 3]) / np.float32(q[i, 0] - q[i, 2]) - kl["angle"]
        ok = ~(dist > q[i, 4] * q[i, 4]) & ~(slope.astype(np.float64) > np.float64(q[i, 4]) * 0.01)
        ok &= (kl["octave"] >= int(q[i, 5]) - 1) & (kl["octave"] <= int(q[i, 5]))
        x = w(desc) ^ w(qd[i])
        d = np.array([[bin(int(v)).count("1") >> 1 for v in row] for row in x]).sum(axis=1)
        d = np.where(ok, d, 1 << 30)
        if ok.any():
            assert bd[i] == d.min()
            assert bi[i] == (int(d.argmin()) if d.min() <= 50 else -1)
        else:
            assert bi[i] == -1
    assert found > 50


# ---- round-2 additions of the line path (tests/golden/lsd_extra.npz, tools/gen_golden_lsd.py) ---------------------------
X = np.load(GOLD / "lsd_extra.npz")


@pytest.mark.parametrize("scale", [0.5, 0.6, 0.9])
def test_lsd_other_scales_match_opencv(scale):
    """Gaussian kernel of flsd at other lsd_scale settings (OpenCV's soft-float exp, bit-exact) and cv::resize(f64) incl.
    the 2x2 area path at exactly 0.5 (CRC of the whole image)."""
    import math
    S = float(np.float32(scale))
    sigma = 0.6 / S
    n = 1 + 2 * int(math.ceil(sigma * math.sqrt(2 * 3.0 * math.log(10.0))))
    k = oracle.gaussian_kernel_f64(n, sigma)
    assert np.array_equal(k, X[f"cv2_kernel_{scale}"])
    img = synth.frame_euroc(9).astype(np.float64)
    mine = oracle.gaussian_blur_f64(img, k)
    h, w = (int(v) for v in X[f"cv2_resize_shape_{scale}"])
    assert crc(oracle.resize_linear_f64(mine, w, h, S, S)) == X[f"cv2_resize_crc_{scale}"]


@pytest.mark.parametrize("seed,scale,refine", [(0, 0.8, 1), (0, 0.8, 2), (3, 0.6, 1), (3, 1.0, 2), (5, 1.0, 0), (5, 0.5, 0)])
def test_lsd_refine_and_scales_equal_committed_reference_outputs(seed, scale, refine):
    """Raw segments of the reference's own lsd.cpp (compiled unmodified, oracle/_ref) for lsd_refine 1 / 2 and lsd_scale
    1.0 / 0.6 / 0.5, committed: the oracle's restatement equals them to the last bit."""
    ref = X[f"ref_lsd_{seed}_{scale}_{refine}"]
    got = oracle.lsd(synth.frame_euroc(seed), scale, refine=refine)
    assert len(ref) > 100 and np.array_equal(got, ref)


@pytest.mark.parametrize("seed", [1, 2, 9])
def test_stereo_line_depth_equals_committed_reference_outputs(seed):
    """Frame::ComputeStereoMatches_Lines of the reference's own Frame.cc (committed): disparities, depths, mvle_l."""
    import test_stereo_lines as SL
    s1, d1, s2, d2 = SL._depth_case(seed)
    n, m12 = oracle.line_match_grid(s1, d1, s2, d2, SL.INV_W, SL.INV_H)
    k, disp, dep, le = oracle.line_stereo_depth(s1, s2, m12, 47.9)
    out = X[f"ref_stereo_lines_{seed}_out"]
    assert k == int(X[f"ref_stereo_lines_{seed}_k"])
    assert np.array_equal(out[:, :2].view(np.uint32), disp.view(np.uint32)) and np.array_equal(out[:, 2:].view(np.uint32), dep.view(np.uint32))
    assert np.array_equal(X[f"ref_stereo_lines_{seed}_le"].view(np.uint64), le.view(np.uint64))
