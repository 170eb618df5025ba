"""GPU: DROP-IN PROOF.  The reference's unmodified src/Frame.cc / src/KeyFrame.cc (+ the same test glue and stand-in
collaborator classes) are compiled twice by oracle/Makefile.ref:

  libplvi_ref*.so     with the reference's own ORBextractor.cc / LineExtractor.cc / lsd.cpp / ... / ORBmatcher.cc / LineMatcher.cpp
  libplvi_dropin*.so  with the PRODUCT's drop-in headers (pl_vi_orbslam3_b200/shim/include first on the include path),
                      shim/src/ORBmatcher.cc, shim/src/LineMatcher.cpp and libplvi_cuda.so underneath.

Every test calls the same glue entry point on both builds and requires identical results: the Frame members the
reference's own constructor fills (mvKeys, mvKeysUn, mDescriptors, mvKeys_Line, mvKeysUn_Line, mDescriptors_Line,
mvKeyLineFunctions, mGrid, scale tables), mvpMapPoints after ORBmatcher::SearchByProjection and matches_12 after
LineMatcher::match (the call pattern of src/Tracking.cc:3957,3990), and the result of every reference-signature search
(ORBmatcher x13, LineMatcher x7), incl. the rectified-stereo branches and Frame::ComputeStereoMatches_Lines.
Bar: bit-exact.
"""
from pathlib import Path

import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import synth
import test_oracle_vs_ref_matchers as T
from test_oracle_vs_ref_matchers import BOUNDS, GRID, INV_SIGMA2, SCALES, pair_features  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
needs_libs = pytest.mark.skipif(not (oracle.ref_available() and oracle.dropin_available()),
                                reason="oracle/_ref (reference + drop-in builds) did not travel")


def both(fn, *a, **kw):
    """fn on the all-reference build and on the drop-in build."""
    ref = fn(*a, **kw)
    with oracle.dropin():
        got = fn(*a, **kw)
    return ref, got


def same(ref, got):
    assert len(ref) == len(got)
    for r, g in zip(ref, got):
        if isinstance(r, np.ndarray):
            assert r.shape == g.shape and np.array_equal(r, g), (r[:16], g[:16])
        else:
            assert r == g, (r, g)


def real_frames():
    out = []
    for name in ("frame_data2_1.npz", "frame_data2_3.npz"):
        p = GOLD / name
        if p.exists():
            z = np.load(p)
            if "img" in z.files:
                out.append(np.ascontiguousarray(z["img"]))
    return out


# ---------------------------------------------------------------------------------------------------------------------
@needs_libs
@pytest.mark.parametrize("seed", [3, 11, 200])
def test_frame_constructor_and_tracking_calls(gpu, seed):
    """Frame::Frame(imGray, ..., ORBextractor*, Lineextractor*, ...) (src/Frame.cc:537-642) on two frames, then
    SearchByProjection(Cur, Last, th, bMono) + LineMatcher::match (src/Tracking.cc:3957,3990): all members equal."""
    f1, f2, _ = synth.warp_pair(seed)

    def run():
        t = oracle.RefTracker()
        n1, n2 = t.frame(f1), t.frame(f2)
        cur, last = t.members(0), t.members(1)
        rng = np.random.RandomState(seed)
        obs0 = (rng.rand(n1) < 0.1).astype(np.uint8)
        res = [t.search(15.0, True, None, obs0, (0.0, 0.0, 0.0), 0.9, True, 0.9),
               t.search(7.0, True, None, None, (0.01, -0.02, 0.0), 0.9, False, 0.8)]
        t.close()
        return n1, n2, cur, last, res

    (n1, n2, cur, last, res), (gn1, gn2, gcur, glast, gres) = both(run)
    assert (n1, n2) == (gn1, gn2) and n1 > 500
    for a, b in ((cur, gcur), (last, glast)):
        assert a.keys() == b.keys()
        for k in a:
            if k == "mvKeysUn_Line":
                # Frame::UndistortKeyLines resize()s the vector and assigns the four end-point fields only
                # (src/Frame.cc:1189-1196); KeyLine's default constructor leaves the other fields uninitialised
                from pl_vi_orbslam3_b200.capi import KEYLINE_DTYPE
                x, y = np.frombuffer(a[k], KEYLINE_DTYPE), np.frombuffer(b[k], KEYLINE_DTYPE)
                for fld in ("startPointX", "startPointY", "endPointX", "endPointY"):
                    assert np.array_equal(x[fld], y[fld]), fld
                assert len(x) > 50 and not np.array_equal(x["startPointX"], np.frombuffer(a["mvKeys_Line"], KEYLINE_DTYPE)["startPointX"])
                continue
            assert np.array_equal(a[k], b[k]), k
    for (k, pts, nl, m12), (gk, gpts, gnl, gm12) in zip(res, gres):
        assert k == gk and np.array_equal(pts, gpts) and nl == gnl and np.array_equal(m12, gm12)
    assert res[0][0] > 200 and res[0][2] > 50


@needs_libs
def test_frame_constructor_other_sizes_and_real_frames(gpu):
    """640x480 / 1280x720 synthetic frames, 2000 features, and the reference's own data2/color frames (committed under
    tests/golden) through the reference's Frame constructor on both builds."""
    imgs = [synth.frame_euroc(5, 640, 480), synth.frame_euroc(6, 1280, 720)] + real_frames()
    for img in imgs:
        def run():
            t = oracle.RefTracker(nfeatures=2000)
            n = t.frame(img, K=(500.0, 500.0, img.shape[1] / 2.0, img.shape[0] / 2.0), dist=(0.0, 0.0, 0.0, 0.0))
            m = t.members(0)
            t.close()
            return n, m
        (n, m), (gn, gm) = both(run)
        assert n == gn and n > 300
        for k in m:
            assert np.array_equal(m[k], gm[k]), (img.shape, k)


@needs_libs
def test_extractor_entry_points(gpu):
    """ORBextractor::operator() / Lineextractor::operator() through the same glue (ref_glue.cpp) on both builds, incl.
    the lapping-area split, mvImagePyramid and other parameter sets."""
    img = synth.frame_euroc(21)
    for kw in (dict(), dict(lapping=(200, 500)), dict(nfeatures=2000, scale_factor=1.1, nlevels=6, ini_th=15, min_th=5)):
        r, g = both(oracle.ref_orb_extract, img, debug=True, **kw)
        assert r["mono_index"] == g["mono_index"]
        assert np.array_equal(r["keypoints"], g["keypoints"]) and np.array_equal(r["descriptors"], g["descriptors"])
        for a, b in zip(r["pyramid"], g["pyramid"]):
            assert np.array_equal(a, b)
    for kw in (dict(), dict(lsd_nfeatures=100), dict(nlevels=1)):
        r, g = both(oracle.ref_line_extract, img, **kw)
        for k in ("keylines", "descriptors", "line_eq"):
            assert np.array_equal(r[k], g[k]), k


# ---- the reference-signature searches --------------------------------------------------------------------------------
@needs_libs
def test_orbmatcher_signatures_on_standin_classes(gpu, pair_features):
    r1, r2, A = pair_features
    k1, d1, k2, d2 = r1["keypoints"], r1["descriptors"], r2["keypoints"], r2["descriptors"]
    for seed, th, nn in ((0, 1.0, 0.8), (1, 3.0, 0.8), (3, 15.0, 0.6)):
        c = T.mappoint_case(r1, r2, A, seed, th)
        same(*both(oracle.ref_search_mappoints, k2, d2, GRID, SCALES, c["proj"], c["viewcos"], c["level"], c["flags"], d1, th, nn, c["blocked"]))
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    for window, nn, ori in ((100, 0.9, True), (30, 0.9, True), (100, 0.7, False)):
        same(*both(oracle.ref_search_init, k1, d1, k2, d2, GRID, prev, window, nn, ori))
    for k, L, lu, nn, ori in ((6, 3, 2, 0.7, True), (10, 4, 2, 0.9, True), (5, 3, 3, 0.75, False)):
        fv1, fv2, mp1, mp2 = T.bow_case(r1, r2, k, L, lu, k + L)
        same(*both(oracle.ref_search_bow_kf_f, k1, d1, mp1, fv1, k2, d2, fv2, nn, ori))
        same(*both(oracle.ref_search_bow_kfkf, k1, d1, mp1, fv1, k2, d2, mp2, fv2, nn, ori))
    for seed, th, ori in ((0, 15.0, True), (1, 7.0, True), (2, 30.0, False)):
        c = T.frame_case(r1, r2, A, seed, th)
        same(*both(oracle.ref_search_frame, k2, d2, GRID, BOUNDS, SCALES, k1, c["uv"], c["flags"], d1, th, ori, c["blocked"]))
    for k, L, lu, coarse, seed in ((6, 3, 2, False, 0), (4, 2, 1, True, 2), (6, 3, 2, False, 3)):
        fv1, fv2, mp1, mp2, F12, ep, sg2 = T.triangulation_case(r1, r2, k, L, lu, seed)
        same(*both(oracle.ref_search_triangulation, k1, d1, mp1, fv1, k2, d2, mp2, fv2, F12, ep, SCALES, sg2, sg2, coarse, True))
    for seed, th, sim3 in ((0, 3.0, False), (2, 3.0, True), (3, 7.5, True)):
        c = T.kf_case(r1, r2, A, seed, th)
        same(*both(oracle.ref_fuse, k2, d2, GRID, BOUNDS, SCALES, INV_SIGMA2, c["uv"], c["level"], c["flags"], d1, th, sim3))
    for seed, th, ratio in ((0, 8, 1.0), (1, 15, 1.5), (3, 15, 0.8)):
        c = T.kf_case(r1, r2, A, seed, float(th))
        same(*both(oracle.ref_search_by_projection_kf, k2, d2, GRID, BOUNDS, SCALES, c["uv"], c["level"], c["flags"], d1, th, ratio, c["matched_in"]))
    for seed, th in ((0, 7.5), (2, 15.0)):
        (uv1, l1, f1), (uv2, l2, f2) = T.sim3_case(r1, r2, A, seed)
        same(*both(oracle.ref_search_by_sim3, k1, d1, uv1, l1, f1, k2, d2, uv2, l2, f2, GRID, BOUNDS, SCALES, th))
    for seed, th, dist, ori in ((0, 10.0, 100, True), (1, 3.0, 64, True), (2, 10.0, 100, False)):
        c = T.reloc_case(r1, r2, A, seed, th)
        same(*both(oracle.ref_search_reloc, k2, d2, GRID, BOUNDS, SCALES, k1, c["uv"], c["level"], c["flags"], d1, th, dist, ori, c["matched_in"]))
    for i in range(32):
        r, g = both(oracle.ref_orb_descriptor_distance, d1[i], d2[i])
        assert r == g


@needs_libs
def test_orbmatcher_signatures_on_the_reference_frame_and_keyframe_classes(gpu, pair_features):
    """The same adapters compiled against the reference's OWN Frame.h / KeyFrame.h (KeyFrame built by its constructor)."""
    r1, r2, A = pair_features
    k1, d1, k2, d2 = r1["keypoints"], r1["descriptors"], r2["keypoints"], r2["descriptors"]
    c = T.mappoint_case(r1, r2, A, 2, 5.0)
    same(*both(oracle.ref_search_mappoints, k2, d2, GRID, SCALES, c["proj"], c["viewcos"], c["level"], c["flags"], d1, 5.0, 0.9, c["blocked"],
               real_frame_bounds=BOUNDS))
    c = T.frame_case(r1, r2, A, 0, 15.0)
    same(*both(oracle.ref_search_frame, k2, d2, GRID, BOUNDS, SCALES, k1, c["uv"], c["flags"], d1, 15.0, True, c["blocked"], real_frame=True))
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    same(*both(oracle.ref_search_init, k1, d1, k2, d2, GRID, prev, 100, 0.9, True, real_frame_bounds=BOUNDS))
    for seed, th, sim3 in ((0, 3.0, False), (3, 7.5, True)):
        c = T.kf_case(r1, r2, A, seed, th)
        same(*both(oracle.ref_fuse_real, k2, d2, BOUNDS, SCALES, INV_SIGMA2, c["uv"], c["level"], c["flags"], d1, th, sim3))
    c = T.kf_case(r1, r2, A, 1, 15.0)
    same(*both(oracle.ref_search_by_projection_kf_real, k2, d2, BOUNDS, SCALES, c["uv"], c["level"], c["flags"], d1, 15, 1.5, c["matched_in"]))
    fv1, fv2, mp1, _ = T.bow_case(r1, r2, 6, 3, 2, 9)
    same(*both(oracle.ref_search_bow_kf_f_real, k1, d1, mp1, fv1, k2, d2, fv2, BOUNDS, 0.7, True))


@needs_libs
def test_rectified_stereo_branches(gpu, pair_features):
    """mvuRight > 0 window test, bForward / bBackward level ranges (src/ORBmatcher.cc:1982-2047) and the 3-dof gate of
    Fuse (:1530-1543) on the reference's own Frame / KeyFrame."""
    r1, r2, A = pair_features
    k1, d1, k2, d2 = r1["keypoints"], r1["descriptors"], r2["keypoints"], r2["descriptors"]
    K = (458.654, 457.296, 367.215, 248.375)
    mbf = 47.9
    rng = np.random.RandomState(5)
    nonzero = 0
    for tz in (0.0, 0.3, -0.3):   # bf / fx = 0.104: |tz| = 0.3 switches bForward / bBackward on
        depth = rng.uniform(2.0, 12.0, len(k1)).astype(np.float32)
        flags = rng.choice([0, 0, 0, 0, 1, 2], len(k1)).astype(np.int32)
        # right coordinates of the current frame: consistent with a depth for 70 % of the features, random for 15 %, none for 15 %
        d2z = rng.uniform(2.0, 12.0, len(k2)).astype(np.float32)
        ur = (k2["x"] - np.float32(mbf) / d2z).astype(np.float32)
        u = rng.rand(len(k2))
        ur = np.where(u < 0.15, np.float32(-1.0), np.where(u < 0.3, rng.uniform(0, 700, len(k2)).astype(np.float32), ur)).astype(np.float32)
        blocked = (rng.rand(len(k2)) < 0.1).astype(np.uint8)
        ref, got = both(oracle.ref_search_frame_stereo, k2, d2, ur, BOUNDS, SCALES, k1, depth, flags, d1, K, mbf, (0.0, 0.0, tz), 15.0, True, blocked)
        same(ref, got)
        nonzero += ref[0]
    assert nonzero > 50
    for seed in (0, 1):
        c = T.kf_case(r1, r2, A, seed, 3.0)
        depth = rng.uniform(2.0, 12.0, len(k1)).astype(np.float32)
        d2z = rng.uniform(2.0, 12.0, len(k2)).astype(np.float32)
        ur = np.where(rng.rand(len(k2)) < 0.3, np.float32(-1.0), (k2["x"] - np.float32(mbf) / d2z)).astype(np.float32)
        ref, got = both(oracle.ref_fuse_stereo, k2, d2, ur, BOUNDS, SCALES, INV_SIGMA2, c["uv"], depth, c["level"], c["flags"], d1, K, mbf, 3.0)
        same(ref, got)
        assert ref[0] > 20


@needs_libs
def test_linematcher_signatures(gpu):
    f1, f2, _ = synth.warp_pair(7)
    l1, l2 = oracle.line_extract(f1), oracle.line_extract(f2)
    a, b = l1["descriptors"], l2["descriptors"]
    for variant in ("nnr", "match", "maplines"):
        for nnr in (0.9, 0.7):
            same(*both(oracle.ref_line_match, a, b, nnr, variant))
    rng = np.random.RandomState(1)
    h1, h2 = (rng.rand(len(a)) < 0.3), (rng.rand(len(b)) < 0.3)
    same(*both(oracle.ref_line_match_mad, a, b, 0.5))
    same(*both(oracle.ref_line_match_mad, a, b, 0.1, h1, h2))
    for i in range(16):
        for which in (0, 1):
            r, g = both(oracle.ref_line_distance, a[i], b[i], which)
            assert r == g


@needs_libs
def test_line_fuse_and_stereo_lines(gpu):
    """LineMatcher::Fuse and Frame::ComputeStereoMatches_Lines (grid fill + matchGrid + the disparity / depth filter that
    follows the search, src/Frame.cc:1408-1529) -- Frame.cc unmodified on both builds."""
    import test_stereo_lines as SL
    import test_oracle_vs_ref as TL
    for seed, th in ((0, 3.0), (1, 8.0), (3, 60.0)):
        kl, desc, sf, q, qd, bad = TL.line_fuse_case(seed, th)
        same(*both(oracle.ref_line_fuse, kl, desc, TL.LINE_BOUNDS, sf, q, qd, bad, th))
    f1 = synth.frame_euroc(31)
    f2 = np.roll(f1, -9, axis=1)
    l1, l2 = oracle.line_extract(f1), oracle.line_extract(f2)
    ref, got = both(oracle.ref_frame_stereo_lines, l1["keylines"], l1["descriptors"], l2["keylines"], l2["descriptors"], 64.0 / 752, 48.0 / 480, 47.9)
    same(ref, got)
    assert ref[0] > 5
    for seed, (n1, n2) in enumerate(SL.SIZES):
        if n1 == 0:
            continue
        s1, d1, s2, d2 = SL.stereo_line_case(seed, n1, n2)
        same(*both(oracle.ref_line_match_grid, s1, d1, s2, d2, SL.INV_W, SL.INV_H))
