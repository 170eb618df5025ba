"""GPU parity of the matchers against THE REFERENCE'S OWN src/ORBmatcher.cc and src/LineMatcher.cpp (compiled unmodified
against stand-in SLAM classes, see tests/test_oracle_vs_ref_matchers.py / test_oracle_vs_ref.py): the CUDA path through
the C ABI vs the committed reference outputs (tests/golden/ref_outputs.npz) and, where the prebuilt libraries
travelled with the snapshot, live runs of the reference.  Bar: bit-exact match tables and counts.
"""
import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200.matchers import FrameView, LineMatcher, ORBmatcher
from test_oracle_vs_ref import frame
import test_oracle_vs_ref as TL
import test_oracle_vs_ref_matchers as T
from test_oracle_vs_ref_matchers import GRID, R, SCALES, pair_features  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu
live = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref did not travel")


@pytest.fixture(scope="module")
def om(gpu):
    m = ORBmatcher(0.9, True, max_pairs=2, max_train=6000, max_query=6000)
    yield m
    m.close()


@pytest.fixture(scope="module")
def lm(gpu):
    m = LineMatcher(max_pairs=4, max_train=512, max_query=512)
    yield m
    m.close()


def cuda_mappoints(om, r2, c, qdesc, nnratio):
    om.mfNNratio = nnratio
    F = FrameView(r2["keypoints"], r2["descriptors"], GRID, c["blocked"])
    n, mt, _ = om.SearchByProjection(F, T.mappoint_queries(c), qdesc, mappoints=True)
    return n, mt


def cuda_init(om, r1, r2, prev, window, nnratio, ori):
    om.mfNNratio, om.mbCheckOrientation = nnratio, ori
    F2 = FrameView(r2["keypoints"], r2["descriptors"], GRID)
    out = om.SearchForInitialization(r1["keypoints"], r1["descriptors"], F2, prev, window)
    om.mbCheckOrientation = True
    return out


def cuda_bow_kf_f(om, r1, r2, fv1, fv2, mp1, nnratio, ori):
    om.mfNNratio, om.mbCheckOrientation = nnratio, ori
    items, q, order = T.bow_queries(r1, fv1, fv2, mp1)
    n, mt, _ = om.SearchByBoW(FrameView(r2["keypoints"], r2["descriptors"], GRID), items, q, r1["descriptors"][order])
    om.mbCheckOrientation = True
    return n, np.where(mt >= 0, order[np.maximum(mt, 0)], -1).astype(np.int32)


def cuda_bow_kfkf(om, r1, r2, fv1, fv2, mp1, mp2, nnratio, ori):
    om.mfNNratio, om.mbCheckOrientation = nnratio, ori
    items, q, order = T.bow_queries(r1, fv1, fv2, mp1)
    n, mq = om.SearchByBoW_KF(FrameView(r2["keypoints"], r2["descriptors"], GRID), mp2 == 1, items, q, r1["descriptors"][order])
    om.mbCheckOrientation = True
    got = np.full(len(r1["keypoints"]), -1, np.int32)
    got[order[mq >= 0]] = mq[mq >= 0]
    return n, got


def test_cuda_equals_reference_orbmatcher_outputs(om, pair_features):
    r1, r2, A = pair_features
    c = T.mappoint_case(r1, r2, A, 1, 3.0)
    n, mt = cuda_mappoints(om, r2, c, r1["descriptors"], 0.8)
    assert n == int(R["orbmatch/mappoints_n"]) and np.array_equal(mt, R["orbmatch/mappoints"])
    k1 = r1["keypoints"]
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    n, m12, pm = cuda_init(om, r1, r2, prev, 100, 0.9, True)
    assert n == int(R["orbmatch/init_n"]) and np.array_equal(m12, R["orbmatch/init"]) and np.array_equal(pm, R["orbmatch/init_prev"])
    fv1, fv2, mp1, mp2 = T.bow_case(r1, r2, 6, 3, 2, 9)
    n, mt = cuda_bow_kf_f(om, r1, r2, fv1, fv2, mp1, 0.7, True)
    assert n == int(R["orbmatch/bow_n"]) and np.array_equal(mt, R["orbmatch/bow"])
    n, m = cuda_bow_kfkf(om, r1, r2, fv1, fv2, mp1, mp2, 0.8, True)
    assert n == int(R["orbmatch/bowkf_n"]) and np.array_equal(m, R["orbmatch/bowkf"])


def test_cuda_equals_reference_linematcher_outputs(lm):
    d1 = oracle.line_extract(frame("synth_0"))["descriptors"]
    d2 = oracle.line_extract(frame("synth_1"))["descriptors"]
    n, m = lm.match(d1, d2, 0.75)
    assert n == int(R["linematch/match_n"]) and np.array_equal(m, R["linematch/match"])
    ms, nm, _ = lm._match_mad_batch([(d1, d2)], 0.5)
    assert nm[0] == int(R["linematch/init_n"]) and np.array_equal(ms[0], R["linematch/init"])
    ms, nm, _ = lm._match_mad_batch([(d1, d2)], 0.1, [(R["linematch/has1"], R["linematch/has2"])])
    assert nm[0] == int(R["linematch/tri_n"]) and np.array_equal(ms[0], R["linematch/tri"])
    kl, desc, sf, q, qd, bad = TL.line_fuse_case(1, 8.0)
    n, bi, _ = lm.FuseSearch(kl, desc, q, qd, TL.line_fuse_flags(q, bad))
    assert n == int(R["linematch/fuse_n"]) and np.array_equal(bi, R["linematch/fuse"])


@live
@pytest.mark.parametrize("seed,th,nnratio", [(0, 1.0, 0.8), (2, 5.0, 0.9), (3, 15.0, 0.6)])
def test_cuda_equals_live_reference_search_by_projection(om, pair_features, seed, th, nnratio):
    r1, r2, A = pair_features
    c = T.mappoint_case(r1, r2, A, seed, th)
    n, mt = cuda_mappoints(om, r2, c, r1["descriptors"], nnratio)
    rn, rmt = oracle.ref_search_mappoints(r2["keypoints"], r2["descriptors"], GRID, SCALES, c["proj"], c["viewcos"], c["level"],
                                          c["flags"], r1["descriptors"], th, nnratio, c["blocked"])
    assert n == rn and np.array_equal(mt, rmt)


@live
@pytest.mark.parametrize("window,nnratio,ori", [(100, 0.9, True), (30, 0.9, True), (100, 0.7, False)])
def test_cuda_equals_live_reference_search_for_initialization(om, pair_features, window, nnratio, ori):
    r1, r2, _ = pair_features
    k1 = r1["keypoints"]
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    n, m12, pm = cuda_init(om, r1, r2, prev, window, nnratio, ori)
    rn, rm12, rpm = oracle.ref_search_init(k1, r1["descriptors"], r2["keypoints"], r2["descriptors"], GRID, prev, window, nnratio, ori)
    assert n == rn and np.array_equal(m12, rm12) and np.array_equal(pm, rpm)


@live
@pytest.mark.parametrize("k,L,levelsup,nnratio,ori", [(10, 4, 2, 0.9, True), (4, 2, 1, 0.6, True), (5, 3, 3, 0.75, False)])
def test_cuda_equals_live_reference_search_by_bow(om, pair_features, k, L, levelsup, nnratio, ori):
    r1, r2, _ = pair_features
    fv1, fv2, mp1, mp2 = T.bow_case(r1, r2, k, L, levelsup, k + L)
    n, mt = cuda_bow_kf_f(om, r1, r2, fv1, fv2, mp1, nnratio, ori)
    rn, rmt = oracle.ref_search_bow_kf_f(r1["keypoints"], r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], fv2, nnratio, ori)
    assert n == rn and np.array_equal(mt, rmt)
    n, m = cuda_bow_kfkf(om, r1, r2, fv1, fv2, mp1, mp2, nnratio, ori)
    rn, rm = oracle.ref_search_bow_kfkf(r1["keypoints"], r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], mp2, fv2,
                                        nnratio, ori)
    assert n == rn and np.array_equal(m, rm)


@live
@pytest.mark.parametrize("seed", range(4))
def test_cuda_equals_live_reference_line_matcher(lm, seed):
    rng = np.random.default_rng(seed)
    n1, n2 = (int(x) for x in rng.integers(2, 300, 2))
    d1, d2 = __import__("test_oracle_vs_ref")._line_desc_pair(rng, n1, n2, [0.0, 0.02, 0.1, 0.3][seed])
    for nnr in (0.75, 1.0):
        n, m = lm.matchNNR(d1, d2, nnr)
        rn, rm = oracle.ref_line_match(d1, d2, nnr, "nnr")
        assert n == rn and np.array_equal(m, rm)
        n, m = lm.match(d1, d2, nnr)
        rn, rm = oracle.ref_line_match(d1, d2, nnr, "maplines")
        assert n == rn and np.array_equal(m, rm)
    ms, nm, _ = lm._match_mad_batch([(d1, d2)], 0.5)
    rn, rm = oracle.ref_line_match_mad(d1, d2, 0.5)
    assert nm[0] == rn and np.array_equal(ms[0], rm)
    h1, h2 = (rng.random(n1) < 0.3).astype(np.uint8), (rng.random(n2) < 0.3).astype(np.uint8)
    ms, nm, _ = lm._match_mad_batch([(d1, d2)], 0.1, [(h1, h2)])
    rn, rm = oracle.ref_line_match_mad(d1, d2, 0.1, h1, h2)
    assert nm[0] == rn and np.array_equal(ms[0], rm)


# ---- the searches behind pose arithmetic (identity poses in the reference run, see test_oracle_vs_ref_matchers.py) ----
def cuda_frame(om, r1, r2, c, ori):
    om.mbCheckOrientation = ori
    F = FrameView(r2["keypoints"], r2["descriptors"], GRID, c["blocked"])
    n, mt, _ = om.SearchByProjection(F, T.frame_queries(r1, c), r1["descriptors"])
    om.mbCheckOrientation = True
    return n, mt


def cuda_triangulation(om, r1, r2, fv1, fv2, mp1, mp2, F12, ep, sg2, coarse):
    k1 = r1["keypoints"]
    start2 = {int(nd): (int(fv2[1][i]), int(fv2[1][i + 1])) for i, nd in enumerate(fv2[0])}
    order, rows = [], []
    for i, nd in enumerate(fv1[0]):
        if int(nd) not in start2:
            continue
        for idx1 in fv1[2][fv1[1][i]:fv1[1][i + 1]]:
            if mp1[idx1]:
                continue
            order.append(int(idx1))
            rows.append((k1["x"][idx1], k1["y"][idx1], start2[int(nd)][0], start2[int(nd)][1], k1["angle"][idx1]))
    qs = np.zeros(len(rows), T.QUERY_DTYPE)
    for j, (u, v, s, e, ang) in enumerate(rows):
        qs[j]["u"], qs[j]["v"], qs[j]["min_level"], qs[j]["max_level"], qs[j]["angle"] = u, v, s, e, ang
    n, mq = om.SearchForTriangulation(FrameView(r2["keypoints"], r2["descriptors"], GRID), mp2, np.asarray(fv2[2], np.int32), qs,
                                      r1["descriptors"][order], F12, ep, SCALES, sg2, coarse)
    got = np.full(len(k1), -1, np.int32)
    got[np.array(order, np.int64)] = mq
    return n, got


def cuda_fuse(om, r2, c, qdesc, chi2):
    n, bi, _ = om.SearchInRadius(FrameView(r2["keypoints"], r2["descriptors"], GRID), T.kf_queries(c), qdesc, T.INV_SIGMA2, chi2, 50)
    return n, bi


def cuda_kf(om, r2, c, qdesc, ratio):
    KF = FrameView(r2["keypoints"], r2["descriptors"], GRID, c["matched_in"])
    n, mt, _ = om.SearchByProjection_KF(KF, T.kf_queries(c), qdesc, ratio)
    return n, mt


def test_cuda_equals_reference_orbmatcher_outputs_projection(om, pair_features):
    r1, r2, A = pair_features
    k1 = r1["keypoints"]
    n, mt = cuda_frame(om, r1, r2, T.frame_case(r1, r2, A, 1, 7.0), True)
    assert n == int(R["orbmatch/frame_n"]) and np.array_equal(mt, R["orbmatch/frame"])
    fv1, fv2, mp1, mp2, F12, ep, sg2 = T.triangulation_case(r1, r2, 6, 3, 2, 3)
    n, m = cuda_triangulation(om, r1, r2, fv1, fv2, mp1, mp2, F12, ep, sg2, False)
    assert n == int(R["orbmatch/tri_n"]) and np.array_equal(m, R["orbmatch/tri"])
    c = T.kf_case(r1, r2, A, 1, 4.0)
    for key, chi2 in (("fuse", 5.99), ("fuse_sim3", 0.0)):
        n, bi = cuda_fuse(om, r2, c, r1["descriptors"], chi2)
        assert n == int(R[f"orbmatch/{key}_n"]) and np.array_equal(bi, R[f"orbmatch/{key}"])
    n, mt = cuda_kf(om, r2, T.kf_case(r1, r2, A, 1, 15.0), r1["descriptors"], 1.5)
    assert n == int(R["orbmatch/kf_n"]) and np.array_equal(mt, R["orbmatch/kf"])


@live
@pytest.mark.parametrize("seed,th,ori", [(0, 15.0, True), (2, 30.0, False)])
def test_cuda_equals_live_reference_search_by_projection_frame(om, pair_features, seed, th, ori):
    r1, r2, A = pair_features
    c = T.frame_case(r1, r2, A, seed, th)
    n, mt = cuda_frame(om, r1, r2, c, ori)
    rn, rmt = oracle.ref_search_frame(r2["keypoints"], r2["descriptors"], GRID, T.BOUNDS, SCALES, r1["keypoints"], c["uv"], c["flags"],
                                      r1["descriptors"], th, ori, c["blocked"])
    assert n == rn and np.array_equal(mt, rmt)


@live
@pytest.mark.parametrize("k,L,levelsup,coarse,seed", [(6, 3, 2, False, 0), (4, 2, 1, True, 2), (3, 2, 2, False, 4)])
def test_cuda_equals_live_reference_search_for_triangulation(om, pair_features, k, L, levelsup, coarse, seed):
    r1, r2, _ = pair_features
    fv1, fv2, mp1, mp2, F12, ep, sg2 = T.triangulation_case(r1, r2, k, L, levelsup, seed)
    n, m = cuda_triangulation(om, r1, r2, fv1, fv2, mp1, mp2, F12, ep, sg2, coarse)
    rn, rm = oracle.ref_search_triangulation(r1["keypoints"], r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], mp2, fv2,
                                             F12, ep, SCALES, sg2, sg2, coarse, True)
    assert n == rn and np.array_equal(m, rm)


@live
@pytest.mark.parametrize("seed,th,sim3", [(0, 3.0, False), (3, 7.5, True)])
def test_cuda_equals_live_reference_fuse(om, pair_features, seed, th, sim3):
    r1, r2, A = pair_features
    c = T.kf_case(r1, r2, A, seed, th)
    n, bi = cuda_fuse(om, r2, c, r1["descriptors"], 0.0 if sim3 else 5.99)
    rn, rbi = oracle.ref_fuse(r2["keypoints"], r2["descriptors"], GRID, T.BOUNDS, SCALES, T.INV_SIGMA2, c["uv"], c["level"], c["flags"],
                              r1["descriptors"], th, sim3)
    assert n == rn and np.array_equal(bi, rbi)


@live
@pytest.mark.parametrize("seed,th,ratio", [(0, 8, 1.0), (1, 15, 1.5), (3, 15, 0.8)])
def test_cuda_equals_live_reference_search_by_projection_keyframe(om, pair_features, seed, th, ratio):
    r1, r2, A = pair_features
    c = T.kf_case(r1, r2, A, seed, float(th))
    n, mt = cuda_kf(om, r2, c, r1["descriptors"], ratio)
    rn, rmt = oracle.ref_search_by_projection_kf(r2["keypoints"], r2["descriptors"], GRID, T.BOUNDS, SCALES, c["uv"], c["level"],
                                                 c["flags"], r1["descriptors"], th, ratio, c["matched_in"])
    assert n == rn and np.array_equal(mt, rmt)


@live
@pytest.mark.parametrize("seed,th", [(0, 3.0), (2, 20.0), (3, 60.0)])
def test_cuda_equals_live_reference_line_fuse(lm, seed, th):
    kl, desc, sf, q, qd, bad = TL.line_fuse_case(seed, th)
    n, bi, _ = lm.FuseSearch(kl, desc, q, qd, TL.line_fuse_flags(q, bad))
    rn, rbi = oracle.ref_line_fuse(kl, desc, TL.LINE_BOUNDS, sf, q, qd, bad, th)
    assert n == rn and np.array_equal(bi, rbi)


def test_cuda_equals_reference_distinctive_descriptors(gpu):
    """plvi_distinctive_descriptors vs the reference's own MapPoint::ComputeDistinctiveDescriptors (committed outputs and,
    where the library travelled, live)."""
    import torch
    from pl_vi_orbslam3_b200.matchers import compute_distinctive_descriptors
    for seed, golden in ((7, True), (0, False), (2, False)):
        desc, counts = T.distinctive_case(seed)
        idx, best = compute_distinctive_descriptors(torch.from_numpy(desc).cuda(), torch.from_numpy(counts).cuda())
        best = best.cpu().numpy()
        if golden:
            assert np.array_equal(best, R["mappoint/distinctive"])
        elif oracle.ref_available():
            for p in range(len(counts)):
                assert np.array_equal(best[p], oracle.ref_distinctive_descriptor(desc[p, :counts[p]])), p
                # MapLine::ComputeDistinctiveDescriptors (src/MapLine.cc:264-329) is the same rule on LBD descriptors
                assert np.array_equal(best[p], oracle.ref_mapline_distinctive_descriptor(desc[p, :counts[p]])), p


def cuda_sim3(om, r1, r2, case, th):
    (uv1, l1, f1), (uv2, l2, f2) = case
    _, bi12, _ = om.SearchInRadius(FrameView(r2["keypoints"], r2["descriptors"], GRID), T.sim3_queries(uv1, l1, f1, th), r1["descriptors"],
                                   T.INV_SIGMA2, 0.0, 100)
    _, bi21, _ = om.SearchInRadius(FrameView(r1["keypoints"], r1["descriptors"], GRID), T.sim3_queries(uv2, l2, f2, th), r2["descriptors"],
                                   T.INV_SIGMA2, 0.0, 100)
    return T.sim3_mutual(bi12, bi21)


def test_cuda_equals_reference_search_by_sim3(om, pair_features):
    """SearchBySim3 = plvi_search_in_radius once per direction (TH_HIGH, no chi2 gate) + the caller's agreement test."""
    r1, r2, A = pair_features
    n, m = cuda_sim3(om, r1, r2, T.sim3_case(r1, r2, A, 1), 7.5)
    assert n == int(R["orbmatch/sim3_n"]) and np.array_equal(m, R["orbmatch/sim3"])
    if oracle.ref_available():
        for seed, th in ((0, 7.5), (2, 15.0)):
            case = T.sim3_case(r1, r2, A, seed)
            n, m = cuda_sim3(om, r1, r2, case, th)
            (uv1, l1, f1), (uv2, l2, f2) = case
            rn, rm = oracle.ref_search_by_sim3(r1["keypoints"], r1["descriptors"], uv1, l1, f1, r2["keypoints"], r2["descriptors"], uv2,
                                               l2, f2, GRID, T.BOUNDS, SCALES, th)
            assert n == rn and np.array_equal(m, rm)


@live
@pytest.mark.parametrize("seed,th,orb_dist,ori", [(0, 10.0, 100, True), (1, 3.0, 64, True), (2, 10.0, 100, False)])
def test_cuda_equals_live_reference_search_by_projection_relocalisation(om, pair_features, seed, th, orb_dist, ori):
    """SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) = mode FRAME with th_dist = ORBdist."""
    r1, r2, A = pair_features
    c = T.reloc_case(r1, r2, A, seed, th)
    om.mbCheckOrientation = ori
    F = FrameView(r2["keypoints"], r2["descriptors"], GRID, c["matched_in"])
    mt, _, nm, _ = om.search_batch(0, [F], [T.reloc_queries(r1, c)], [r1["descriptors"]], orb_dist)
    om.mbCheckOrientation = True
    rn, rmt = oracle.ref_search_reloc(r2["keypoints"], r2["descriptors"], GRID, T.BOUNDS, SCALES, r1["keypoints"], c["uv"], c["level"],
                                      c["flags"], r1["descriptors"], th, orb_dist, ori, c["matched_in"])
    assert int(nm[0]) == rn and np.array_equal(mt[0], rmt)
