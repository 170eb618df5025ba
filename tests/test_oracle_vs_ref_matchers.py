"""CPU: the oracle's ORB searches against THE REFERENCE'S OWN src/ORBmatcher.cc.

oracle/_ref/libplvi_ref_orbmatcher.so = ORBmatcher.cc compiled unmodified, where it lies, with the stand-in Frame /
KeyFrame / MapPoint of oracle/cvmini/slam_mock_orb.h force-included in place of the reference's headers (which need
DBoW2's vocabulary, g2o, Sophus, boost, the Atlas).  The stand-ins carry plain data; Frame::GetFeaturesInArea
(Frame.cc, not compilable for the same reason) is the oracle's restatement.  Called: SearchByProjection(F,
vpMapPoints, th), SearchByProjection(CurrentFrame, LastFrame), SearchByProjection(pKF, Scw, ...), SearchForInitialization,
SearchByBoW(KF, F), SearchByBoW(KF, KF), SearchForTriangulation, Fuse (both overloads), SearchBySim3, DescriptorDistance.  The oracle /
CUDA boundary starts at the projected point: poses are the identity and map points sit at (u, v, 1) before a unit
pinhole camera, so that the reference's own pose and projection arithmetic is exact.

Bar: bit-exact (match tables and counts).  The committed outputs (tests/golden/ref_outputs.npz, tools/gen_golden_ref.py)
run everywhere; the live tests run where the library exists.
"""
from pathlib import Path

import numpy as np
import pytest

import oracle
from oracle import QUERY_DTYPE
from pl_vi_orbslam3_b200 import synth
from pl_vi_orbslam3_b200.matchers import frame_grid, ORBmatcher
from pl_vi_orbslam3_b200.vocabulary import ORBVocabulary

GOLD = Path(__file__).resolve().parent / "golden"
R = np.load(GOLD / "ref_outputs.npz")
GRID = frame_grid(0, 752, 0, 480)
SCALES = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2))]).astype(np.float32), dtype=np.float32)
needs_ref = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def pair_features():
    f1, f2, A = synth.warp_pair(3)
    return oracle.orb_extract(f1), oracle.orb_extract(f2), A


# ---- case builders shared by the live tests and tools/gen_golden_ref.py ----------------------------------------------
def mappoint_case(r1, r2, A, seed, th):
    """Map points = the features of frame 1 projected into frame 2 by the warp; random view cosines / flags."""
    rng = np.random.RandomState(seed)
    k = r1["keypoints"]
    n = len(k)
    proj = np.stack([A[0, 0] * k["x"] + A[0, 1] * k["y"] + A[0, 2], A[1, 0] * k["x"] + A[1, 1] * k["y"] + A[1, 2]], 1).astype(np.float32)
    viewcos = np.where(rng.rand(n) < 0.5, 0.9995, 0.9).astype(np.float32)
    level = k["octave"].astype(np.int32)
    flags = ((rng.rand(n) < 0.1) * 1 + (rng.rand(n) < 0.15) * 2 + (rng.rand(n) < 0.05) * 4).astype(np.int32)
    blocked = (rng.rand(len(r2["keypoints"])) < 0.1).astype(np.uint8)
    return dict(proj=proj, viewcos=viewcos, level=level, flags=flags, blocked=blocked, th=np.float32(th))


def mappoint_queries(c):
    """Caller side of SearchByProjection(F, vpMapPoints, th) in front of the oracle / plvi_search_by_projection:
    r = RadiusByViewingCos (2.5 above 0.998, else 4.0), x th unless th == 1, x scale factor of the predicted level;
    levels level-1 .. level (src/ORBmatcher.cc:66-74)."""
    n = len(c["proj"])
    q = np.zeros(n, QUERY_DTYPE)
    q["u"], q["v"] = c["proj"][:, 0], c["proj"][:, 1]
    r = np.where(c["viewcos"] > np.float32(0.998), np.float32(2.5), np.float32(4.0)).astype(np.float32)
    if c["th"] != 1.0:
        r = (r * np.float32(c["th"])).astype(np.float32)
    q["radius"] = (r * SCALES[c["level"]]).astype(np.float32)
    q["min_level"], q["max_level"] = c["level"] - 1, c["level"]
    q["flags"] = ((c["flags"] & 1) | ((c["flags"] & 4) >> 2)) | (c["flags"] & 2)
    return q


def bow_case(r1, r2, k, L, levelsup, seed):
    v = ORBVocabulary.random_tree(k=k, L=L, seed=seed)
    rng = np.random.RandomState(seed)
    fv1 = oracle.bow_transform(v.as_oracle_dict(), r1["descriptors"], levelsup)["fv"]
    fv2 = oracle.bow_transform(v.as_oracle_dict(), r2["descriptors"], levelsup)["fv"]
    u1, u2 = rng.rand(len(r1["keypoints"])), rng.rand(len(r2["keypoints"]))
    mp1 = np.where(u1 < 0.75, 1, np.where(u1 < 0.85, 2, 0)).astype(np.uint8)   # 1 good, 2 bad, 0 none
    mp2 = np.where(u2 < 0.75, 1, np.where(u2 < 0.85, 2, 0)).astype(np.uint8)
    return fv1, fv2, mp1, mp2


def bow_queries(r1, fv1, fv2, mp1):
    """Caller side of the walk over the two FeatureVectors (src/ORBmatcher.cc:286-300): one query per keyframe feature
    of a common node, in the reference's order."""
    start2 = {int(nd): (int(fv2[1][i]), int(fv2[1][i + 1])) for i, nd in enumerate(fv2[0])}
    order, rows = [], []
    for i, nd in enumerate(fv1[0]):
        if int(nd) not in start2:
            continue
        for idx1 in fv1[2][fv1[1][i]:fv1[1][i + 1]]:
            order.append(int(idx1))
            rows.append((start2[int(nd)][0], start2[int(nd)][1], r1["keypoints"]["angle"][idx1], 0 if mp1[idx1] == 1 else 1))
    q = np.zeros(len(rows), QUERY_DTYPE)
    for j, (s, e, ang, fl) in enumerate(rows):
        q[j]["min_level"], q[j]["max_level"], q[j]["angle"], q[j]["flags"] = s, e, ang, fl
    return np.asarray(fv2[2], np.int32), q, np.array(order, np.int64)


def oracle_mappoints(r2, c, nnratio):
    return oracle.search_mappoints(r2["keypoints"], r2["descriptors"], GRID, mappoint_queries(c), c["qdesc"], 100, nnratio, c["blocked"])


def oracle_bow_kf_f(r1, r2, fv1, fv2, mp1, nnratio, check_ori):
    items, q, order = bow_queries(r1, fv1, fv2, mp1)
    n, mt = oracle.search_bow(r2["keypoints"], r2["descriptors"], items, q, r1["descriptors"][order], 50, nnratio, check_ori)
    return n, np.where(mt >= 0, order[np.maximum(mt, 0)], -1).astype(np.int32)


# ---- live ------------------------------------------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("seed,th,nnratio", [(0, 1.0, 0.8), (1, 3.0, 0.8), (2, 5.0, 0.9), (3, 15.0, 0.6)])
def test_live_reference_search_by_projection_mappoints(pair_features, seed, th, nnratio):
    r1, r2, A = pair_features
    c = mappoint_case(r1, r2, A, seed, th)
    c["qdesc"] = r1["descriptors"]
    n, mt = oracle.ref_search_mappoints(r2["keypoints"], r2["descriptors"], GRID, SCALES, c["proj"], c["viewcos"], c["level"],
                                        c["flags"], c["qdesc"], th, nnratio, c["blocked"])
    on, omt = oracle_mappoints(r2, c, nnratio)
    assert n == on and np.array_equal(mt, omt)
    assert n > 50


@needs_ref
@pytest.mark.parametrize("window,nnratio,check_ori", [(100, 0.9, True), (30, 0.9, True), (100, 0.7, False), (10, 0.9, True)])
def test_live_reference_search_for_initialization(pair_features, window, nnratio, check_ori):
    r1, r2, _ = pair_features
    k1 = r1["keypoints"]
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    n, m12, pm = oracle.ref_search_init(k1, r1["descriptors"], r2["keypoints"], r2["descriptors"], GRID, prev, window, nnratio, check_ori)
    q = ORBmatcher.init_queries(k1, prev, window)
    on, om12, oq = oracle.search_init(r2["keypoints"], r2["descriptors"], GRID, q, r1["descriptors"], 50, nnratio, check_ori)
    assert n == on and np.array_equal(m12, om12)
    assert np.array_equal(pm[:, 0], oq["u"]) and np.array_equal(pm[:, 1], oq["v"])
    if window == 100:
        assert n > 50
        # second round, as Tracking::MonocularInitialization does with the updated mvbPrevMatched
        n2, m12b, _ = oracle.ref_search_init(k1, r1["descriptors"], r2["keypoints"], r2["descriptors"], GRID, pm, window, nnratio, check_ori)
        on2, om12b, _ = oracle.search_init(r2["keypoints"], r2["descriptors"], GRID, ORBmatcher.init_queries(k1, pm, window),
                                           r1["descriptors"], 50, nnratio, check_ori)
        assert n2 == on2 and np.array_equal(m12b, om12b)


@needs_ref
@pytest.mark.parametrize("k,L,levelsup,nnratio,check_ori", [(6, 3, 2, 0.7, True), (10, 4, 2, 0.9, True), (4, 2, 1, 0.6, True),
                                                            (5, 3, 3, 0.75, False), (3, 2, 2, 0.95, True)])
def test_live_reference_search_by_bow(pair_features, k, L, levelsup, nnratio, check_ori):
    r1, r2, _ = pair_features
    fv1, fv2, mp1, mp2 = bow_case(r1, r2, k, L, levelsup, k + L)
    n, mt = oracle.ref_search_bow_kf_f(r1["keypoints"], r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], fv2, nnratio,
                                       check_ori)
    on, omt = oracle_bow_kf_f(r1, r2, fv1, fv2, mp1, nnratio, check_ori)
    assert n == on and np.array_equal(mt, omt)
    n, m = oracle.ref_search_bow_kfkf(r1["keypoints"], r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], mp2, fv2, nnratio,
                                      check_ori)
    on, om = oracle.search_bow_kfkf(r1["keypoints"], r1["descriptors"], mp1 == 1, fv1, r2["keypoints"], r2["descriptors"], mp2 == 1, fv2,
                                    nnratio, check_ori)
    assert n == on and np.array_equal(m, om)
    if L == 2:
        assert n > 10


@needs_ref
def test_live_reference_orb_descriptor_distance(pair_features):
    r1, r2, _ = pair_features
    for i in range(64):
        assert oracle.ref_orb_descriptor_distance(r1["descriptors"][i], r2["descriptors"][i]) == \
            oracle.hamming256(r1["descriptors"][i], r2["descriptors"][i])


def frame_case(r1, r2, A, seed, th):
    """SearchByProjection(CurrentFrame, LastFrame): last-frame points projected into the current frame by the warp."""
    rng = np.random.RandomState(seed)
    k = r1["keypoints"]
    uv = np.stack([A[0, 0] * k["x"] + A[0, 1] * k["y"] + A[0, 2], A[1, 0] * k["x"] + A[1, 1] * k["y"] + A[1, 2]], 1).astype(np.float32)
    flags = rng.choice([0, 0, 0, 0, 1, 2], len(k)).astype(np.int32)
    blocked = (rng.rand(len(r2["keypoints"])) < 0.15).astype(np.uint8)
    return dict(uv=uv, flags=flags, blocked=blocked, th=np.float32(th))


BOUNDS = (0.0, 752.0, 0.0, 480.0)


def frame_queries(r1, c):
    """Caller side of SearchByProjection(CurrentFrame, LastFrame, th, true) (src/ORBmatcher.cc:1999-2023): points outside
    the image bounds are dropped, radius = th * scale factor of the last octave, levels octave-1 .. octave+1."""
    k = r1["keypoints"]
    q = np.zeros(len(k), QUERY_DTYPE)
    q["u"], q["v"] = c["uv"][:, 0], c["uv"][:, 1]
    q["radius"] = (np.float32(c["th"]) * SCALES[k["octave"]]).astype(np.float32)
    q["min_level"], q["max_level"] = k["octave"] - 1, k["octave"] + 1
    q["angle"] = k["angle"]
    outside = (q["u"] < BOUNDS[0]) | (q["u"] > BOUNDS[1]) | (q["v"] < BOUNDS[2]) | (q["v"] > BOUNDS[3])
    q["flags"] = (c["flags"] & 2) | ((c["flags"] & 1) | outside)
    return q


@needs_ref
@pytest.mark.parametrize("seed,th,check_ori", [(0, 15.0, True), (1, 7.0, True), (2, 30.0, False), (3, 15.0, True)])
def test_live_reference_search_by_projection_frame(pair_features, seed, th, check_ori):
    r1, r2, A = pair_features
    c = frame_case(r1, r2, A, seed, th)
    n, mt = oracle.ref_search_frame(r2["keypoints"], r2["descriptors"], GRID, BOUNDS, SCALES, r1["keypoints"], c["uv"], c["flags"],
                                    r1["descriptors"], th, check_ori, c["blocked"])
    on, omt = oracle.search_frame(r2["keypoints"], r2["descriptors"], GRID, frame_queries(r1, c), r1["descriptors"], 100, check_ori,
                                  c["blocked"])
    assert n == on and np.array_equal(mt, omt)
    assert n > 100


def triangulation_case(r1, r2, k, L, levelsup, seed):
    v = ORBVocabulary.random_tree(k=k, L=L, seed=k + 3)
    rng = np.random.RandomState(seed)
    fv1 = oracle.bow_transform(v.as_oracle_dict(), r1["descriptors"], levelsup)["fv"]
    fv2 = oracle.bow_transform(v.as_oracle_dict(), r2["descriptors"], levelsup)["fv"]
    mp1 = (rng.rand(len(r1["keypoints"])) < 0.4).astype(np.uint8)
    mp2 = (rng.rand(len(r2["keypoints"])) < 0.4).astype(np.uint8)
    F12 = (np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32) + rng.normal(0, 2e-4, (3, 3)).astype(np.float32)) * np.float32(rng.uniform(0.5, 3))
    ep = (np.float32(rng.uniform(100, 600)), np.float32(rng.uniform(100, 400)))
    sg2 = (SCALES * SCALES * np.float32(4.0 if seed == 3 else 1.0)).astype(np.float32)
    return fv1, fv2, mp1, mp2, F12, ep, sg2


@needs_ref
@pytest.mark.parametrize("k,L,levelsup,coarse,seed", [(6, 3, 2, False, 0), (10, 4, 2, False, 1), (4, 2, 1, True, 2), (6, 3, 2, False, 3),
                                                      (3, 2, 2, False, 4)])
def test_live_reference_search_for_triangulation(pair_features, k, L, levelsup, coarse, seed):
    r1, r2, _ = pair_features
    fv1, fv2, mp1, mp2, F12, ep, sg2 = triangulation_case(r1, r2, k, L, levelsup, seed)
    n, m = oracle.ref_search_triangulation(r1["keypoints"], r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], mp2, fv2,
                                           F12, ep, SCALES, sg2, sg2, coarse, True)
    on, om = oracle.search_triangulation(r1["keypoints"], r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], mp2, fv2,
                                         F12, ep, SCALES, sg2, coarse, True)
    assert n == on and np.array_equal(m, om)
    if coarse:
        assert n > 20


INV_SIGMA2 = (np.float32(1.0) / (SCALES * SCALES)).astype(np.float32)


def kf_case(r1, r2, A, seed, th):
    """Fuse / SearchByProjection(KF, Scw): map points = frame-1 features projected into keyframe 2 by the warp (+ a
    little noise so that the chi2 gate of Fuse bites), predicted level = their octave (+1 sometimes)."""
    rng = np.random.RandomState(seed)
    k = r1["keypoints"]
    uv = np.stack([A[0, 0] * k["x"] + A[0, 1] * k["y"] + A[0, 2], A[1, 0] * k["x"] + A[1, 1] * k["y"] + A[1, 2]], 1)
    uv = (uv + rng.normal(0, 1.0, uv.shape)).astype(np.float32)
    level = np.minimum(k["octave"] + (rng.rand(len(k)) < 0.3), 7).astype(np.int32)
    flags = (rng.rand(len(k)) < 0.1).astype(np.int32)
    matched_in = (rng.rand(len(r2["keypoints"])) < 0.15).astype(np.uint8)
    return dict(uv=uv, level=level, flags=flags, matched_in=matched_in, th=th)


def kf_queries(c):
    """Caller side of the keyframe projection searches (src/ORBmatcher.cc:1473-1505, 520-548): IsInImage (x >= min, x < max),
    radius = th * scale factor of the predicted level, levels level-1 .. level."""
    q = np.zeros(len(c["uv"]), QUERY_DTYPE)
    q["u"], q["v"] = c["uv"][:, 0], c["uv"][:, 1]
    q["radius"] = (np.float32(c["th"]) * SCALES[c["level"]]).astype(np.float32)
    q["min_level"], q["max_level"] = c["level"] - 1, c["level"]
    inside = (q["u"] >= BOUNDS[0]) & (q["u"] < BOUNDS[1]) & (q["v"] >= BOUNDS[2]) & (q["v"] < BOUNDS[3])
    q["flags"] = (c["flags"] & 1) | ~inside
    return q


@needs_ref
@pytest.mark.parametrize("seed,th,sim3", [(0, 3.0, False), (1, 4.0, False), (2, 3.0, True), (3, 7.5, True)])
def test_live_reference_fuse(pair_features, seed, th, sim3):
    r1, r2, A = pair_features
    c = kf_case(r1, r2, A, seed, th)
    n, bi = oracle.ref_fuse(r2["keypoints"], r2["descriptors"], GRID, BOUNDS, SCALES, INV_SIGMA2, c["uv"], c["level"], c["flags"],
                            r1["descriptors"], th, sim3)
    on, obi, _ = oracle.search_in_radius(r2["keypoints"], r2["descriptors"], GRID, kf_queries(c), r1["descriptors"], INV_SIGMA2,
                                         0.0 if sim3 else 5.99, 50)
    assert n == on and np.array_equal(bi, obi)
    assert n > 50


@needs_ref
@pytest.mark.parametrize("seed,th,ratio", [(0, 8, 1.0), (1, 15, 1.5), (2, 30, 1.0), (3, 15, 0.8)])
def test_live_reference_search_by_projection_keyframe(pair_features, seed, th, ratio):
    """SearchByProjection(pKF, Scw, ...) CLAIMS features while it iterates (vpMatched[idx] != NULL is skipped,
    src/ORBmatcher.cc:554-555, 575): it is the sequential search of search_frame without the rotation check, not the
    independent per-point search of Fuse."""
    r1, r2, A = pair_features
    c = kf_case(r1, r2, A, seed, float(th))
    n, mt = oracle.ref_search_by_projection_kf(r2["keypoints"], r2["descriptors"], GRID, BOUNDS, SCALES, c["uv"], c["level"], c["flags"],
                                               r1["descriptors"], th, ratio, c["matched_in"])
    on, omt = oracle.search_frame(r2["keypoints"], r2["descriptors"], GRID, kf_queries(c), r1["descriptors"], int(np.floor(50 * ratio)),
                                  False, c["matched_in"])
    assert n == on and np.array_equal(mt, omt)
    assert n > 100


def sim3_case(r1, r2, A, seed):
    """SearchBySim3: the map points of keyframe 1 projected into keyframe 2 by the warp, those of keyframe 2 into
    keyframe 1 by its inverse (+ noise); some features without / with bad map points, some already matched."""
    rng = np.random.RandomState(seed)
    Ai = np.linalg.inv(np.vstack([A, [0, 0, 1]]))[:2]
    out = []
    for r, M in ((r1, A), (r2, Ai)):
        k = r["keypoints"]
        uv = np.stack([M[0, 0] * k["x"] + M[0, 1] * k["y"] + M[0, 2], M[1, 0] * k["x"] + M[1, 1] * k["y"] + M[1, 2]], 1)
        uv = (uv + rng.normal(0, 0.7, uv.shape)).astype(np.float32)
        level = np.minimum(k["octave"] + (rng.rand(len(k)) < 0.3), 7).astype(np.int32)
        flags = ((rng.rand(len(k)) < 0.15) * 1 + (rng.rand(len(k)) < 0.05) * 2).astype(np.int32)
        out.append((uv, level, flags))
    out[0][2][:] |= ((rng.rand(len(r1["keypoints"])) < 0.05) * 4).astype(np.int32)
    return out


def sim3_queries(uv, level, flags, th):
    """Caller side of one direction of SearchBySim3 (src/ORBmatcher.cc:1783-1825)."""
    q = kf_queries(dict(uv=uv, level=level, flags=np.zeros(len(uv), np.int32), th=th))
    q["flags"] |= (flags != 0)
    return q


def sim3_mutual(bi12, bi21):
    """The agreement test of src/ORBmatcher.cc:1944-1957 on the two best-index arrays."""
    m = np.full(len(bi12), -1, np.int32)
    ok = bi12 >= 0
    ok[ok] = bi21[bi12[ok]] == np.nonzero(ok)[0]
    m[ok] = bi12[ok]
    return int(ok.sum()), m


@needs_ref
@pytest.mark.parametrize("seed,th", [(0, 7.5), (1, 3.0), (2, 15.0)])
def test_live_reference_search_by_sim3(pair_features, seed, th):
    r1, r2, A = pair_features
    (uv1, l1, f1), (uv2, l2, f2) = sim3_case(r1, r2, A, seed)
    n, m = oracle.ref_search_by_sim3(r1["keypoints"], r1["descriptors"], uv1, l1, f1, r2["keypoints"], r2["descriptors"], uv2, l2, f2,
                                     GRID, BOUNDS, SCALES, th)
    _, bi12, _ = oracle.search_in_radius(r2["keypoints"], r2["descriptors"], GRID, sim3_queries(uv1, l1, f1, th), r1["descriptors"],
                                         INV_SIGMA2, 0.0, 100)
    _, bi21, _ = oracle.search_in_radius(r1["keypoints"], r1["descriptors"], GRID, sim3_queries(uv2, l2, f2, th), r2["descriptors"],
                                         INV_SIGMA2, 0.0, 100)
    on, om = sim3_mutual(bi12, bi21)
    assert n == on and np.array_equal(m, om)
    assert n > 100


def reloc_queries(r1, c):
    """Caller side of the relocalisation overload (src/ORBmatcher.cc:2205-2232): bounds test as in the frame-to-frame
    search, radius = th * scale factor of the predicted level, levels pred-1 .. pred+1; every claim blocks."""
    k = r1["keypoints"]
    q = np.zeros(len(k), QUERY_DTYPE)
    q["u"], q["v"] = c["uv"][:, 0], c["uv"][:, 1]
    q["radius"] = (np.float32(c["th"]) * SCALES[c["level"]]).astype(np.float32)
    q["min_level"], q["max_level"] = c["level"] - 1, c["level"] + 1
    q["angle"] = k["angle"]
    outside = (q["u"] < BOUNDS[0]) | (q["u"] > BOUNDS[1]) | (q["v"] < BOUNDS[2]) | (q["v"] > BOUNDS[3])
    q["flags"] = (c["flags"] != 0) | outside
    return q


def reloc_case(r1, r2, A, seed, th):
    c = kf_case(r1, r2, A, seed, th)
    rng = np.random.RandomState(seed + 50)
    c["flags"] = ((rng.rand(len(c["uv"])) < 0.1) * 1 + (rng.rand(len(c["uv"])) < 0.05) * 2 + (rng.rand(len(c["uv"])) < 0.1) * 4).astype(np.int32)
    return c


@needs_ref
@pytest.mark.parametrize("seed,th,orb_dist,check_ori", [(0, 10.0, 100, True), (1, 3.0, 64, True), (2, 10.0, 100, False)])
def test_live_reference_search_by_projection_relocalisation(pair_features, seed, th, orb_dist, check_ori):
    r1, r2, A = pair_features
    c = reloc_case(r1, r2, A, seed, th)
    n, mt = oracle.ref_search_reloc(r2["keypoints"], r2["descriptors"], GRID, BOUNDS, SCALES, r1["keypoints"], c["uv"], c["level"],
                                    c["flags"], r1["descriptors"], th, orb_dist, check_ori, c["matched_in"])
    on, omt = oracle.search_frame(r2["keypoints"], r2["descriptors"], GRID, reloc_queries(r1, c), r1["descriptors"], orb_dist, check_ori,
                                  c["matched_in"])
    assert n == on and np.array_equal(mt, omt)
    assert n > 100


@needs_ref
def test_live_reference_tracking_searches_on_the_reference_frame_class(pair_features):
    """The three tracking searches once more, now with ORBmatcher.cc compiled against the reference's OWN Frame class
    (libplvi_ref_frame.so): AssignFeaturesToGrid and GetFeaturesInArea underneath are the reference's code as well."""
    r1, r2, A = pair_features
    c = mappoint_case(r1, r2, A, 2, 5.0)
    c["qdesc"] = r1["descriptors"]
    n, mt = oracle.ref_search_mappoints(r2["keypoints"], r2["descriptors"], GRID, SCALES, c["proj"], c["viewcos"], c["level"], c["flags"],
                                        c["qdesc"], 5.0, 0.9, c["blocked"], real_frame_bounds=BOUNDS)
    on, omt = oracle_mappoints(r2, c, 0.9)
    assert n == on and np.array_equal(mt, omt) and n > 50
    c = frame_case(r1, r2, A, 0, 15.0)
    n, mt = oracle.ref_search_frame(r2["keypoints"], r2["descriptors"], GRID, BOUNDS, SCALES, r1["keypoints"], c["uv"], c["flags"],
                                    r1["descriptors"], 15.0, True, c["blocked"], real_frame=True)
    on, omt = oracle.search_frame(r2["keypoints"], r2["descriptors"], GRID, frame_queries(r1, c), r1["descriptors"], 100, True, c["blocked"])
    assert n == on and np.array_equal(mt, omt) and n > 100
    k1 = r1["keypoints"]
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    n, m12, pm = oracle.ref_search_init(k1, r1["descriptors"], r2["keypoints"], r2["descriptors"], GRID, prev, 100, 0.9, True,
                                        real_frame_bounds=BOUNDS)
    on, om12, oq = oracle.search_init(r2["keypoints"], r2["descriptors"], GRID, ORBmatcher.init_queries(k1, prev, 100), r1["descriptors"],
                                      50, 0.9, True)
    assert n == on and np.array_equal(m12, om12) and np.array_equal(pm[:, 0], oq["u"]) and n > 50


@needs_ref
def test_live_reference_keyframe_searches_on_the_reference_keyframe_class(pair_features):
    """Fuse (both overloads), SearchByProjection(pKF, Scw, ...) and SearchByBoW(pKF, F) with ORBmatcher.cc compiled
    against the reference's OWN KeyFrame / Frame classes (libplvi_ref_frame.so; the keyframe is built by its own
    constructor, AddMapPoint / GetMapPoint / GetFeaturesInArea / IsInImage are the reference's code)."""
    r1, r2, A = pair_features
    for seed, th, sim3 in ((0, 3.0, False), (3, 7.5, True)):
        c = kf_case(r1, r2, A, seed, th)
        n, bi = oracle.ref_fuse_real(r2["keypoints"], r2["descriptors"], BOUNDS, SCALES, INV_SIGMA2, c["uv"], c["level"], c["flags"],
                                     r1["descriptors"], th, sim3)
        on, obi, _ = oracle.search_in_radius(r2["keypoints"], r2["descriptors"], GRID, kf_queries(c), r1["descriptors"], INV_SIGMA2,
                                             0.0 if sim3 else 5.99, 50)
        assert n == on and np.array_equal(bi, obi) and n > 50
    c = kf_case(r1, r2, A, 1, 15.0)
    n, mt = oracle.ref_search_by_projection_kf_real(r2["keypoints"], r2["descriptors"], BOUNDS, SCALES, c["uv"], c["level"], c["flags"],
                                                    r1["descriptors"], 15, 1.5, c["matched_in"])
    on, omt = oracle.search_frame(r2["keypoints"], r2["descriptors"], GRID, kf_queries(c), r1["descriptors"], 75, False, c["matched_in"])
    assert n == on and np.array_equal(mt, omt) and n > 100
    fv1, fv2, mp1, _ = bow_case(r1, r2, 6, 3, 2, 9)
    n, mt = oracle.ref_search_bow_kf_f_real(r1["keypoints"], r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], fv2, BOUNDS,
                                            0.7, True)
    on, omt = oracle_bow_kf_f(r1, r2, fv1, fv2, mp1, 0.7, True)
    assert n == on and np.array_equal(mt, omt) and n > 50


# ---- committed outputs of the reference (run everywhere) -------------------------------------------------------------
def test_oracle_equals_reference_orbmatcher_outputs(pair_features):
    r1, r2, A = pair_features
    c = mappoint_case(r1, r2, A, 1, 3.0)
    c["qdesc"] = r1["descriptors"]
    n, mt = oracle_mappoints(r2, c, 0.8)
    assert n == int(R["orbmatch/mappoints_n"]) and np.array_equal(mt, R["orbmatch/mappoints"])
    k1 = r1["keypoints"]
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    n, m12, q = oracle.search_init(r2["keypoints"], r2["descriptors"], GRID, ORBmatcher.init_queries(k1, prev, 100), r1["descriptors"],
                                   50, 0.9, True)
    assert n == int(R["orbmatch/init_n"]) and np.array_equal(m12, R["orbmatch/init"])
    assert np.array_equal(np.stack([q["u"], q["v"]], 1), R["orbmatch/init_prev"])
    fv1, fv2, mp1, mp2 = bow_case(r1, r2, 6, 3, 2, 9)
    n, mt = oracle_bow_kf_f(r1, r2, fv1, fv2, mp1, 0.7, True)
    assert n == int(R["orbmatch/bow_n"]) and np.array_equal(mt, R["orbmatch/bow"])
    n, m = oracle.search_bow_kfkf(k1, r1["descriptors"], mp1 == 1, fv1, r2["keypoints"], r2["descriptors"], mp2 == 1, fv2, 0.8, True)
    assert n == int(R["orbmatch/bowkf_n"]) and np.array_equal(m, R["orbmatch/bowkf"])
    c = frame_case(r1, r2, A, 1, 7.0)
    n, mt = oracle.search_frame(r2["keypoints"], r2["descriptors"], GRID, frame_queries(r1, c), r1["descriptors"], 100, True, c["blocked"])
    assert n == int(R["orbmatch/frame_n"]) and np.array_equal(mt, R["orbmatch/frame"])
    fv1, fv2, mp1, mp2, F12, ep, sg2 = triangulation_case(r1, r2, 6, 3, 2, 3)
    n, m = oracle.search_triangulation(k1, r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], mp2, fv2, F12, ep, SCALES, sg2,
                                       False, True)
    assert n == int(R["orbmatch/tri_n"]) and np.array_equal(m, R["orbmatch/tri"])
    c = kf_case(r1, r2, A, 1, 4.0)
    for key, chi2 in (("fuse", 5.99), ("fuse_sim3", 0.0)):
        n, bi, _ = oracle.search_in_radius(r2["keypoints"], r2["descriptors"], GRID, kf_queries(c), r1["descriptors"], INV_SIGMA2, chi2, 50)
        assert n == int(R[f"orbmatch/{key}_n"]) and np.array_equal(bi, R[f"orbmatch/{key}"])
    c = kf_case(r1, r2, A, 1, 15.0)
    n, mt = oracle.search_frame(r2["keypoints"], r2["descriptors"], GRID, kf_queries(c), r1["descriptors"], 75, False, c["matched_in"])
    assert n == int(R["orbmatch/kf_n"]) and np.array_equal(mt, R["orbmatch/kf"])
    (uv1, l1, f1), (uv2, l2, f2) = sim3_case(r1, r2, A, 1)
    _, bi12, _ = oracle.search_in_radius(r2["keypoints"], r2["descriptors"], GRID, sim3_queries(uv1, l1, f1, 7.5), r1["descriptors"],
                                         INV_SIGMA2, 0.0, 100)
    _, bi21, _ = oracle.search_in_radius(k1, r1["descriptors"], GRID, sim3_queries(uv2, l2, f2, 7.5), r2["descriptors"], INV_SIGMA2, 0.0, 100)
    n, m = sim3_mutual(bi12, bi21)
    assert n == int(R["orbmatch/sim3_n"]) and np.array_equal(m, R["orbmatch/sim3"])


# ---- MapPoint::ComputeDistinctiveDescriptors: the reference's own MapPoint.cc + MapPoint.h over stand-in KeyFrame / Map --
def distinctive_case(seed, M=48, max_obs=40):
    """M map points with 1..max_obs observed descriptors each: noisy copies of a base descriptor (ties are frequent: the
    medians are small integers); every fifth point has identical observations."""
    rng = np.random.RandomState(seed)
    counts = rng.randint(1, max_obs + 1, M).astype(np.int32)
    desc = np.zeros((M, max_obs, 32), np.uint8)
    for p in range(M):
        base = rng.randint(0, 256, 32).astype(np.uint8)
        for i in range(counts[p]):
            d = base.copy()
            for bit in rng.permutation(256)[:rng.randint(0, 60)]:
                d[bit // 8] ^= 1 << (bit % 8)
            desc[p, i] = d
        if p % 5 == 0:
            desc[p, :counts[p]] = desc[p, 0]
    return desc, counts


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_live_reference_distinctive_descriptors(seed):
    desc, counts = distinctive_case(seed)
    rng = np.random.RandomState(seed + 100)
    for p in range(len(counts)):
        d = desc[p, :counts[p]]
        r = oracle.ref_distinctive_descriptor(d)
        assert np.array_equal(r, d[oracle.distinctive_descriptor(d)])
        bad = (rng.rand(len(d)) < 0.3).astype(np.uint8)       # observations from bad keyframes are left out
        r = oracle.ref_distinctive_descriptor(d, bad)
        good = d[bad == 0]
        assert (r is None and len(good) == 0) or np.array_equal(r, good[oracle.distinctive_descriptor(good)])


@needs_ref
@pytest.mark.parametrize("seed", [3, 4])
def test_live_reference_mapline_distinctive_descriptors(seed):
    """MapLine::ComputeDistinctiveDescriptors (src/MapLine.cc:264-329, MapLine.cc + MapLine.h compiled unmodified): the same
    least-median rule over LBD descriptors with ORBmatcher::DescriptorDistance (:305)."""
    desc, counts = distinctive_case(seed)
    rng = np.random.RandomState(seed + 100)
    for p in range(len(counts)):
        d = desc[p, :counts[p]]
        r = oracle.ref_mapline_distinctive_descriptor(d)
        assert np.array_equal(r, d[oracle.distinctive_descriptor(d)])
        assert np.array_equal(r, oracle.ref_distinctive_descriptor(d))
        bad = (rng.rand(len(d)) < 0.3).astype(np.uint8)
        r = oracle.ref_mapline_distinctive_descriptor(d, bad)
        good = d[bad == 0]
        assert (r is None and len(good) == 0) or np.array_equal(r, good[oracle.distinctive_descriptor(good)])


def test_oracle_equals_reference_distinctive_outputs():
    desc, counts = distinctive_case(7)
    for p in range(len(counts)):
        d = desc[p, :counts[p]]
        assert np.array_equal(d[oracle.distinctive_descriptor(d)], R["mappoint/distinctive"][p])


# ---- the reference's real frames (tests/golden/frame_data2_{1,3}.npz = /root/reference/data2/color/{1,3}.png, gray) ----
def real_pair():
    a = np.load(GOLD / "frame_data2_1.npz")["img"]
    b = np.load(GOLD / "frame_data2_3.npz")["img"]
    return a, b


@needs_ref
def test_live_reference_matchers_on_real_frames():
    """Initialisation matching (points and lines) between two frames of the reference's own sequence: features from the
    reference's extractors, matches from the reference's matchers, against the oracle end to end."""
    a, b = real_pair()
    h, w = a.shape
    grid = frame_grid(0, w, 0, h)
    ra, rb = oracle.ref_orb_extract(a, nfeatures=1000), oracle.ref_orb_extract(b, nfeatures=1000)
    oa, ob = oracle.orb_extract(a, nfeatures=1000), oracle.orb_extract(b, nfeatures=1000)
    ka = ra["keypoints"]
    prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    n, m12, pm = oracle.ref_search_init(ka, ra["descriptors"], rb["keypoints"], rb["descriptors"], grid, prev, 100, 0.9, True)
    on, om12, oq = oracle.search_init(ob["keypoints"], ob["descriptors"], grid, ORBmatcher.init_queries(oa["keypoints"], prev, 100),
                                      oa["descriptors"], 50, 0.9, True)
    assert n == on and np.array_equal(m12, om12) and n > 30
    la, lb = oracle.ref_line_extract(a), oracle.ref_line_extract(b)
    pa, pb = oracle.line_extract(a), oracle.line_extract(b)
    n, m = oracle.ref_line_match(la["descriptors"], lb["descriptors"], 0.75, "match")
    on, om = oracle.line_match(pa["descriptors"], pb["descriptors"], 0.75)
    assert n == on and np.array_equal(m, om)
    n, m = oracle.ref_line_match_mad(la["descriptors"], lb["descriptors"], 0.5)
    on, om, _ = oracle.line_match_mad(pa["descriptors"], pb["descriptors"], 0.5)
    assert n == on and np.array_equal(m, om) and n > 20
