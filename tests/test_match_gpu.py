"""GPU parity: CUDA Hamming searches (through the C ABI) vs the CPU oracle.

Bar: bit-exact -- distances, match lists and match counts are integer work.
"""
import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import FrameView, LineMatcher, ORBmatcher, frame_grid, synth
from pl_vi_orbslam3_b200.capi import QUERY_DTYPE

pytestmark = pytest.mark.gpu

GRID = frame_grid(0, 752, 0, 480)
SCALES = np.float32(1.2) ** np.arange(8, dtype=np.float32)


@pytest.fixture(scope="module")
def pair_features():
    """C3: ORB features (from the oracle) of a synthetic frame and its affine warp."""
    f1, f2, A = synth.warp_pair(3)
    r1, r2 = oracle.orb_extract(f1), oracle.orb_extract(f2)
    return r1, r2, A


def _proj_queries(r1, A, th=15.0, lo=-1, hi=+1):
    k = r1["keypoints"]
    q = np.zeros(len(k), QUERY_DTYPE)
    q["u"] = (A[0, 0] * k["x"] + A[0, 1] * k["y"] + A[0, 2]).astype(np.float32)
    q["v"] = (A[1, 0] * k["x"] + A[1, 1] * k["y"] + A[1, 2]).astype(np.float32)
    q["radius"] = np.float32(th) * SCALES[k["octave"]]
    q["min_level"] = k["octave"] + lo
    q["max_level"] = k["octave"] + hi
    q["angle"] = k["angle"]
    return q


@pytest.fixture(scope="module")
def om(gpu):
    m = ORBmatcher(0.9, True, max_pairs=8, max_train=6000, max_query=6000)
    yield m
    m.close()


@pytest.fixture(scope="module")
def lm(gpu):
    m = LineMatcher(max_pairs=64, max_train=512, max_query=512)
    yield m
    m.close()


def test_descriptor_distance_bit_exact(om, lm):
    rng = np.random.RandomState(0)
    a = rng.randint(0, 256, (4096, 32)).astype(np.uint8)
    b = rng.randint(0, 256, (4096, 32)).astype(np.uint8)
    b[:100] = a[:100]
    b[100:200] = ~a[100:200]
    ref = np.array([oracle.hamming256(a[i], b[i]) for i in range(len(a))])
    assert np.array_equal(om.DescriptorDistance(a, b), ref)
    assert np.array_equal(lm.distance(a, b), ref)
    ref25 = np.array([oracle.hamming256(a[i], b[i], shift25=True) for i in range(len(a))])
    assert np.array_equal(lm.DescriptorDistance(a, b), ref25)
    assert ref[:100].max() == 0 and ref[100:200].min() == 256


@pytest.mark.parametrize("th", [15.0, 30.0])
@pytest.mark.parametrize("ori", [True, False])
def test_search_by_projection_frame(om, pair_features, th, ori):
    r1, r2, A = pair_features
    q = _proj_queries(r1, A, th)
    om.mbCheckOrientation = ori
    F = FrameView(r2["keypoints"], r2["descriptors"], GRID)
    n, mt, mq = om.SearchByProjection(F, q, r1["descriptors"])
    rn, rmt = oracle.search_frame(r2["keypoints"], r2["descriptors"], GRID, q, r1["descriptors"], 100, ori)
    om.mbCheckOrientation = True
    assert n == rn and n > 300
    assert np.array_equal(mt, rmt)


def test_search_frame_blocked_and_flags(om, pair_features):
    r1, r2, A = pair_features
    q = _proj_queries(r1, A, 15.0)
    rng = np.random.RandomState(1)
    q["flags"] = rng.choice([0, 0, 0, 1, 2], len(q)).astype(np.int32)
    blocked = (rng.rand(len(r2["keypoints"])) < 0.2).astype(np.uint8)
    F = FrameView(r2["keypoints"], r2["descriptors"], GRID, blocked)
    n, mt, mq = om.SearchByProjection(F, q, r1["descriptors"])
    rn, rmt = oracle.search_frame(r2["keypoints"], r2["descriptors"], GRID, q, r1["descriptors"], 100, True, blocked)
    assert n == rn
    assert np.array_equal(mt, rmt)


@pytest.mark.parametrize("nnratio", [0.8, 0.6])
def test_search_by_projection_mappoints(om, pair_features, nnratio):
    r1, r2, A = pair_features
    q = _proj_queries(r1, A, 4.0 * 2.5, lo=-1, hi=0)
    om.mfNNratio = nnratio
    F = FrameView(r2["keypoints"], r2["descriptors"], GRID)
    n, mt, mq = om.SearchByProjection(F, q, r1["descriptors"], mappoints=True)
    om.mfNNratio = 0.9
    rn, rmt = oracle.search_mappoints(r2["keypoints"], r2["descriptors"], GRID, q, r1["descriptors"], 100, nnratio)
    assert n == rn and n > 100
    assert np.array_equal(mt, rmt)


@pytest.mark.parametrize("window", [100, 30])
def test_search_for_initialization(om, pair_features, window):
    r1, r2, A = pair_features
    k1 = r1["keypoints"]
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    F2 = FrameView(r2["keypoints"], r2["descriptors"], GRID)
    n, m12, prev_out = om.SearchForInitialization(k1, r1["descriptors"], F2, prev, window)
    q = ORBmatcher.init_queries(k1, prev, window)
    rn, rm12, rq = oracle.search_init(r2["keypoints"], r2["descriptors"], GRID, q, r1["descriptors"], 50, 0.9, True)
    assert n == rn and n > 30
    assert np.array_equal(m12, rm12)
    assert np.array_equal(prev_out[:, 0], rq["u"]) and np.array_equal(prev_out[:, 1], rq["v"])


def _tie_heavy_case(seed, n=1500, nq=1200):
    """Lattice keypoints with descriptors from a tiny codebook: distance ties and
    overlapping windows everywhere -> exercises tie-breaks, claims and steals."""
    rng = np.random.RandomState(seed)
    keys = np.zeros(n, oracle.KEYPOINT_DTYPE)
    keys["x"] = rng.randint(20, 732, n).astype(np.float32)
    keys["y"] = rng.randint(20, 460, n).astype(np.float32)
    keys["octave"] = rng.randint(0, 3, n)
    keys["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    book = rng.randint(0, 256, (6, 32)).astype(np.uint8)
    desc = book[rng.randint(0, 6, n)].copy()
    flip = rng.rand(n) < 0.5
    desc[flip, 0] ^= 1
    q = np.zeros(nq, QUERY_DTYPE)
    q["u"] = rng.uniform(0, 752, nq).astype(np.float32)
    q["v"] = rng.uniform(0, 480, nq).astype(np.float32)
    q["radius"] = rng.choice([12.0, 25.0, 60.0], nq).astype(np.float32)
    q["min_level"] = rng.choice([-1, 0, 1], nq)
    q["max_level"] = q["min_level"] + rng.choice([0, 1, 2], nq)
    q["angle"] = rng.uniform(0, 360, nq).astype(np.float32)
    qd = book[rng.randint(0, 6, nq)].copy()
    qd[rng.rand(nq) < 0.3, 1] ^= 3
    return keys, desc, q, qd


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_search_tie_heavy_all_modes(om, seed):
    keys, desc, q, qd = _tie_heavy_case(seed)
    F = FrameView(keys, desc, GRID)
    n, mt, _ = om.SearchByProjection(F, q, qd)
    rn, rmt = oracle.search_frame(keys, desc, GRID, q, qd, 100, True)
    assert n == rn and np.array_equal(mt, rmt)
    om.mfNNratio = 1.0
    n, mt, _ = om.SearchByProjection(F, q, qd, mappoints=True)
    om.mfNNratio = 0.9
    rn, rmt = oracle.search_mappoints(keys, desc, GRID, q, qd, 100, 1.0)
    assert n == rn and np.array_equal(mt, rmt)
    q2 = q.copy()
    q2["min_level"], q2["max_level"] = 0, 0
    mt, mq, nm, qs = om.search_batch(2, [F], [q2], [qd], 50)
    rn, rm12, rq = oracle.search_init(keys, desc, GRID, q2, qd, 50, 0.9, True)
    assert nm[0] == rn and np.array_equal(mq[0], rm12)


def test_search_batch_of_pairs(om):
    cases = [_tie_heavy_case(10 + i, n=700 + 50 * i, nq=500 + 30 * i) for i in range(6)]
    frames = [FrameView(c[0], c[1], GRID) for c in cases]
    mt, mq, nm, _ = om.search_batch(0, frames, [c[2] for c in cases], [c[3] for c in cases], 100)
    for i, c in enumerate(cases):
        rn, rmt = oracle.search_frame(c[0], c[1], GRID, c[2], c[3], 100, True)
        assert nm[i] == rn and np.array_equal(mt[i], rmt)


def test_search_empty_sets(om):
    keys, desc, q, qd = _tie_heavy_case(5, n=50, nq=40)
    F0 = FrameView(keys[:0], desc[:0], GRID)
    n, mt, mq = om.SearchByProjection(F0, q, qd)
    assert n == 0 and len(mt) == 0 and (mq == -1).all()
    n, mt, mq = om.SearchByProjection(FrameView(keys, desc, GRID), q[:0], qd[:0])
    assert n == 0 and (mt == -1).all()


def _line_descs(rng, n, dup_of=None):
    d = rng.randint(0, 256, (n, 32)).astype(np.uint8)
    if dup_of is not None and n and len(dup_of):
        idx = rng.randint(0, len(dup_of), n)
        noisy = dup_of[idx].copy()
        noisy[np.arange(n), rng.randint(0, 32, n)] ^= (1 << rng.randint(0, 8, n)).astype(np.uint8)
        take = rng.rand(n) < 0.6
        d[take] = noisy[take]
    return d


@pytest.mark.parametrize("n1,n2", [(200, 200), (200, 137), (1, 5), (5, 1), (0, 7), (7, 0), (2, 2), (400, 380)])
def test_line_match_bit_exact(lm, n1, n2):
    rng = np.random.RandomState(n1 * 1000 + n2)
    d1 = _line_descs(rng, n1)
    d2 = _line_descs(rng, n2, d1)
    if n2 > 10:
        d2[3] = d2[7]  # exact duplicate train rows: tie -> lowest index
    for nnr in (0.9, 0.75):
        n, m = lm.match(d1, d2, nnr)
        rn, rm = oracle.line_match(d1, d2, nnr)
        assert n == rn and np.array_equal(m, rm)
        n, m = lm.matchNNR(d1, d2, nnr)
        rn, rm = oracle.match_nnr(d1, d2, nnr)
        assert n == rn and np.array_equal(m, rm)


def test_line_match_batch(lm):
    rng = np.random.RandomState(77)
    pairs = []
    for i in range(48):
        d1 = _line_descs(rng, rng.randint(0, 220))
        pairs.append((d1, _line_descs(rng, rng.randint(0, 220), d1)))
    ms, nm = lm.match_batch(pairs, 0.9)
    for i, (a, b) in enumerate(pairs):
        rn, rm = oracle.line_match(a, b, 0.9)
        assert nm[i] == rn and np.array_equal(ms[i], rm)


def _bow_case(r1, r2, seed, nodes=40):
    """Synthetic DBoW2-like feature vectors: a 'vocabulary node' per feature (from two descriptor
    bytes, so that similar descriptors tend to share a node), frame features grouped by node."""
    rng = np.random.RandomState(seed)
    node1 = (r1["descriptors"][:, 0].astype(np.int32) * 3 + r1["descriptors"][:, 5]) % nodes
    node2 = (r2["descriptors"][:, 0].astype(np.int32) * 3 + r2["descriptors"][:, 5]) % nodes
    items, start, end = [], {}, {}
    for nd in range(nodes):
        idx = np.nonzero(node2 == nd)[0]
        start[nd], end[nd] = len(items), len(items) + len(idx)
        items.extend(idx.tolist())
    order = [i for nd in range(nodes) for i in np.nonzero(node1 == nd)[0] if end[nd] > start[nd]]
    q = np.zeros(len(order), QUERY_DTYPE)
    k1 = r1["keypoints"]
    for j, i in enumerate(order):
        q[j]["min_level"], q[j]["max_level"] = start[node1[i]], end[node1[i]]
        q[j]["angle"] = k1["angle"][i]
        q[j]["flags"] = int(rng.rand() < 0.1)
    return np.array(items, np.int32), q, r1["descriptors"][order]


@pytest.mark.parametrize("nodes,nnratio", [(40, 0.7), (6, 0.9), (300, 0.7)])
def test_search_by_bow(om, pair_features, nodes, nnratio):
    r1, r2, _ = pair_features
    items, q, qd = _bow_case(r1, r2, nodes, nodes)
    om.mfNNratio = nnratio
    F = FrameView(r2["keypoints"], r2["descriptors"], GRID)
    n, mt, mq = om.SearchByBoW(F, items, q, qd)
    om.mfNNratio = 0.9
    rn, rmt = oracle.search_bow(r2["keypoints"], r2["descriptors"], items, q, qd, 50, nnratio, True)
    assert n == rn and np.array_equal(mt, rmt)
    if nodes == 6:
        assert n > 20


# ---- LineMatcher::SerachForInitialize / SearchForTriangulation (kNN-2 + MAD threshold) and
# MapPoint::ComputeDistinctiveDescriptors
def _line_sets(seed, n1, n2):
    rng = np.random.RandomState(seed)
    a = rng.randint(0, 256, (n1, 32)).astype(np.uint8)
    b = rng.randint(0, 256, (n2, 32)).astype(np.uint8)
    k = min(n1, n2) * 2 // 3
    sel = rng.permutation(n2)[:k]
    b[sel] = a[:k]
    flips = rng.randint(0, 40, k)                       # noisy copies: 0..40 flipped bits
    for j, f in zip(sel, flips):
        bits = rng.permutation(256)[:f]
        for bit in bits:
            b[j, bit // 8] ^= 1 << (bit % 8)
    return a, b


@pytest.mark.parametrize("n1,n2,seed", [(200, 200, 0), (199, 150, 1), (64, 201, 2), (2, 2, 3), (7, 1, 4)])
def test_line_search_for_initialize_and_triangulation(gpu, n1, n2, seed):
    from pl_vi_orbslam3_b200 import LineMatcher
    a, b = _line_sets(seed, n1, n2)
    lm = LineMatcher(max_pairs=2, max_train=256, max_query=256)
    try:
        rn, rm, rmad = oracle.line_match_mad(a, b, 0.5)
        n, pairs = lm.SerachForInitialize(a, b)
        assert n == rn and pairs == [(i, int(t)) for i, t in enumerate(rm) if t >= 0]
        rng = np.random.RandomState(seed + 50)
        h1, h2 = (rng.rand(n1) < 0.2).astype(np.uint8), (rng.rand(n2) < 0.2).astype(np.uint8)
        rn, rm, _ = oracle.line_match_mad(a, b, 0.1, h1, h2)
        n, pairs = lm.SearchForTriangulation(a, b, h1, h2)
        assert n == rn and pairs == [(i, int(t)) for i, t in enumerate(rm) if t >= 0]
        m, nm, mad = lm._match_mad_batch([(a, b), (b, a)], 0.5)
        assert np.array_equal(mad[0], rmad) and np.array_equal(mad[1], oracle.line_match_mad(b, a, 0.5)[2])
        assert np.array_equal(m[1], oracle.line_match_mad(b, a, 0.5)[1])
    finally:
        lm.close()


def test_line_mad_on_extracted_lines(gpu):
    from pl_vi_orbslam3_b200 import LineMatcher, Lineextractor
    f1, f2, _ = synth.warp_pair(5)
    l = Lineextractor(200, 0, 0.8, 2, 2.0, 0)
    lm = LineMatcher(max_pairs=1, max_train=256, max_query=256)
    try:
        _, d1, _ = l(f1)
        _, d2, _ = l(f2)
        n, pairs = lm.SerachForInitialize(d1, d2)
        rn, rm, _ = oracle.line_match_mad(d1, d2, 0.5)
        assert n == rn > 20 and pairs == [(i, int(t)) for i, t in enumerate(rm) if t >= 0]
    finally:
        l.close()
        lm.close()


def test_compute_distinctive_descriptors(gpu):
    import torch
    from pl_vi_orbslam3_b200.matchers import compute_distinctive_descriptors
    rng = np.random.RandomState(8)
    M, cap = 300, 70
    counts = rng.randint(0, cap + 1, M).astype(np.int32)
    counts[:4] = (0, 1, 2, cap)
    desc = np.zeros((M, cap, 32), np.uint8)
    for p in range(M):
        base = rng.randint(0, 256, 32).astype(np.uint8)
        for i in range(counts[p]):
            d = base.copy()
            for bit in rng.permutation(256)[:rng.randint(0, 60)]:
                d[bit // 8] ^= 1 << (bit % 8)
            desc[p, i] = d
    desc[5, :counts[5]] = desc[5, 0]                      # all identical: index 0 wins
    idx, best = compute_distinctive_descriptors(torch.from_numpy(desc).cuda(), torch.from_numpy(counts).cuda())
    idx, best = idx.cpu().numpy(), best.cpu().numpy()
    for p in range(M):
        r = oracle.distinctive_descriptor(desc[p, :counts[p]])
        assert idx[p] == r, p
        if r >= 0:
            assert np.array_equal(best[p], desc[p, r])


# ---- ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) through plvi_search_by_bow_kf, with real
# FeatureVectors from the DBoW2 transform oracle on a synthetic vocabulary
@pytest.mark.parametrize("k,L,levelsup,nnratio", [(6, 3, 2, 0.8), (10, 4, 2, 0.9), (4, 2, 1, 0.6)])
def test_search_by_bow_keyframes(om, pair_features, k, L, levelsup, nnratio):
    from pl_vi_orbslam3_b200.vocabulary import ORBVocabulary
    r1, r2, _ = pair_features
    v = ORBVocabulary.random_tree(k=k, L=L, seed=k)
    rng = np.random.RandomState(L)
    fv1 = oracle.bow_transform(v.as_oracle_dict(), r1["descriptors"], levelsup)["fv"]
    fv2 = oracle.bow_transform(v.as_oracle_dict(), r2["descriptors"], levelsup)["fv"]
    mp1 = (rng.rand(len(r1["keypoints"])) < 0.8).astype(np.uint8)
    mp2 = (rng.rand(len(r2["keypoints"])) < 0.8).astype(np.uint8)
    # caller side of the reference's walk over the two FeatureVectors: KF2 groups + one query per KF1 feature of a common node
    items, start2 = fv2[2], {int(nd): (int(fv2[1][i]), int(fv2[1][i + 1])) for i, nd in enumerate(fv2[0])}
    order, q = [], []
    for i, nd in enumerate(fv1[0]):
        if int(nd) not in start2:
            continue
        for idx1 in fv1[2][fv1[1][i]:fv1[1][i + 1]]:
            order.append(int(idx1))
            q.append((start2[int(nd)][0], start2[int(nd)][1], r1["keypoints"]["angle"][idx1], 0 if mp1[idx1] else 1))
    qs = np.zeros(len(q), QUERY_DTYPE)
    for j, (s, e, ang, fl) in enumerate(q):
        qs[j]["min_level"], qs[j]["max_level"], qs[j]["angle"], qs[j]["flags"] = s, e, ang, fl
    om.mfNNratio = nnratio
    KF2 = FrameView(r2["keypoints"], r2["descriptors"], GRID)
    n, mq = om.SearchByBoW_KF(KF2, mp2, items, qs, r1["descriptors"][order])
    om.mfNNratio = 0.9
    got = np.full(len(r1["keypoints"]), -1, np.int32)
    for j, idx1 in enumerate(order):
        if mq[j] >= 0:
            got[idx1] = mq[j]
    rn, rm = oracle.search_bow_kfkf(r1["keypoints"], r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], mp2, fv2,
                                    nnratio, True)
    assert n == rn and np.array_equal(got, rm)
    if L <= 3:
        assert n > 10


# ---- ORBmatcher::SearchForTriangulation (mono pinhole path: Hamming + epipole + epipolar-line tests)
@pytest.mark.parametrize("k,L,levelsup,coarse,seed", [(6, 3, 2, False, 0), (10, 4, 2, False, 1), (4, 2, 1, True, 2), (6, 3, 2, False, 3)])
def test_search_for_triangulation(om, pair_features, k, L, levelsup, coarse, seed):
    from pl_vi_orbslam3_b200.vocabulary import ORBVocabulary
    r1, r2, _ = pair_features
    v = ORBVocabulary.random_tree(k=k, L=L, seed=k + 3)
    rng = np.random.RandomState(seed)
    fv1 = oracle.bow_transform(v.as_oracle_dict(), r1["descriptors"], levelsup)["fv"]
    fv2 = oracle.bow_transform(v.as_oracle_dict(), r2["descriptors"], levelsup)["fv"]
    mp1 = (rng.rand(len(r1["keypoints"])) < 0.4).astype(np.uint8)
    mp2 = (rng.rand(len(r2["keypoints"])) < 0.4).astype(np.uint8)
    # a fundamental matrix of a (noisy) sideways motion: epipolar lines roughly horizontal; random scale
    F12 = (np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32) + rng.normal(0, 2e-4, (3, 3)).astype(np.float32)) * np.float32(rng.uniform(0.5, 3))
    ep = (np.float32(rng.uniform(100, 600)), np.float32(rng.uniform(100, 400)))
    sf2 = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    sg2 = (sf2 * sf2 * np.float32(4.0 if seed == 3 else 1.0)).astype(np.float32)
    items, start2 = fv2[2], {int(nd): (int(fv2[1][i]), int(fv2[1][i + 1])) for i, nd in enumerate(fv2[0])}
    order, q = [], []
    k1 = r1["keypoints"]
    for i, nd in enumerate(fv1[0]):
        if int(nd) not in start2:
            continue
        for idx1 in fv1[2][fv1[1][i]:fv1[1][i + 1]]:
            if mp1[idx1]:
                continue                                   # only features without a map point are queried
            order.append(int(idx1))
            q.append((k1["x"][idx1], k1["y"][idx1], start2[int(nd)][0], start2[int(nd)][1], k1["angle"][idx1]))
    qs = np.zeros(len(q), QUERY_DTYPE)
    for j, (u, vv, s, e, ang) in enumerate(q):
        qs[j]["u"], qs[j]["v"], qs[j]["min_level"], qs[j]["max_level"], qs[j]["angle"] = u, vv, s, e, ang
    KF2 = FrameView(r2["keypoints"], r2["descriptors"], GRID)
    n, mq = om.SearchForTriangulation(KF2, mp2, items, qs, r1["descriptors"][order], F12, ep, sf2, sg2, coarse)
    got = np.full(len(k1), -1, np.int32)
    for j, idx1 in enumerate(order):
        got[idx1] = mq[j]
    rn, rm = oracle.search_triangulation(k1, r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], mp2, fv2, F12, ep,
                                         sf2, sg2, coarse, True)
    assert n == rn and np.array_equal(got, rm)
    if coarse:
        assert n > 20


INV_SIGMA2 = (np.float32(1.0) / (SCALES * SCALES)).astype(np.float32)


@pytest.mark.parametrize("chi2,th_dist", [(5.99, 50), (0.0, 50), (0.0, 100)])
def test_search_in_radius_fuse_and_sim3_patterns(om, pair_features, chi2, th_dist):
    """Fuse (chi2 gate, TH_LOW), Fuse(Scw) / SearchByProjection(KF, Scw) (no gate, TH_LOW) and one direction of
    SearchBySim3 (no gate, TH_HIGH) on the warped pair: projected points of frame 1 searched in frame 2."""
    r1, r2, A = pair_features
    q = _proj_queries(r1, A, th=3.0, lo=-1, hi=0)
    q["flags"][::7] = 1          # points rejected before the search (bad / already in the keyframe / out of image)
    F = FrameView(r2["keypoints"], r2["descriptors"], GRID)
    n, bi, bd = om.SearchInRadius(F, q, r1["descriptors"], INV_SIGMA2, chi2, th_dist)
    rn, rbi, rbd = oracle.search_in_radius(r2["keypoints"], r2["descriptors"], GRID, q, r1["descriptors"], INV_SIGMA2, chi2, th_dist)
    assert n == rn and n > 100
    assert np.array_equal(bi, rbi) and np.array_equal(bd, rbd)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_search_in_radius_tie_heavy(om, seed):
    keys, desc, q, qd = _tie_heavy_case(seed)
    q["max_level"] = q["min_level"] + 1
    F = FrameView(keys, desc, GRID)
    for chi2, th in ((5.99, 50), (0.0, 100), (50.0, 256)):
        n, bi, bd = om.SearchInRadius(F, q, qd, INV_SIGMA2, chi2, th)
        rn, rbi, rbd = oracle.search_in_radius(keys, desc, GRID, q, qd, INV_SIGMA2, chi2, th)
        assert n == rn and np.array_equal(bi, rbi) and np.array_equal(bd, rbd)
    # SearchBySim3's mutual agreement (src/ORBmatcher.cc:1944-1957) from the two directions
    n12, m12, _ = om.SearchInRadius(F, q, qd, INV_SIGMA2, 0.0, 100)
    assert n12 == int((m12 >= 0).sum())


def test_search_in_radius_empty(om):
    keys, desc, q, qd = _tie_heavy_case(3, n=50, nq=40)
    n, bi, bd = om.SearchInRadius(FrameView(keys[:0], desc[:0], GRID), q, qd, INV_SIGMA2)
    assert n == 0 and (bi == -1).all() and (bd == 256).all()
    n, bi, bd = om.SearchInRadius(FrameView(keys, desc, GRID), q[:0], qd[:0], INV_SIGMA2)
    assert n == 0 and len(bi) == 0


def _line_fuse_case(seed):
    """Keylines + LBD descriptors of a synthetic frame (oracle) and map lines = perturbed copies of them."""
    rng = np.random.RandomState(seed)
    r = oracle.line_extract(synth.frame_euroc(40 + seed))
    kl, desc = r["keylines"], r["descriptors"]
    nq = 300
    src = rng.randint(0, len(kl), nq)
    q = np.zeros((nq, 6), np.float32)
    jit = rng.uniform(-6, 6, (nq, 4)).astype(np.float32)
    q[:, 0] = kl["startPointX"][src] + jit[:, 0]
    q[:, 1] = kl["startPointY"][src] + jit[:, 1]
    q[:, 2] = kl["endPointX"][src] + jit[:, 2]
    q[:, 3] = kl["endPointY"][src] + jit[:, 3]
    q[:, 4] = rng.choice([3.0, 8.0, 20.0, 60.0], nq).astype(np.float32) * np.float32(1.2)
    q[:, 5] = kl["octave"][src] + rng.randint(0, 2, nq)
    q[::17, 2] = q[::17, 0]                      # vertical projections: x1 == x2 -> division by zero in the slope
    qd = desc[src].copy()
    flip = rng.rand(nq, 32) < 0.08
    qd ^= (flip * rng.randint(0, 256, (nq, 32))).astype(np.uint8)
    flags = (rng.rand(nq) < 0.1).astype(np.uint8)
    return kl, desc, q, qd, flags


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_line_fuse_search(lm, seed):
    kl, desc, q, qd, flags = _line_fuse_case(seed)
    n, bi, bd = lm.FuseSearch(kl, desc, q, qd, flags)
    rn, rbi, rbd = oracle.line_fuse_search(kl, desc, q, qd, flags, 50)
    assert n == rn and n > 20
    assert np.array_equal(bi, rbi) and np.array_equal(bd, rbd)
    n0, bi0, bd0 = lm.FuseSearch(kl[:0], desc[:0], q, qd)
    assert n0 == 0 and (bi0 == -1).all()
