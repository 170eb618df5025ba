"""CPU: the oracle restatement against THE REFERENCE'S OWN CODE.

oracle/_ref/libplvi_ref.so = the unmodified reference sources (ORBextractor.cc, LSD/lsd.cpp, LineExtractor.cc,
LSDDetector_custom.cpp, binary_descriptor_custom.cpp) compiled where they lie, against the OpenCV/Eigen stand-in of
oracle/cvmini (OpenCV primitives = the scalar models pinned against cv2).  tests/golden/ref_outputs.npz holds its
outputs (tools/gen_golden_ref.py), so the first test runs everywhere; the live tests run where the library exists
(build container: built from /root/reference; GPU box: the prebuilt .so travels with the snapshot).

Bar: everything bit-exact.  That includes the results of the host libm functions the reference calls on float
arguments (cosf / sinf in rBRIEF, in LSD's region sums and in LBD; atan2f in KeyLine::angle): the oracle restates
glibc's algorithms operation by operation (oracle/oracle_common.h, namespace glibcm; exhaustively equal to this
image's libm, tools/scan_libm.c).
"""
from pathlib import Path

import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import synth

GOLD = Path(__file__).resolve().parent / "golden"
R = np.load(GOLD / "ref_outputs.npz")
CASES = [c.split("|") for c in R["cases"]]


def frame(name):
    if name.startswith("synth_"):
        p = name.split("_")
        return synth.frame_euroc(int(p[1]), *(int(v) for v in p[2:4])) if len(p) > 2 else synth.frame_euroc(int(p[1]))
    return np.load(GOLD / f"frame_{name}.npz")["img"]


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b).max(initial=0)


def check_orb(got, kp, desc, mono):
    assert len(got["keypoints"]) == len(kp)
    assert got["mono_index"] == int(mono)
    for f in kp.dtype.names:
        assert np.array_equal(got["keypoints"][f], kp[f]), f
    assert np.array_equal(got["descriptors"], desc)


def check_lines(got, kl, desc, eq):
    assert len(got["keylines"]) == len(kl)
    for f in kl.dtype.names:
        assert np.array_equal(got["keylines"][f], kl[f]), f
    assert np.array_equal(got["descriptors"], desc)
    assert np.array_equal(got["line_eq"], eq)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_equals_reference_outputs(case):
    name, nf, lap0, lap1, lnf = case[0], int(case[1]), int(case[2]), int(case[3]), int(case[4])
    img = frame(name)
    check_orb(oracle.orb_extract(img, nfeatures=nf, lapping=(lap0, lap1)), R[f"{name}/orb_kp"], R[f"{name}/orb_desc"],
              R[f"{name}/orb_mono"])
    assert np.array_equal(oracle.lsd(img), R[f"{name}/lsd_raw"])
    check_lines(oracle.line_extract(img, lsd_nfeatures=lnf), R[f"{name}/line_kl"], R[f"{name}/line_desc"],
                R[f"{name}/line_eq"])


def test_clipped_line_iterator_count_is_covered():
    """synth_0 / synth_10 with all lines kept contain a segment whose rounded endpoint falls outside the octave
    image: cv::LineIterator clips it (numOfPixels is one less than max(|dx|, |dy|) + 1)."""
    hit = 0
    for name in ("synth_0", "synth_10"):
        k = R[f"{name}/line_kl"]
        ax, ay = np.rint(k["sPointInOctaveX"]).astype(int), np.rint(k["sPointInOctaveY"]).astype(int)
        bx, by = np.rint(k["ePointInOctaveX"]).astype(int), np.rint(k["ePointInOctaveY"]).astype(int)
        hit += int((np.maximum(abs(bx - ax), abs(by - ay)) + 1 != k["numOfPixels"]).sum())
    assert hit >= 2


def test_clip_line_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    import ctypes as C
    rng = np.random.RandomState(1)
    L = oracle.lib()
    for _ in range(20000):
        w, h = int(rng.randint(1, 60)), int(rng.randint(1, 60))
        p = [int(v) for v in rng.randint(-40, 100, size=4)]
        r, a, b = cv2.clipLine((0, 0, w, h), (p[0], p[1]), (p[2], p[3]))
        c = [C.c_int(v) for v in p]
        r2 = L.plvio_clip_line(w, h, *[C.byref(v) for v in c])
        assert bool(r) == bool(r2)
        if r:
            assert (a, b) == ((c[0].value, c[1].value), (c[2].value, c[3].value))


needs_ref = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libplvi_ref.so not built")


@needs_ref
@pytest.mark.parametrize("seed,w,h,nf", [(1, 752, 480, 1000), (2, 752, 480, 2000), (4, 640, 480, 2000),
                                         (5, 701, 413, 1000), (6, 1280, 720, 1000)])
def test_live_reference_orb(seed, w, h, nf):
    img = synth.frame_euroc(seed, w, h)
    r = oracle.ref_orb_extract(img, nfeatures=nf, debug=True)
    o = oracle.orb_extract(img, nfeatures=nf, debug=True)
    check_orb(o, r["keypoints"], r["descriptors"], r["mono_index"])
    for a, b in zip(o["pyramid"], r["pyramid"]):
        assert np.array_equal(a, b)


@needs_ref
def test_live_reference_orb_edge_images():
    rng = np.random.RandomState(5)
    for img in (rng.randint(0, 256, (480, 752)).astype(np.uint8), np.full((480, 752), 128, np.uint8)):
        r = oracle.ref_orb_extract(img)
        check_orb(oracle.orb_extract(img), r["keypoints"], r["descriptors"], r["mono_index"])


@needs_ref
@pytest.mark.parametrize("seed,w,h,lnf,levels", [(1, 752, 480, 200, 2), (2, 752, 480, 0, 2), (7, 752, 480, 0, 2),
                                                 (4, 640, 480, 100, 2), (5, 701, 413, 200, 1), (6, 1280, 720, 200, 2)])
def test_live_reference_lines(seed, w, h, lnf, levels):
    img = synth.frame_euroc(seed, w, h)
    r = oracle.ref_line_extract(img, lsd_nfeatures=lnf, nlevels=levels)
    check_lines(oracle.line_extract(img, lsd_nfeatures=lnf, nlevels=levels), r["keylines"], r["descriptors"], r["line_eq"])
    assert np.array_equal(oracle.lsd(img), oracle.ref_lsd(img))


@needs_ref
@pytest.mark.parametrize("lsd_scale", [1.0, 0.6, 0.5, 0.9, 0.33])
def test_live_reference_lines_lsd_scales(lsd_scale):
    """lsd_scale 1.0 (Examples/Stereo-Line/UMA_ueye.yaml; no blur, no resize) and the smaller settings the reference's
    yaml comments recommend: other Gaussian sizes (9 / 11 / 15 taps) through the reference's own flsd."""
    img = synth.frame_euroc(9)
    r = oracle.ref_line_extract(img, lsd_nfeatures=150, lsd_scale=lsd_scale)
    check_lines(oracle.line_extract(img, lsd_nfeatures=150, lsd_scale=lsd_scale), r["keylines"], r["descriptors"], r["line_eq"])
    assert np.array_equal(oracle.lsd(img, lsd_scale), oracle.ref_lsd(img, lsd_scale))
    assert len(r["keylines"]) > 50


@needs_ref
@pytest.mark.parametrize("refine", [1, 2])
@pytest.mark.parametrize("seed,w,h,lsd_scale", [(1, 752, 480, 0.8), (4, 640, 480, 0.8), (6, 752, 480, 0.6), (8, 752, 480, 1.0)])
def test_live_reference_lines_lsd_refine(refine, seed, w, h, lsd_scale):
    """lsd_refine 1 / 2 (refine, reduce_region_radius, rect_improve, rect_nfa, nfa: src/LSD/lsd.cpp:784-1134, with the
    vendored code's integer steps and its "(double(n) + 1)" term): the oracle's restatement against the reference's own
    lsd.cpp, raw segments and the whole extractor."""
    img = synth.frame_euroc(seed, w, h)
    a, b = oracle.lsd(img, lsd_scale, refine=refine), oracle.ref_lsd(img, lsd_scale, refine)
    assert len(a) > 200 and np.array_equal(a, b)
    assert len(a) != len(oracle.lsd(img, lsd_scale))          # refine changes the segment set
    r = oracle.ref_line_extract(img, lsd_refine=refine, lsd_scale=lsd_scale)
    check_lines(oracle.line_extract(img, lsd_refine=refine, lsd_scale=lsd_scale), r["keylines"], r["descriptors"], r["line_eq"])


@needs_ref
def test_live_reference_lines_edge_images():
    rng = np.random.RandomState(5)
    noise = rng.randint(0, 256, (480, 752)).astype(np.uint8)
    r = oracle.ref_line_extract(noise)
    check_lines(oracle.line_extract(noise), r["keylines"], r["descriptors"], r["line_eq"])
    flat = np.full((480, 752), 128, np.uint8)
    assert len(oracle.ref_line_extract(flat)["keylines"]) == 0 and len(oracle.line_extract(flat)["keylines"]) == 0


def test_restated_libm_equals_host_libm():
    """glibcm::sinf / cosf / atan2f (oracle_common.h) against the libm the oracle library is linked with, on 4M
    sampled arguments (tools/scan_libm.cpp does the exhaustive scan).  Only meaningful on glibc 2.28 .. 2.40,
    x86-64 with FMA3 -- the build image and the GPU box; skipped elsewhere."""
    import platform
    ver = tuple(int(v) for v in platform.libc_ver()[1].split(".")[:2]) if platform.libc_ver()[0] == "glibc" else (0, 0)
    flags = open("/proc/cpuinfo").read() if Path("/proc/cpuinfo").exists() else ""
    if not ((2, 28) <= ver <= (2, 40)) or " fma " not in flags or platform.machine() != "x86_64":
        pytest.skip(f"host libm {ver} / CPU is not the reference platform model")
    L = oracle.lib()
    rng = np.random.RandomState(3)
    n = 1 << 21
    x = np.concatenate([rng.uniform(-7, 7, n), rng.uniform(-119, 119, n // 2),
                        rng.standard_normal(n // 2) * 1e-3]).astype(np.float32)
    s1, c1, s2, c2 = (np.empty_like(x) for _ in range(4))
    L.plvio_glibc_sincosf(oracle._p(x), len(x), oracle._p(s1), oracle._p(c1))
    L.plvio_host_sincosf(oracle._p(x), len(x), oracle._p(s2), oracle._p(c2))
    assert np.array_equal(s1.view(np.uint32), s2.view(np.uint32)) and np.array_equal(c1.view(np.uint32), c2.view(np.uint32))
    y = (rng.uniform(-800, 800, n) * rng.choice([1.0, 0.125, 1e-3], n)).astype(np.float32)
    xx = (rng.uniform(-800, 800, n) * rng.choice([1.0, 0.125, 1e-3], n)).astype(np.float32)
    y[:64] = 0.0
    xx[64:128] = 0.0
    xx[128:192] = 1.0
    r1, r2 = np.empty_like(y), np.empty_like(y)
    L.plvio_glibc_atan2f(oracle._p(y), oracle._p(xx), n, oracle._p(r1))
    L.plvio_host_atan2f(oracle._p(y), oracle._p(xx), n, oracle._p(r2))
    assert np.array_equal(r1.view(np.uint32), r2.view(np.uint32))


def _write_voc(v, path, trailing_newline=False):
    v.save_text(path)
    if not trailing_newline:
        txt = Path(path).read_text().rstrip("\n")
        Path(path).write_text(txt)


@needs_ref
@pytest.mark.parametrize("k,L,stop,early,scoring,weighting,levelsup", [(10, 4, 0.0, 0.0, 0, 0, 4), (8, 3, 0.1, 0.0, 0, 0, 2),
                                                                       (10, 4, 0.05, 0.1, 0, 0, 3), (6, 5, 0.0, 0.3, 1, 1, 4),
                                                                       (9, 4, 0.02, 0.05, 5, 0, 2), (10, 3, 0.0, 0.0, 0, 3, 1)])
def test_live_reference_dbow2_transform(tmp_path, k, L, stop, early, scoring, weighting, levelsup):
    """Frame::ComputeBoW: the oracle against the reference's own DBoW2 sources (TemplatedVocabulary::loadFromTextFile
    reads the ORBvoc text format written by vocabulary.save_text, then transform): BowVector ids and values (doubles,
    bit for bit) and FeatureVector.  The file is written without a trailing newline: the reference's loader turns a
    final empty line into an extra child of the root with an uninitialised descriptor (TemplatedVocabulary.h:1378-1392),
    and for a full tree that extra node outgrows the reserve() and leaves m_words dangling."""
    from pl_vi_orbslam3_b200.vocabulary import ORBVocabulary
    v = ORBVocabulary.random_tree(k=k, L=L, seed=k + L, stop_fraction=stop, early_leaf_fraction=early, scoring=scoring,
                                  weighting=weighting)
    path = tmp_path / "voc.txt"
    _write_voc(v, path)
    rng = np.random.RandomState(L)
    desc = rng.randint(0, 256, (1500, 32)).astype(np.uint8)
    desc[:200] = v.desc[rng.randint(1, len(v.desc), 200)]          # exact hits on node descriptors: distance ties at 0
    a = oracle.bow_transform(v.as_oracle_dict(), desc, levelsup)
    b = oracle.ref_bow_transform(path, desc, levelsup)
    assert b["words"] == v.words
    assert np.array_equal(a["bow"][0], b["bow"][0]) and np.array_equal(a["bow"][1], b["bow"][1])
    for x, y in zip(a["fv"], b["fv"]):
        assert np.array_equal(x, y)
    # the round trip of the text format through load_text is the same tree
    w = ORBVocabulary.load_text(path)
    assert np.array_equal(w.parent, v.parent) and np.array_equal(w.desc[1:], v.desc[1:]) and np.array_equal(w.weight[1:], v.weight[1:])


def test_oracle_equals_reference_dbow2_outputs():
    """Committed outputs of the reference's DBoW2 (tools/gen_golden_ref.py) for a seeded synthetic vocabulary."""
    from pl_vi_orbslam3_b200.vocabulary import ORBVocabulary
    v = ORBVocabulary.random_tree(k=8, L=4, seed=21, stop_fraction=0.03, early_leaf_fraction=0.1)
    desc = oracle.orb_extract(frame("synth_0"))["descriptors"]
    a = oracle.bow_transform(v.as_oracle_dict(), desc, 4)
    assert np.array_equal(a["bow"][0], R["bow/words"]) and np.array_equal(a["bow"][1], R["bow/values"])
    assert np.array_equal(a["fv"][0], R["bow/fv_nodes"]) and np.array_equal(a["fv"][1], R["bow/fv_start"])
    assert np.array_equal(a["fv"][2], R["bow/fv_features"])


# ---- LineMatcher.cpp of the reference, compiled unmodified against the stand-in SLAM classes (cvmini/slam_mock.h) ----
def _line_desc_pair(rng, n1, n2, p):
    """n2 noisy copies (bit-flip probability p) of random rows of n1 random 256-bit descriptors."""
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    noise = (rng.random((n2, 32, 8)) < p).astype(np.uint8)
    d2 = d1[rng.integers(0, n1, n2)] ^ np.packbits(noise, axis=2).reshape(n2, 32)
    return d1, d2


@needs_ref
@pytest.mark.parametrize("seed", range(12))
def test_live_reference_line_matcher(seed):
    """matchNNR, both match() overloads, SerachForInitialize, SearchForTriangulation(KF, KF), distance and
    DescriptorDistance of the reference itself against the oracle restatement (ties included: p = 0 gives
    duplicate descriptors, i.e. equal distances and equal NN12 differences)."""
    rng = np.random.default_rng(seed)
    n1, n2 = (int(x) for x in rng.integers(2, 300, 2))
    p = [0.0, 0.02, 0.1, 0.3][seed % 4]
    d1, d2 = _line_desc_pair(rng, n1, n2, p)
    for nnr in (0.75, 0.9, 1.0):
        n, m = oracle.ref_line_match(d1, d2, nnr, "nnr")
        on, om = oracle.match_nnr(d1, d2, nnr)
        assert n == on and np.array_equal(m, om)
        for variant in ("match", "maplines"):
            n, m = oracle.ref_line_match(d1, d2, nnr, variant)
            on, om = oracle.line_match(d1, d2, nnr)
            assert n == on and np.array_equal(m, om)
    n, m = oracle.ref_line_match_mad(d1, d2, 0.5)
    on, om, _ = oracle.line_match_mad(d1, d2, 0.5)
    assert n == on and np.array_equal(m, om)
    h1, h2 = (rng.random(n1) < 0.3).astype(np.uint8), (rng.random(n2) < 0.3).astype(np.uint8)
    for a, b in ((None, None), (h1, h2)):
        n, m = oracle.ref_line_match_mad(d1, d2, 0.1, a, b)
        on, om, _ = oracle.line_match_mad(d1, d2, 0.1, a, b)
        assert n == on and np.array_equal(m, om)
    for i in range(min(n1, n2, 16)):
        assert oracle.ref_line_distance(d1[i], d2[i], 0) == oracle.hamming256(d1[i], d2[i])
        assert oracle.ref_line_distance(d1[i], d2[i], 1) == oracle.hamming256(d1[i], d2[i], shift25=True)


def test_oracle_equals_reference_line_matcher_outputs():
    """Committed outputs of the reference's LineMatcher (tools/gen_golden_ref.py) on the LBD descriptors of two
    synthetic frames."""
    d1 = oracle.line_extract(frame("synth_0"))["descriptors"]
    d2 = oracle.line_extract(frame("synth_1"))["descriptors"]
    n, m = oracle.line_match(d1, d2, 0.75)
    assert n == int(R["linematch/match_n"]) and np.array_equal(m, R["linematch/match"])
    n, m, _ = oracle.line_match_mad(d1, d2, 0.5)
    assert n == int(R["linematch/init_n"]) and np.array_equal(m, R["linematch/init"])
    h1, h2 = R["linematch/has1"], R["linematch/has2"]
    n, m, _ = oracle.line_match_mad(d1, d2, 0.1, h1, h2)
    assert n == int(R["linematch/tri_n"]) and np.array_equal(m, R["linematch/tri"])
    kl, desc, sf, q, qd, bad = line_fuse_case(1, 8.0)
    n, bi, _ = oracle.line_fuse_search(kl, desc, q, qd, line_fuse_flags(q, bad), 50)
    assert n == int(R["linematch/fuse_n"]) and np.array_equal(bi, R["linematch/fuse"])


def line_fuse_case(seed, th):
    """Keylines + LBD descriptors of a synthetic frame and map lines = perturbed copies of them, projected endpoints
    given directly (the reference runs with an identity pose and a unit pinhole, see oracle.ref_line_fuse)."""
    rng = np.random.RandomState(seed)
    r = oracle.line_extract(synth.frame_euroc(40 + seed))
    kl, desc = r["keylines"], r["descriptors"]
    nq = 300
    src = rng.randint(0, len(kl), nq)
    sf = np.array([1, 2, 4, 8], np.float32)
    q = np.zeros((nq, 6), np.float32)
    jit = rng.uniform(-6, 6, (nq, 4)).astype(np.float32)
    q[:, 0] = kl["startPointX"][src] + jit[:, 0]
    q[:, 1] = kl["startPointY"][src] + jit[:, 1]
    q[:, 2] = kl["endPointX"][src] + jit[:, 2]
    q[:, 3] = kl["endPointY"][src] + jit[:, 3]
    q[::17, 2] = q[::17, 0]                      # vertical projections: x1 == x2 -> division by zero in the slope
    q[:, 5] = kl["octave"][src] + rng.randint(0, 2, nq)
    q[:, 4] = np.float32(th) * sf[q[:, 5].astype(np.int32)]   # radius = th * mvScaleFactors[nPredictedLevel]
    qd = desc[src].copy()
    qd ^= ((rng.rand(nq, 32) < 0.08) * rng.randint(0, 256, (nq, 32))).astype(np.uint8)
    bad = (rng.rand(nq) < 0.1).astype(np.uint8)
    return kl, desc, sf, q, qd, bad


LINE_BOUNDS = (0.0, 752.0, 0.0, 480.0)


def line_fuse_flags(q, bad):
    """Caller side of LineMatcher::Fuse in front of the search (src/LineMatcher.cpp:414-429): both projected endpoints
    must lie inside [mnMinX, mnMaxX] x [mnMinY, mnMaxY]."""
    b = LINE_BOUNDS
    out = (q[:, 0] < b[0]) | (q[:, 0] > b[1]) | (q[:, 1] < b[2]) | (q[:, 1] > b[3]) | \
          (q[:, 2] < b[0]) | (q[:, 2] > b[1]) | (q[:, 3] < b[2]) | (q[:, 3] > b[3])
    return (bad.astype(bool) | out).astype(np.uint8)


@needs_ref
@pytest.mark.parametrize("seed,th", [(0, 3.0), (1, 8.0), (2, 20.0), (3, 60.0)])
def test_live_reference_line_fuse(seed, th):
    kl, desc, sf, q, qd, bad = line_fuse_case(seed, th)
    n, bi = oracle.ref_line_fuse(kl, desc, LINE_BOUNDS, sf, q, qd, bad, th)
    on, obi, _ = oracle.line_fuse_search(kl, desc, q, qd, line_fuse_flags(q, bad), 50)
    assert n == on and np.array_equal(bi, obi)
    if th >= 8.0:
        assert n > 20
