"""Stereo line search: Frame::ComputeStereoMatches_Lines' grid fill + LineMatcher::matchGrid
(src/Frame.cc:1421-1448, src/LineIterator.cpp, src/gridStructure.cpp, src/LineMatcher.cpp:191-272).

CPU: the oracle's bitmap restatement against the reference's own GridStructure / LineIterator / matchGrid
(oracle/_ref/libplvi_ref.so, compiled unmodified; live where the library exists) and against committed reference
outputs (tests/golden/ref_match_grid.npz, tools/gen_golden_match_grid.py).  GPU: plvi_line_match_grid through the
C ABI against the oracle, and against the live reference.  Bar: bit-exact match tables and counts.
"""
from pathlib import Path

import numpy as np
import pytest

import oracle

GOLD = Path(__file__).resolve().parent / "golden" / "ref_match_grid.npz"
needs_ref = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built")
W, H = 752, 480
INV_W, INV_H = 64 / float(W), 48 / float(H)   # src/Frame.cc:208-209
SIZES = [(200, 200), (150, 90), (7, 300), (64, 2), (1, 1), (0, 5), (5, 0), (33, 33)]


def stereo_line_case(seed, n1, n2, w=W, h=H):
    """Right lines uniform over the image; every left line = a right line shifted by a disparity of 0..100 px
    (so that it falls into the 8-cell window to the left of ... the right line's cells) plus endpoint jitter,
    some swapped end points, some lines inside one grid cell (NaN direction), some out of the image; the left
    descriptor = the right one with a few bits flipped, with duplicates so that pre-emption and ties occur."""
    rng = np.random.RandomState(seed)
    lim = np.array([w, h, w, h], np.float32)
    seg2 = (rng.uniform(0, 1, (n2, 4)) * lim).astype(np.float32)
    d2 = rng.randint(0, 256, (n2, 32)).astype(np.uint8)
    if n2 >= 4:
        d2[1] = d2[0]                                   # tie on the best distance
        seg2[1] = seg2[0] + np.float32(1.5)
        seg2[2, 2:] = seg2[2, :2] + np.float32(2.0)     # short line: a single cell
        seg2[3] = np.float32([-30, 20, 900, 500])       # leaves the image: cells outside the grid are dropped
    if n2 == 0:
        seg1 = (rng.uniform(0, 1, (n1, 4)) * lim).astype(np.float32)
        return seg1, rng.randint(0, 256, (n1, 32)).astype(np.uint8), seg2, d2
    idx = rng.randint(0, n2, n1)
    seg1 = seg2[idx] + rng.normal(0, 2.0, (n1, 4)).astype(np.float32)
    disp = rng.uniform(0, 100, n1).astype(np.float32)
    seg1[:, 0] += disp
    seg1[:, 2] += disp
    swap = rng.rand(n1) < 0.2
    seg1[swap] = seg1[swap][:, [2, 3, 0, 1]]
    seg1 = np.clip(seg1, 0, lim - 1).astype(np.float32)
    d1 = d2[idx].copy()
    nflip = rng.randint(0, 40, n1)
    for i in range(n1):
        bits = rng.randint(0, 256, nflip[i])
        for b in bits:
            d1[i, b >> 3] ^= np.uint8(1 << (b & 7))
    if n1 >= 3:
        seg1[2, 2:] = seg1[2, :2]                       # zero-length left line: NaN direction passes the |cos| test
    return seg1, d1, seg2, d2


def test_oracle_match_grid_against_committed_reference_outputs():
    R = np.load(GOLD)
    total = 0
    for k, (n1, n2) in enumerate(SIZES):
        s1, d1, s2, d2 = stereo_line_case(100 + k, n1, n2)
        n, m = oracle.line_match_grid(s1, d1, s2, d2, INV_W, INV_H)
        assert n == int(R[f"n_{k}"]) and np.array_equal(m, R[f"m12_{k}"]), k
        assert n == int((m >= 0).sum())
        total += n
    assert total > 150


@needs_ref
def test_oracle_match_grid_against_live_reference():
    total = 0
    for seed in range(48):
        n1, n2 = SIZES[seed % len(SIZES)]
        s1, d1, s2, d2 = stereo_line_case(seed, n1, n2)
        for win in ((7, 0, 2, 2), (3, 3, 0, 1)):
            a = oracle.line_match_grid(s1, d1, s2, d2, INV_W, INV_H, window=win)
            b = oracle.ref_line_match_grid(s1, d1, s2, d2, INV_W, INV_H, window=win)
            assert a[0] == b[0] and np.array_equal(a[1], b[1]), (seed, win)
            total += a[0]
    assert total > 1000


def test_oracle_match_grid_other_grid_shape():
    s1, d1, s2, d2 = stereo_line_case(7, 120, 140, 1280, 720)
    n, m = oracle.line_match_grid(s1, d1, s2, d2, 32 / 1280.0, 24 / 720.0, grid_rows=24, grid_cols=32)
    assert n == int((m >= 0).sum()) and n > 10
    if oracle.ref_available():
        rn, rm = oracle.ref_line_match_grid(s1, d1, s2, d2, 32 / 1280.0, 24 / 720.0, grid_rows=24, grid_cols=32)
        assert rn == n and np.array_equal(rm, m)


@pytest.mark.gpu
def test_match_grid_gpu_against_oracle(gpu):
    from pl_vi_orbslam3_b200 import LineMatcher
    lm = LineMatcher(max_pairs=64, max_train=512, max_query=512)
    try:
        cases = [stereo_line_case(seed, *SIZES[seed % len(SIZES)]) for seed in range(48)]
        for win in ((7, 0, 2, 2), (3, 3, 0, 1)):
            ms, ns = lm.matchGrid_batch(cases, INV_W, INV_H, window=win)
            total = 0
            for k, c in enumerate(cases):
                rn, rm = oracle.line_match_grid(*c, INV_W, INV_H, window=win)
                assert ns[k] == rn and np.array_equal(ms[k], rm), (k, win)
                total += rn
            assert total > 500
        c = stereo_line_case(7, 120, 140, 1280, 720)
        n, m = lm.matchGrid(*c, 32 / 1280.0, 24 / 720.0, grid_rows=24, grid_cols=32)
        rn, rm = oracle.line_match_grid(*c, 32 / 1280.0, 24 / 720.0, grid_rows=24, grid_cols=32)
        assert n == rn and np.array_equal(m, rm)
        with pytest.raises(Exception):
            lm.matchGrid(*c, INV_W, INV_H, grid_rows=65, grid_cols=64)
        for k in (0, 3, 5, 6, 7):      # host-buffer entry (the shim's call), including the empty sides
            rn, rm = oracle.line_match_grid(*cases[k], INV_W, INV_H)
            n, m = lm.matchGrid_host(*cases[k], INV_W, INV_H)
            assert n == rn and np.array_equal(m, rm), k
    finally:
        lm.close()


@pytest.mark.gpu
def test_match_grid_gpu_on_extracted_lines(gpu):
    """A synthetic frame and its copy shifted 12 px to the left as the right image: keylines and LBD descriptors
    from the CUDA Lineextractor, the search against the oracle (and the live reference where it exists)."""
    from pl_vi_orbslam3_b200 import LineMatcher, Lineextractor, synth
    left = synth.frame_euroc(11)
    right = np.ascontiguousarray(np.roll(left, -12, axis=1))
    l = Lineextractor(200, 0, 0.8, 2, 2.0, 0)
    lm = LineMatcher(max_pairs=1, max_train=256, max_query=256)
    try:
        k1, d1, _ = l(left)
        k2, d2, _ = l(right)
        seg = lambda k: np.stack([k["startPointX"], k["startPointY"], k["endPointX"], k["endPointY"]], 1).astype(np.float32)
        n, m = lm.matchGrid(seg(k1), d1, seg(k2), d2, INV_W, INV_H)
        rn, rm = oracle.line_match_grid(seg(k1), d1, seg(k2), d2, INV_W, INV_H)
        assert n == rn > 30 and np.array_equal(m, rm)
        if oracle.ref_available():
            fn, fm = oracle.ref_line_match_grid(seg(k1), d1, seg(k2), d2, INV_W, INV_H)
            assert fn == n and np.array_equal(fm, m)
    finally:
        l.close()
        lm.close()
