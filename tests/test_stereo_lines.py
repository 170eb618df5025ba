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


def _keylines(seg):
    kl = np.zeros(len(seg), oracle.KEYLINE_DTYPE)
    for j, f in enumerate(("startPointX", "startPointY", "endPointX", "endPointY")):
        kl[f] = seg[:, j]
    return kl


def _depth_case(seed):
    n1, n2 = SIZES[seed % len(SIZES)]
    s1, d1, s2, d2 = stereo_line_case(seed, n1, n2)
    if seed % 3 == 1 and n1:      # some left lines shifted vertically: partial / no overlap with their right line
        s1 = s1.copy()
        s1[::2, 1] += np.float32(9.0)
        s1[::2, 3] += np.float32(9.0)
    if seed % 4 == 2 and n2 > 4:  # horizontal right lines: the division by sp_r(1) - ep_r(1) = 0
        s2 = s2.copy()
        s2[4:8, 3] = s2[4:8, 1]
    return s1, d1, s2, d2


@needs_ref
def test_oracle_stereo_line_depth_against_live_reference():
    """Grid fill + matchGrid + the disparity / overlap / depth filter + mvle_l: the oracle's restatement against the
    reference's own Frame::ComputeStereoMatches_Lines (src/Frame.cc:1408-1500, Frame.cc compiled unmodified)."""
    total = 0
    for seed in range(32):
        s1, d1, s2, d2 = _depth_case(seed)
        k, out, le = oracle.ref_frame_stereo_lines(_keylines(s1), d1, _keylines(s2), d2, INV_W, INV_H, 47.9)
        n, m12 = oracle.line_match_grid(s1, d1, s2, d2, INV_W, INV_H)
        k2, disp, dep, le2 = oracle.line_stereo_depth(s1, s2, m12, 47.9)
        assert k == k2, seed
        assert np.array_equal(out[:, :2].view(np.uint32), disp.view(np.uint32)) and np.array_equal(out[:, 2:].view(np.uint32), dep.view(np.uint32)), seed
        assert np.array_equal(le.view(np.uint64), le2.view(np.uint64)), seed
        total += k
    assert total > 300


@pytest.mark.gpu
def test_stereo_line_depth_gpu(gpu):
    """plvi_line_stereo_depth (device batch) and plvi_line_stereo_depth_host against the oracle and the live reference."""
    import torch
    from pl_vi_orbslam3_b200 import LineMatcher
    from pl_vi_orbslam3_b200.capi import check, lib, ptr
    lm = LineMatcher(max_pairs=64, max_train=512, max_query=512)
    try:
        cases = [_depth_case(seed) for seed in range(32)]
        P, S = len(cases), 320
        seg1 = np.zeros((P, S, 4), np.float32); seg2 = np.zeros((P, S, 4), np.float32); segu = np.zeros((P, S, 4), np.float32)
        m12 = np.full((P, S), -1, np.int32); n1 = np.zeros(P, np.int32); n2 = np.zeros(P, np.int32)
        rng = np.random.RandomState(3)
        want = []
        for k, (s1, d1, s2, d2) in enumerate(cases):
            n1[k], n2[k] = len(s1), len(s2)
            seg1[k, :len(s1)] = s1; seg2[k, :len(s2)] = s2
            segu[k, :len(s1)] = s1 + rng.normal(0, 0.7, s1.shape).astype(np.float32)      # "undistorted" end points
            _, m = oracle.line_match_grid(s1, d1, s2, d2, INV_W, INV_H)
            m12[k, :len(s1)] = m
            want.append(oracle.line_stereo_depth(s1, s2, m, 47.9, segu[k, :len(s1)]))
        dev = lambda a: torch.from_numpy(a).cuda()
        t = [dev(a) for a in (seg1, n1, seg2, n2, m12, segu)]
        disp = torch.zeros((P, S, 2), dtype=torch.float32, device="cuda"); dep = torch.zeros_like(disp)
        le = torch.zeros((P, S, 3), dtype=torch.float64, device="cuda"); nd = torch.zeros(P, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        check(lib().plvi_line_stereo_depth(lm._h, P, ptr(t[0]), ptr(t[1]), S, ptr(t[2]), ptr(t[3]), S, ptr(t[4]), ptr(t[5]), 47.9,
                                           ptr(disp), ptr(dep), ptr(le), ptr(nd)))
        lm.sync()
        disp, dep, le, nd = disp.cpu().numpy(), dep.cpu().numpy(), le.cpu().numpy(), nd.cpu().numpy()
        total = 0
        for k, (kk, wd, wp, wl) in enumerate(want):
            n = n1[k]
            assert nd[k] == kk, k
            assert np.array_equal(disp[k, :n].view(np.uint32), wd.view(np.uint32)) and np.array_equal(dep[k, :n].view(np.uint32), wp.view(np.uint32)), k
            assert np.array_equal(le[k, :n].view(np.uint64), wl.view(np.uint64)), k
            total += kk
        assert total > 300
        for k in (0, 1, 2, 5, 6, 9):     # host-buffer entry, incl. the empty sides; the live reference where it travelled
            s1, d1, s2, d2 = cases[k]
            n, m = lm.matchGrid_host(s1, d1, s2, d2, INV_W, INV_H)
            kk, gd, gp, gl = lm.stereo_depth_host(s1, s2, m, 47.9)
            wk, wd, wp, wl = oracle.line_stereo_depth(s1, s2, m, 47.9)
            assert kk == wk and np.array_equal(gd.view(np.uint32), wd.view(np.uint32)) and np.array_equal(gp.view(np.uint32), wp.view(np.uint32))
            assert np.array_equal(gl.view(np.uint64), wl.view(np.uint64))
            if oracle.ref_available():
                rk, rout, rle = oracle.ref_frame_stereo_lines(_keylines(s1), d1, _keylines(s2), d2, INV_W, INV_H, 47.9)
                assert rk == kk and np.array_equal(rout[:, :2].view(np.uint32), gd.view(np.uint32)) and np.array_equal(rle.view(np.uint64), gl.view(np.uint64))
    finally:
        lm.close()


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
