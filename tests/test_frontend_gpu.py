"""GPU parity of the batched front end bench.py times (pl_vi_orbslam3_b200/frontend.py): C3 pairs through FrontEnd.step
(device-resident frames) and FrontEnd.step_host (the host-buffer C ABI: plvi_line_extract_batch_async +
plvi_orb_extract_batch_async, searches on the handles' device results) against the oracle and, when it travelled, the
reference's own ORBmatcher / LineMatcher code (oracle/_ref).

Bar: bit-exact (integer / index work).
"""
import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import frame_grid, synth
from pl_vi_orbslam3_b200.capi import KEYPOINT_DTYPE, QUERY_DTYPE
from pl_vi_orbslam3_b200.frontend import FrontEnd
from pl_vi_orbslam3_b200.matchers import ORBmatcher

pytestmark = pytest.mark.gpu

W, H = 752, 480
SCALES = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2))]).astype(np.float32), dtype=np.float32)


def _valid(out, key, counts_key, i):
    return out[key][i][: out[counts_key][i]]


def _check_pairs_against_oracle(out, A, B):
    grid = frame_grid(0, W, 0, H)
    for p in range(B // 2):
        n1, n2 = int(out["counts"][2 * p]), int(out["counts"][2 * p + 1])
        assert n1 > 500 and n2 > 500
        k1 = np.ascontiguousarray(out["kps"][2 * p][:n1]).view(KEYPOINT_DTYPE).reshape(-1)
        k2 = np.ascontiguousarray(out["kps"][2 * p + 1][:n2]).view(KEYPOINT_DTYPE).reshape(-1)
        d1, d2 = out["desc"][2 * p][:n1], out["desc"][2 * p + 1][:n2]
        q = np.zeros(n1, QUERY_DTYPE)
        q["u"] = (A[0] * k1["x"] + A[1] * k1["y"]) + A[2]
        q["v"] = (A[3] * k1["x"] + A[4] * k1["y"]) + A[5]
        q["radius"] = np.float32(15.0) * SCALES[k1["octave"]]
        q["min_level"], q["max_level"], q["angle"] = k1["octave"] - 1, k1["octave"] + 1, k1["angle"]
        q["flags"] = ((q["u"] < 0) | (q["u"] > W) | (q["v"] < 0) | (q["v"] > H)).astype(np.int32)
        n, mt = oracle.search_frame(k2, d2, grid, q, d1, 100, True)
        assert n == out["nmatches"][p] and n > 100
        assert np.array_equal(mt, out["match_train"][p][:n2])
        qi = ORBmatcher.init_queries(k1, np.stack([k1["x"], k1["y"]], 1), 100.0)
        ni, m12, _ = oracle.search_init(k2, d2, grid, qi, d1, 50, 0.9, True)
        assert ni == out["init_nmatches"][p] and ni > 20
        assert np.array_equal(m12, out["init_matches"][p][:n1])
        l1, l2 = int(out["line_counts"][2 * p]), int(out["line_counts"][2 * p + 1])
        ld1, ld2 = out["line_desc"][2 * p][:l1], out["line_desc"][2 * p + 1][:l2]
        nl, lm12 = oracle.line_match(ld1, ld2, 0.9)
        assert nl == out["line_nmatches"][p] and nl > 10
        assert np.array_equal(lm12, out["line_matches"][p][:l1])
        if oracle.ref_available():   # the reference's own searches on the same features
            rn, rmt = oracle.ref_search_frame(k2, d2, grid, (0.0, float(W), 0.0, float(H)), SCALES, k1, np.stack([q["u"], q["v"]], 1),
                                              np.zeros(n1, np.int32), d1, 15.0, True)
            assert rn == n and np.array_equal(rmt, mt)
            rni, rm12, _ = oracle.ref_search_init(k1, d1, k2, d2, grid, np.stack([k1["x"], k1["y"]], 1), 100, 0.9, True)
            assert rni == ni and np.array_equal(rm12, m12)
            rnl, rlm = oracle.ref_line_match(ld1, ld2, 0.9, "match")
            assert rnl == nl and np.array_equal(rlm, lm12)


@pytest.mark.parametrize("share", ["1", "0", "sched"])
def test_c3_pairs_device_path_and_host_buffer_path(gpu, share, monkeypatch):
    """share = 1: the frames are uploaded once (plvi_orb_extract_batch_async_from_line reads the line handle's copy);
    0: each extractor call uploads them; sched: the scheduling hooks between the two pipelines (ORB pyramid ahead of
    the line pipeline, ORB kernels behind the stage event AND the resident-block counter of the region-growing kernel,
    band speculation instead of the small-batch schedule) -- results are the same bits."""
    import torch
    kw = {}
    if share == "sched":
        share = "1"
        monkeypatch.setenv("PLVI_SKEW", "3")
        monkeypatch.setenv("PLVI_PYR_FIRST", "1")
        kw["band_run_max"] = 0
    monkeypatch.setenv("PLVI_SHARE_UPLOAD", share)
    B = 6
    frames = synth.pair_batch(B, W, H, base_seed=40, workers=1, cache=False)
    A = synth.warp_affine(W, H).astype(np.float32).reshape(6)
    fe = FrontEnd(B, w=W, h=H, pairs=True, affine=A, out_sets=2, **kw)
    try:
        d = torch.from_numpy(frames).cuda()
        with torch.cuda.stream(fe.stream):
            fe.step(d)
        fe.stream.synchronize()
        out = {k: v.cpu().numpy() for k, v in fe.outputs().items()}
        _check_pairs_against_oracle(out, A, B)

        hf = torch.from_numpy(frames).pin_memory()
        hf2 = torch.from_numpy(np.ascontiguousarray(frames[::-1])).pin_memory()
        ios = [fe.alloc_host_io() for _ in range(2)]
        # three calls back to back without waiting: the uploads of calls 2 and 3 overlap the kernels of the calls before
        # (two staging buffers); the middle call carries other frames, so a stale buffer would show
        fe.step_host(hf, ios[0])
        fe.step_host(hf2, ios[1])
        ev = fe.host_done_events()
        fe.step_host(hf, ios[0])
        fe.sync_host()
        for e in ev:
            assert e.query()
        host = {k: v.numpy() for k, v in ios[0].items()}
        for k in ("counts", "mono", "line_counts", "nmatches", "init_nmatches", "line_nmatches"):
            assert np.array_equal(host[k], out[k]), k
        for i in range(B):
            for k, c in (("kps", "counts"), ("desc", "counts"), ("keylines", "line_counts"), ("line_desc", "line_counts"),
                         ("line_eq", "line_counts")):
                assert np.array_equal(_valid(host, k, c, i).view(np.uint8), _valid(out, k, c, i).view(np.uint8)), (k, i)
        for p in range(B // 2):
            assert np.array_equal(host["match_train"][p][: out["counts"][2 * p + 1]], out["match_train"][p][: out["counts"][2 * p + 1]])
            assert np.array_equal(host["init_matches"][p][: out["counts"][2 * p]], out["init_matches"][p][: out["counts"][2 * p]])
            assert np.array_equal(host["line_matches"][p][: out["line_counts"][2 * p]], out["line_matches"][p][: out["line_counts"][2 * p]])
        # the reversed batch is pairs (warp, frame): extraction results must equal the forward run frame by frame
        h2 = {k: v.numpy() for k, v in ios[1].items()}
        assert np.array_equal(h2["counts"], out["counts"][::-1]) and np.array_equal(h2["line_counts"], out["line_counts"][::-1])
        for i in range(B):
            assert np.array_equal(_valid(h2, "desc", "counts", i), _valid(out, "desc", "counts", B - 1 - i))
            assert np.array_equal(_valid(h2, "line_desc", "line_counts", i), _valid(out, "line_desc", "line_counts", B - 1 - i))
    finally:
        fe.close()


def test_sequence_mode_host_buffer_path_equals_device_path(gpu):
    import torch
    B = 5
    frames = synth.seq_batch(B, 640, 480, base_seed=7, workers=1, cache=False)
    fe = FrontEnd(B, w=640, h=480, nfeatures=2000)
    try:
        with torch.cuda.stream(fe.stream):
            fe.step(torch.from_numpy(frames).cuda())
        fe.stream.synchronize()
        out = {k: v.cpu().numpy() for k, v in fe.outputs().items()}
        io = fe.alloc_host_io()
        fe.step_host(torch.from_numpy(frames).pin_memory(), io)
        fe.sync_host()
        host = {k: v.numpy() for k, v in io.items()}
        for k in ("counts", "line_counts", "nmatches", "line_nmatches"):
            assert np.array_equal(host[k], out[k]), k
        grid = frame_grid(0, 640, 0, 480)
        for p in range(B - 1):
            n1, n2 = out["counts"][p], out["counts"][p + 1]
            assert np.array_equal(host["match_train"][p][:n2], out["match_train"][p][:n2])
            k1 = np.ascontiguousarray(out["kps"][p][:n1]).view(KEYPOINT_DTYPE).reshape(-1)
            k2 = np.ascontiguousarray(out["kps"][p + 1][:n2]).view(KEYPOINT_DTYPE).reshape(-1)
            q = np.zeros(n1, QUERY_DTYPE)
            q["u"], q["v"] = k1["x"], k1["y"]
            q["radius"] = np.float32(15.0) * SCALES[k1["octave"]]
            q["min_level"], q["max_level"], q["angle"] = k1["octave"] - 1, k1["octave"] + 1, k1["angle"]
            n, mt = oracle.search_frame(k2, out["desc"][p + 1][:n2], grid, q, out["desc"][p][:n1], 100, True)
            assert n == out["nmatches"][p] and np.array_equal(mt, out["match_train"][p][:n2])
    finally:
        fe.close()
