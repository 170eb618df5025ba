"""GPU: BASELINE.json's full-size configurations, checked through size-independent
properties (batch invariance, duplicate frames, shard invariance, determinism, Hamming
identities) plus oracle spot checks on sampled frames.

  config 4: 640x480 and 1280x720 frames at 2000 ORB features / 8 levels, batched 256 per GPU
  config 5: EuRoC-shaped sequence sharded by contiguous frame range (a 256-frame slice here)
  configs 2+3: the batched front-end step (ORB + lines + frame-to-frame matching)
"""
import zlib

import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import Lineextractor, ORBextractor, synth
from pl_vi_orbslam3_b200.frontend import FrontEnd, shard_range

pytestmark = pytest.mark.gpu


def _crc(*arrays):
    c = 0
    for a in arrays:
        c = zlib.crc32(np.ascontiguousarray(a).tobytes(), c)
    return c


def _frame_checksums(kps, desc, counts):
    return np.array([_crc(kps[i, :counts[i]], desc[i, :counts[i]]) for i in range(len(counts))], np.uint32)


@pytest.mark.parametrize("w,h,batch,check", [(640, 480, 256, (0, 100, 255)), (1280, 720, 64, (0, 63))])
def test_config4_batched_orb_2000_features(gpu, w, h, batch, check):
    frames = synth.frame_batch(batch, w, h, base_seed=500, distinct=8)
    frames[17] = frames[3]                                   # planted duplicate
    e = ORBextractor(2000, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=batch)
    try:
        kps, desc, counts, mono = e.extract_batch(frames)
        sums = _frame_checksums(kps, desc, counts)
        assert (counts > 1500).all() and (counts <= e.capacity).all() and np.array_equal(counts, mono)
        assert sums[17] == sums[3]                            # identical frames -> identical features
        # determinism: a second run gives the same checksum of checksums
        k2, d2, c2, _ = e.extract_batch(frames)
        assert _crc(_frame_checksums(k2, d2, c2)) == _crc(sums)
        # batch invariance: a frame alone == the same frame inside the batch
        for i in check:
            m1, k1, d1 = e(frames[i])
            assert _crc(k1, d1) == sums[i]
        # oracle spot check
        i = check[-1]
        ref = oracle.orb_extract(frames[i], nfeatures=2000)
        assert counts[i] == len(ref["keypoints"])
        for fld in ("x", "y", "octave", "response", "angle"):
            assert np.array_equal(kps[i, :counts[i]][fld], ref["keypoints"][fld]), fld
        assert np.array_equal(desc[i, :counts[i]], ref["descriptors"])      # every ORB bit
        # structural: octaves ascending (mono fill), responses in FAST range
        k0 = kps[0, :counts[0]]
        assert (np.diff(k0["octave"]) >= 0).all() and k0["response"].min() >= 7
    finally:
        e.close()


def test_config5_sequence_shards_equal_full_batch(gpu):
    """A contiguous shard of the sequence gives the same per-frame results as the full batch."""
    n = 256
    frames = synth.frame_batch(n, 752, 480, base_seed=900, distinct=8)
    e = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=n)
    l = Lineextractor(200, 0, 0.8, 2, 2.0, 0, max_batch=n)
    try:
        kps, desc, counts, _ = e.extract_batch(frames)
        kl, ld, eq, lc = l.extract_batch(frames)
        full = _frame_checksums(kps, desc, counts)
        lfull = np.array([_crc(kl[i, :lc[i]], ld[i, :lc[i]], eq[i, :lc[i]]) for i in range(n)], np.uint32)
        assert (lc == 200).all()
        for world in (2, 8):
            got, lgot = [], []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                k, d, c, _ = e.extract_batch(frames[lo:hi])
                got.append(_frame_checksums(k, d, c))
                a, b_, c_, d_ = l.extract_batch(frames[lo:hi])
                lgot.append(np.array([_crc(a[i, :d_[i]], b_[i, :d_[i]], c_[i, :d_[i]]) for i in range(hi - lo)], np.uint32))
            assert np.array_equal(np.concatenate(got), full)
            assert np.array_equal(np.concatenate(lgot), lfull)
        # oracle spot check of the lines of one frame
        ref = oracle.line_extract(frames[200])
        assert np.array_equal(kl[200, :200]["startPointX"], ref["keylines"]["startPointX"])
        assert np.array_equal(ld[200, :200], ref["descriptors"])            # every LBD bit
        assert np.array_equal(kl[200, :200].view(np.uint8), ref["keylines"].view(np.uint8))
    finally:
        e.close()
        l.close()


def test_frontend_step_matches_oracle(gpu):
    """configs 2+3 through the batched pipeline (device-resident buffers, two streams)."""
    import torch
    from pl_vi_orbslam3_b200.capi import QUERY_DTYPE
    from pl_vi_orbslam3_b200.matchers import frame_grid
    n = 6
    f1, f2, _ = synth.warp_pair(3)
    frames = np.stack([f1, f2, synth.frame_euroc(1), synth.frame_euroc(1), f2, f1])
    fe = FrontEnd(n)
    try:
        d = torch.from_numpy(frames).cuda()
        with torch.cuda.stream(fe.stream):
            launches = fe.step(d)
        fe.stream.synchronize()
        assert launches >= 25
        out = {k: v.cpu().numpy() for k, v in fe.outputs().items()}
        counts, lcounts = out["counts"], out["line_counts"]
        kps = out["kps"].view(np.uint8).reshape(n, fe.orb.capacity, 28).copy().view(oracle.KEYPOINT_DTYPE)[..., 0]
        grid = frame_grid(0, 752, 0, 480)
        refs = [oracle.orb_extract(f) for f in frames]
        lrefs = [oracle.line_extract(f) for f in frames]
        for p in range(n - 1):
            k = refs[p]["keypoints"]
            q = np.zeros(len(k), QUERY_DTYPE)
            q["u"], q["v"] = k["x"], k["y"]
            q["radius"] = np.float32(15.0) * (np.float32(1.2) ** k["octave"].astype(np.float32))
            q["min_level"], q["max_level"], q["angle"] = k["octave"] - 1, k["octave"] + 1, k["angle"]
            rn, rmt = oracle.search_frame(refs[p + 1]["keypoints"], refs[p + 1]["descriptors"], grid, q,
                                          refs[p]["descriptors"], 100, True)
            assert out["nmatches"][p] == rn
            assert np.array_equal(out["match_train"][p, :counts[p + 1]], rmt)
            ln, lm12 = oracle.line_match(lrefs[p]["descriptors"], lrefs[p + 1]["descriptors"], 0.9)
            assert out["line_nmatches"][p] == ln
            assert np.array_equal(out["line_matches"][p, :lcounts[p]], lm12)
        # identical consecutive frames (2,3): every keypoint matches itself at distance 0
        same = out["match_train"][2, :counts[3]]
        assert (same >= 0).sum() >= 0.9 * counts[3] and np.array_equal(kps[2, :counts[2]], kps[3, :counts[3]])
    finally:
        fe.close()
