"""GPU parity: CUDA ORB path (through the C ABI) vs the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): pyramid pixels, FAST candidates/scores and keypoint
sets bit-exact; orientation within 1e-3 rad; ORB descriptor bits >= 99.9 %.
"""
import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import ORBextractor, synth

pytestmark = pytest.mark.gpu

ANGLE_TOL_DEG = np.degrees(1e-3)
DESC_BIT_AGREEMENT = 0.999


def _bits_equal(a, b):
    return 1.0 - np.unpackbits(a ^ b).mean() if a.size else 1.0


def _check_frame(kps, desc, mono, ref, exact_desc_required=False):
    rk, rd = ref["keypoints"], ref["descriptors"]
    assert len(kps) == len(rk), (len(kps), len(rk))
    assert mono == ref["mono_index"]
    for fld in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps[fld], rk[fld]), fld
    da = np.abs(kps["angle"] - rk["angle"])
    da = np.minimum(da, 360.0 - da)
    assert da.max(initial=0.0) <= ANGLE_TOL_DEG, da.max()
    agree = _bits_equal(desc, rd)
    assert agree >= DESC_BIT_AGREEMENT, agree
    return float(da.max(initial=0.0)), agree


@pytest.fixture(scope="module")
def ext(gpu):
    e = ORBextractor(1000, 1.2, 8, 20, 7, max_width=752, max_height=480, max_batch=8)
    yield e
    e.close()


def test_pyramid_blur_candidates_bit_exact(ext):
    img = synth.frame_euroc(0)
    ext(img)
    ref = oracle.orb_extract(img, debug=True)
    for l in range(8):
        got = ext.read_level(0, l, 752, 480)
        assert np.array_equal(got, ref["pyramid"][l]), f"pyramid level {l}"
        got = ext.read_level(0, l, 752, 480, blurred=True)
        assert np.array_equal(got, ref["blurred"][l]), f"blurred level {l}"
        c = ext.read_candidates(0, l)
        rc = oracle.grid_fast(ref["pyramid"][l]).astype(np.int32)
        c = c[np.lexsort((c[:, 0], c[:, 1]))]
        rc = rc[np.lexsort((rc[:, 0], rc[:, 1]))]
        assert np.array_equal(c, rc), f"FAST candidates level {l}: {len(c)} vs {len(rc)}"


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 7])
def test_orb_single_frame_matches_oracle(ext, seed):
    img = synth.frame_euroc(seed)
    mono, kps, desc = ext(img)
    ref = oracle.orb_extract(img)
    _check_frame(kps, desc, mono, ref)


def test_orb_batch_equals_single(ext):
    frames = np.stack([synth.frame_euroc(s) for s in range(10, 18)])
    kps, desc, counts, mono = ext.extract_batch(frames)
    for i in range(len(frames)):
        ref = oracle.orb_extract(frames[i])
        n = counts[i]
        _check_frame(kps[i, :n], desc[i, :n], mono[i], ref)


def test_orb_lapping_area_reverse_fill(ext):
    img = synth.frame_euroc(5)
    for lap in ((0, 1000), (300, 500)):
        mono, kps, desc = ext(img, None, lap)
        ref = oracle.orb_extract(img, lapping=lap)
        _check_frame(kps, desc, mono, ref)


def test_orb_flat_and_noise_images(ext):
    flat = np.full((480, 752), 128, np.uint8)
    mono, kps, desc = ext(flat)
    assert mono == 0 and len(kps) == 0
    rng = np.random.RandomState(3)
    noise = rng.randint(0, 256, (480, 752)).astype(np.uint8)
    mono, kps, desc = ext(noise)
    _check_frame(kps, desc, mono, oracle.orb_extract(noise))


def test_orb_empty_image_returns_minus_one(ext):
    mono, kps, desc = ext(np.zeros((0, 0), np.uint8))
    assert mono == -1 and len(kps) == 0


def test_orb_strided_input(ext):
    big = np.zeros((480, 800), np.uint8)
    img = synth.frame_euroc(21)
    big[:, :752] = img
    view = big[:, :752]
    kps, desc, counts, mono = ext.extract_batch.__func__(ext, np.ascontiguousarray(view)[None])
    _check_frame(kps[0, :counts[0]], desc[0, :counts[0]], mono[0], oracle.orb_extract(img))


def test_orb_device_resident_api(ext):
    import torch
    frames = np.stack([synth.frame_euroc(s) for s in (30, 31)])
    d = torch.from_numpy(frames).cuda()
    import ctypes
    st = torch.cuda.ExternalStream(ext.stream)
    with torch.cuda.stream(st):
        kps, desc, counts, mono = ext.extract_batch_device(d)
    st.synchronize()
    kps = kps.cpu().numpy().view(np.uint8).reshape(2, ext.capacity, 28).copy().view(oracle.KEYPOINT_DTYPE)[..., 0]
    desc, counts, mono = desc.cpu().numpy(), counts.cpu().numpy(), mono.cpu().numpy()
    for i in range(2):
        n = counts[i]
        _check_frame(kps[i, :n], desc[i, :n], mono[i], oracle.orb_extract(frames[i]))


@pytest.mark.parametrize("w,h,nfeat", [(640, 480, 2000), (1280, 720, 2000), (752, 480, 5000)])
def test_orb_other_configs(gpu, w, h, nfeat):
    e = ORBextractor(nfeat, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2)
    try:
        frames = np.stack([synth.frame_euroc(40 + i, w, h) for i in range(2)])
        kps, desc, counts, mono = e.extract_batch(frames)
        for i in range(2):
            ref = oracle.orb_extract(frames[i], nfeatures=nfeat)
            n = counts[i]
            _check_frame(kps[i, :n], desc[i, :n], mono[i], ref)
    finally:
        e.close()


def test_orb_smaller_image_on_bigger_handle(gpu):
    e = ORBextractor(1000, 1.2, 8, 20, 7, max_width=1280, max_height=720, max_batch=2)
    try:
        for (w, h) in ((752, 480), (640, 480), (1280, 720)):
            img = synth.frame_euroc(50, w, h)
            mono, kps, desc = e(img)
            _check_frame(kps, desc, mono, oracle.orb_extract(img))
    finally:
        e.close()


@pytest.mark.parametrize("w,h", [(750, 481), (641, 479), (333, 250)])
def test_orb_odd_sizes_and_unaligned_device_input(gpu, w, h):
    """Widths that are not multiples of 4 exercise the unaligned tile-load paths; the device API
    reads the caller's buffer in place (row stride = width, odd base offset)."""
    import torch
    e = ORBextractor(1000, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2)
    try:
        frames = np.stack([synth.frame_euroc(80 + i, w, h) for i in range(2)])
        kps, desc, counts, mono = e.extract_batch(frames)
        refs = [oracle.orb_extract(f) for f in frames]
        for i in range(2):
            _check_frame(kps[i, :counts[i]], desc[i, :counts[i]], mono[i], refs[i])
        # device-resident, deliberately misaligned by one byte
        flat = torch.zeros(frames.size + 8, dtype=torch.uint8, device="cuda")
        flat[1:1 + frames.size] = torch.from_numpy(frames.reshape(-1)).cuda()
        view = flat[1:1 + frames.size].view(2, h, w)
        st = torch.cuda.ExternalStream(e.stream)
        with torch.cuda.stream(st):
            dk, dd, dc, dm = e.extract_batch_device(view)
        st.synchronize()
        dk = dk.cpu().numpy().view(np.uint8).reshape(2, e.capacity, 28).copy().view(oracle.KEYPOINT_DTYPE)[..., 0]
        dd, dc, dm = dd.cpu().numpy(), dc.cpu().numpy(), dm.cpu().numpy()
        for i in range(2):
            _check_frame(dk[i, :dc[i]], dd[i, :dc[i]], dm[i], refs[i])
    finally:
        e.close()


@pytest.mark.parametrize("nfeat,sf,nlev,ini,mn", [(500, 1.5, 4, 30, 10), (1500, 1.1, 10, 15, 5), (300, 2.0, 3, 20, 7)])
def test_orb_other_extractor_parameters(gpu, nfeat, sf, nlev, ini, mn):
    e = ORBextractor(nfeat, sf, nlev, ini, mn, max_batch=1)
    try:
        img = synth.frame_euroc(90)
        mono, kps, desc = e(img)
        ref = oracle.orb_extract(img, nfeatures=nfeat, scale_factor=sf, nlevels=nlev, ini_th=ini, min_th=mn)
        _check_frame(kps, desc, mono, ref)
        assert np.array_equal(e.features_per_level(), ref["plan"]["quota"])
        assert np.array_equal(e.GetScaleFactors(), ref["plan"]["scale"])
    finally:
        e.close()


def test_cuda_graph_replay_is_used_and_identical(gpu):
    """The launch sequence of a batch is captured into a CUDA graph on first use and replayed afterwards; the
    results are the ones of plain launches (profiling switches graphs off)."""
    from pl_vi_orbslam3_b200 import Lineextractor
    frames = np.stack([synth.frame_euroc(40 + s) for s in range(4)])
    e = ORBextractor(1000, 1.2, 8, 20, 7, max_width=752, max_height=480, max_batch=4)
    le = Lineextractor(200, 0, 0.8, 2, 2.0, 0, max_width=752, max_height=480, max_batch=4)
    try:
        # the host-buffer entry points upload into two alternating staging buffers: one graph per buffer
        a = [e.extract_batch(frames) for _ in range(4)]
        la = [le.extract_batch(frames) for _ in range(4)]
        cap, rep = e.graph_stats()
        assert cap == 2 and rep == 2, (cap, rep)
        lcap, lrep = le.graph_stats()
        assert lcap == 2 and lrep == 2, (lcap, lrep)
        from pl_vi_orbslam3_b200.capi import lib
        lib().plvi_orb_set_profile(e._h, 1)
        lib().plvi_line_set_profile(le._h, 1)
        b = e.extract_batch(frames)          # plain launches
        lb = le.extract_batch(frames)
        assert e.graph_stats() == (2, 2) and le.graph_stats() == (2, 2)
        # rows beyond a frame's count are never written (stale device memory of whichever result set the call used)
        def same(x, y, cnt_idx):
            cx, cy = x[cnt_idx], y[cnt_idx]
            if not np.array_equal(cx, cy):
                return False
            for j, (u, v) in enumerate(zip(x, y)):
                if u.ndim == 1:
                    if not np.array_equal(u, v):
                        return False
                    continue
                for i, c in enumerate(cx):
                    if not np.array_equal(np.ascontiguousarray(u[i, :c]).view(np.uint8), np.ascontiguousarray(v[i, :c]).view(np.uint8)):
                        return False
            return True
        for x in a:
            assert same(x, b, 2)          # (kps, desc, counts, mono)
        for x in la:
            assert same(x, lb, 3)         # (keylines, desc, line_eq, counts)
        # another batch size is another graph
        lib().plvi_orb_set_profile(e._h, 0)
        e.extract_batch(frames[:2])
        assert e.graph_stats()[0] == 3
    finally:
        e.close()
        le.close()


def test_pyramid_ahead_and_wait_after_pyramid_give_the_same_result(ext):
    """Scheduling hooks of the batched front end: plvi_orb_pyramid_device builds the pyramid ahead of the extraction call
    (which then skips that stage), plvi_orb_wait_event_after_pyramid places an event wait behind the pyramid.  Neither
    changes a bit of the result; a pyramid built for OTHER images is not reused."""
    import torch
    from pl_vi_orbslam3_b200.capi import check, lib, ptr
    frames = np.stack([synth.frame_euroc(s) for s in (3, 4, 5, 6)])
    d = torch.from_numpy(frames).cuda()
    other = torch.from_numpy(np.stack([synth.frame_euroc(s) for s in (7, 8, 9, 10)])).cuda()
    base = ext.extract_batch_device(d)
    torch.cuda.synchronize()   # the handle has its own stream
    base = [t.clone() for t in base]
    torch.cuda.synchronize()
    n, h, w = d.shape
    # pyramid ahead, same images
    check(lib().plvi_orb_pyramid_device(ext._h, ptr(d), n, w, h, d.stride(1), d.stride(0)))
    got = ext.extract_batch_device(d)
    torch.cuda.synchronize()
    for a, b in zip(base, got):
        assert torch.equal(a.contiguous().view(torch.uint8), b.contiguous().view(torch.uint8))   # bytes: class_id -1 reads as NaN
    # pyramid of other images: the extraction call rebuilds its own
    check(lib().plvi_orb_pyramid_device(ext._h, ptr(other), n, w, h, other.stride(1), other.stride(0)))
    got = ext.extract_batch_device(d)
    torch.cuda.synchronize()
    for a, b in zip(base, got):
        assert torch.equal(a.contiguous().view(torch.uint8), b.contiguous().view(torch.uint8))   # bytes: class_id -1 reads as NaN
    # event wait behind the pyramid (an event that has already completed)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    check(lib().plvi_orb_wait_event_after_pyramid(ext._h, ptr(ev.cuda_event)))
    got = ext.extract_batch_device(d)
    torch.cuda.synchronize()
    for a, b in zip(base, got):
        assert torch.equal(a.contiguous().view(torch.uint8), b.contiguous().view(torch.uint8))   # bytes: class_id -1 reads as NaN
