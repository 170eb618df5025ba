"""GPU parity against THE REFERENCE'S OWN CODE: the CUDA path (through the C ABI) vs the outputs of
oracle/_ref/libplvi_ref.so (reference sources compiled unmodified, see tests/test_oracle_vs_ref.py) -- the committed
vectors of tests/golden/ref_outputs.npz and, where the prebuilt library travelled with the snapshot, live runs.

Bar: keypoints (all fields incl. angle), ORB descriptors, all KeyLine fields, LBD bytes and line equations bit-exact.
"""
from pathlib import Path

import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import Lineextractor, ORBextractor, synth
from test_oracle_vs_ref import CASES, R, frame

pytestmark = pytest.mark.gpu


def check_orb(kps, desc, mono, rk, rd, rmono):
    assert len(kps) == len(rk) and mono == int(rmono)
    for f in rk.dtype.names:
        assert np.array_equal(kps[f], rk[f]), f
    assert np.array_equal(desc, rd)


def check_lines(kl, desc, eq, rk, rd, req):
    assert len(kl) == len(rk)
    for f in rk.dtype.names:
        assert np.array_equal(kl[f], rk[f]), f
    assert np.array_equal(desc, rd)
    assert np.array_equal(eq, req)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cuda_equals_reference_outputs(gpu, case):
    name, nf, lap0, lap1, lnf = case[0], int(case[1]), int(case[2]), int(case[3]), int(case[4])
    img = frame(name)
    h, w = img.shape
    e = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    try:
        mono, kps, desc = e(img, vLappingArea=(lap0, lap1))
        check_orb(kps, desc, mono, R[f"{name}/orb_kp"], R[f"{name}/orb_desc"], R[f"{name}/orb_mono"])
    finally:
        e.close()
    le = Lineextractor(lnf, 0, 0.8, 2, 2.0, 0, max_width=w, max_height=h, max_batch=1)
    try:
        kl, ld, eq = le(img)
        check_lines(kl, ld, eq, R[f"{name}/line_kl"], R[f"{name}/line_desc"], R[f"{name}/line_eq"])
    finally:
        le.close()


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libplvi_ref.so did not travel")
def test_cuda_equals_live_reference_batch(gpu):
    frames = np.stack([synth.frame_euroc(100 + s) for s in range(8)])
    e = ORBextractor(1000, 1.2, 8, 20, 7, max_width=752, max_height=480, max_batch=8)
    le = Lineextractor(200, 0, 0.8, 2, 2.0, 0, max_width=752, max_height=480, max_batch=8)
    try:
        kps, desc, counts, mono = e.extract_batch(frames)
        kl, ld, eq, lc = le.extract_batch(frames)
        for i in range(len(frames)):
            r = oracle.ref_orb_extract(frames[i])
            n = counts[i]
            check_orb(kps[i, :n], desc[i, :n], mono[i], r["keypoints"], r["descriptors"], r["mono_index"])
            rl = oracle.ref_line_extract(frames[i])
            m = lc[i]
            check_lines(kl[i, :m], ld[i, :m], eq[i, :m], rl["keylines"], rl["descriptors"], rl["line_eq"])
    finally:
        e.close()
        le.close()


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libplvi_ref.so did not travel")
def test_cuda_bow_equals_reference_dbow2(gpu, tmp_path):
    """Frame::ComputeBoW on the GPU against the reference's own DBoW2 (loadFromTextFile + transform)."""
    import torch
    from pl_vi_orbslam3_b200.vocabulary import ORBVocabulary
    v = ORBVocabulary.random_tree(k=10, L=5, seed=4, stop_fraction=0.03, early_leaf_fraction=0.05)
    path = tmp_path / "voc.txt"
    v.save_text(path)
    path.write_text(path.read_text().rstrip("\n"))
    frames = np.stack([synth.frame_euroc(s) for s in (3, 9)])
    e = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=2)
    try:
        kps, desc, counts, _ = e.extract_batch_device(torch.from_numpy(frames).cuda())
        out = {k: t.cpu().numpy() for k, t in v.transform(desc, counts, 4).items()}
        d, c = desc.cpu().numpy(), counts.cpu().numpy()
        for f in range(2):
            r = oracle.ref_bow_transform(path, d[f, :c[f]], 4)
            bw, bv = r["bow"]
            assert out["bow_count"][f] == len(bw)
            assert np.array_equal(out["bow_words"][f, :len(bw)], bw) and np.array_equal(out["bow_values"][f, :len(bw)], bv)
            fn, fs, ff = r["fv"]
            assert out["fv_count"][f] == len(fn)
            assert np.array_equal(out["fv_nodes"][f, :len(fn)], fn) and np.array_equal(out["fv_start"][f, :len(fn) + 1], fs)
            assert np.array_equal(out["fv_features"][f, :fs[-1]], ff)
    finally:
        e.close()
        v.close()


@pytest.mark.parametrize("seed,scale,refine", [(0, 0.8, 1), (0, 0.8, 2), (3, 0.6, 1), (3, 1.0, 2), (5, 1.0, 0), (5, 0.5, 0)])
def test_cuda_lsd_equals_committed_reference_refine_and_scale_outputs(gpu, seed, scale, refine):
    """Raw LSD segments of the CUDA path against the reference's own lsd.cpp outputs for lsd_refine 1 / 2 and lsd_scale
    1.0 / 0.6 / 0.5 (tests/golden/lsd_extra.npz, tools/gen_golden_lsd.py).  refine 0: equal; refine > 0: same count and end
    points within 0.5 px (tree sums over lanes, CUDA libm in the NFA: DESIGN.md section 4)."""
    from pathlib import Path
    from pl_vi_orbslam3_b200 import Lineextractor, synth
    X = np.load(Path(__file__).resolve().parent / "golden" / "lsd_extra.npz")
    ref = X[f"ref_lsd_{seed}_{scale}_{refine}"]
    e = Lineextractor(0, refine, scale, 1, 2.0, 0, max_batch=1)
    try:
        e.set_debug(True)
        e(synth.frame_euroc(seed))
        got = e.read_lsd(0, 0, "segments", 752, 480)
        assert len(got) == len(ref)
        if refine == 0:
            assert np.array_equal(got, ref)
        else:
            assert np.abs(got - ref).max() <= 0.5
    finally:
        e.close()
