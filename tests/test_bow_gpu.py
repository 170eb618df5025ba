"""GPU parity of Frame::ComputeBoW (DBoW2 transform) through the C ABI against the oracle: per-feature
word / weight / node, BowVector (ids and values, bit-exact doubles) and FeatureVector."""
import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import ORBextractor, synth
from pl_vi_orbslam3_b200.vocabulary import ORBVocabulary

pytestmark = pytest.mark.gpu


def _check(v, out, desc, counts, levelsup):
    o = {k: t.cpu().numpy() for k, t in out.items()}
    for f in range(len(counts)):
        n = int(counts[f])
        r = oracle.bow_transform(v.as_oracle_dict(), desc[f, :n], levelsup)
        assert np.array_equal(o["word_id"][f, :n], r["word_id"])
        assert np.array_equal(o["word_weight"][f, :n], r["word_weight"])
        assert np.array_equal(o["node_id"][f, :n], r["node_id"])
        bw, bv = r["bow"]
        assert o["bow_count"][f] == len(bw)
        assert np.array_equal(o["bow_words"][f, :len(bw)], bw)
        assert np.array_equal(o["bow_values"][f, :len(bw)], bv)                   # same sums in the same order
        fn, fs, ff = r["fv"]
        assert o["fv_count"][f] == len(fn)
        assert np.array_equal(o["fv_nodes"][f, :len(fn)], fn)
        assert np.array_equal(o["fv_start"][f, :len(fn) + 1], fs)
        assert np.array_equal(o["fv_features"][f, :fs[-1]], ff)


@pytest.mark.parametrize("k,L,levelsup,nfeat", [(10, 4, 2, 1000), (10, 6, 4, 1000), (8, 5, 4, 2500)])
def test_bow_transform_after_extraction(gpu, k, L, levelsup, nfeat):
    import torch
    v = ORBVocabulary.random_tree(k=k, L=L, seed=L, stop_fraction=0.05)
    frames = np.stack([synth.frame_euroc(s) for s in (0, 1, 7)])
    e = ORBextractor(nfeat, 1.2, 8, 20, 7, max_batch=3)
    try:
        kps, desc, counts, _ = e.extract_batch_device(torch.from_numpy(frames).cuda())
        out = v.transform(desc, counts, levelsup)
        torch.cuda.synchronize()
        _check(v, out, desc.cpu().numpy(), counts.cpu().numpy(), levelsup)
    finally:
        e.close()
        v.close()


@pytest.mark.parametrize("scoring,weighting", [(0, 0), (1, 1), (5, 0), (0, 2), (5, 3)])
def test_bow_variants_irregular_tree(gpu, scoring, weighting):
    import torch
    v = ORBVocabulary.random_tree(k=5, L=5, seed=9, stop_fraction=0.2, early_leaf_fraction=0.25, scoring=scoring, weighting=weighting)
    rng = np.random.RandomState(3)
    desc = rng.randint(0, 256, (2, 700, 32)).astype(np.uint8)
    desc[0, 50:120] = v.desc[rng.randint(1, len(v.desc), 70)]                    # ties / repeated words
    desc[1, :300] = desc[1, 300:600]
    counts = np.array([700, 613], np.int32)
    try:
        out = v.transform(torch.from_numpy(desc).cuda(), torch.from_numpy(counts).cuda(), 3)
        torch.cuda.synchronize()
        _check(v, out, desc, counts, 3)
        out0 = v.transform(torch.from_numpy(desc).cuda(), torch.from_numpy(np.array([0, 5], np.int32)).cuda(), 3)
        torch.cuda.synchronize()
        assert out0["bow_count"].cpu().numpy()[0] == 0 and out0["fv_count"].cpu().numpy()[0] == 0   # empty frame
    finally:
        v.close()
