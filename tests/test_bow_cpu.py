"""CPU: the oracle's DBoW2 transform against an independent dictionary-based model, and the
ORBvoc text format round trip.  (The reference ships no vocabulary and no test vectors.)"""
import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200.vocabulary import ORBVocabulary


def _model(v, feats, levelsup):
    """TemplatedVocabulary::transform written with python dicts (double arithmetic = python float)."""
    n = len(v.parent)
    children = [[] for _ in range(n)]
    for i in range(1, n):
        children[v.parent[i]].append(i)
    wid, nw = {}, 0
    for i in range(1, n):
        if not children[i]:
            wid[i] = nw
            nw += 1
    bits = np.unpackbits(v.desc, axis=1)
    bow, fv = {}, {}
    for f, q in enumerate(np.unpackbits(feats, axis=1)):
        node, level, nid = 0, 0, 0
        while True:
            level += 1
            ch = children[node]
            d = [int((bits[c] != q).sum()) for c in ch]
            node = ch[int(np.argmin(d))]                 # first minimum
            if level == v.L - levelsup:
                nid = node
            if not children[node]:
                break
        w = float(v.weight[node])
        if w > 0:
            if v.weighting <= 1:
                bow[wid[node]] = bow.get(wid[node], 0.0) + w if wid[node] in bow else w
            else:
                bow.setdefault(wid[node], w)
            fv.setdefault(nid, []).append(f)
    norm_mode = 2 if v.scoring == 1 else (0 if v.scoring == 5 else 1)
    keys = sorted(bow)
    if v.weighting <= 1 and bow and norm_mode == 0:
        for k in keys:
            bow[k] /= float(len(bow))
    if norm_mode:
        norm = 0.0
        for k in keys:
            norm += abs(bow[k]) if norm_mode == 1 else bow[k] * bow[k]
        if norm_mode == 2:
            norm = float(np.sqrt(norm))
        if norm > 0:
            for k in keys:
                bow[k] /= norm
    return keys, [bow[k] for k in keys], sorted(fv), [fv[k] for k in sorted(fv)]


@pytest.mark.parametrize("scoring,weighting", [(0, 0), (0, 1), (1, 0), (5, 0), (0, 2), (0, 3), (5, 3)])
def test_oracle_transform_matches_dict_model(scoring, weighting):
    v = ORBVocabulary.random_tree(k=4, L=4, seed=5, stop_fraction=0.15, early_leaf_fraction=0.2, scoring=scoring, weighting=weighting)
    rng = np.random.RandomState(1)
    feats = rng.randint(0, 256, (300, 32)).astype(np.uint8)
    feats[10:40] = v.desc[rng.randint(1, len(v.desc), 30)]          # exact node descriptors: ties and repeats
    for levelsup in (1, 2, 6):
        r = oracle.bow_transform(v.as_oracle_dict(), feats, levelsup)
        words, vals, nodes, lists = _model(v, feats, levelsup)
        assert list(r["bow"][0]) == words
        assert np.array_equal(r["bow"][1], np.array(vals))
        assert list(r["fv"][0]) == nodes
        for i, lst in enumerate(lists):
            assert list(r["fv"][2][r["fv"][1][i]:r["fv"][1][i + 1]]) == lst
        if scoring in (0,) and len(vals):
            assert abs(sum(vals) - 1.0) < 1e-12                       # L1 normalised


def test_vocabulary_text_roundtrip(tmp_path):
    v = ORBVocabulary.random_tree(k=3, L=3, seed=2, stop_fraction=0.1)
    p = tmp_path / "voc.txt"
    v.save_text(p)
    w = ORBVocabulary.load_text(p)
    assert (w.k, w.L, w.scoring, w.weighting) == (3, 3, 0, 0) and w.words == v.words == 27
    assert np.array_equal(w.parent, v.parent) and np.array_equal(w.desc[1:], v.desc[1:]) and np.array_equal(w.weight, v.weight)   # the root has no line
    assert open(p).readline().split() == ["3", "3", "0", "0"]
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.txt"
        bad.write_text("50 6 0 0\n")
        ORBVocabulary.load_text(bad)
