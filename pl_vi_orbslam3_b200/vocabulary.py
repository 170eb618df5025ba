"""ORBVocabulary (include/ORBVocabulary.h = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>):
host mirror holding the tree as flat arrays, the text format of ORBvoc.txt
(TemplatedVocabulary::loadFromTextFile / saveToTextFile, TemplatedVocabulary.h:1338-1470) and
`transform` on device-resident descriptors (Frame::ComputeBoW)."""
import ctypes as C

import numpy as np

from .capi import check, lib, ptr


class ORBVocabulary:
    def __init__(self, k, L, scoring, weighting, parent, is_leaf, desc, weight, device=0):
        self.k, self.L, self.scoring, self.weighting = int(k), int(L), int(scoring), int(weighting)
        self.parent = np.ascontiguousarray(parent, np.int32)
        self.is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.weight = np.ascontiguousarray(weight, np.float64)
        assert len(self.parent) == len(self.is_leaf) == len(self.desc) == len(self.weight)
        self.device = device
        self._h = None     # device copy, made on first use

    def _handle(self):
        if self._h is None:
            h = C.c_void_p()
            check(lib().plvi_vocab_create(C.byref(h), self.k, self.L, self.scoring, self.weighting, len(self.parent),
                                          ptr(self.parent), ptr(self.is_leaf), ptr(self.desc), ptr(self.weight), self.device))
            self._h = h
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            lib().plvi_vocab_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def words(self):
        return int(self.is_leaf[1:].sum())

    def as_oracle_dict(self):
        return dict(k=self.k, L=self.L, scoring=self.scoring, weighting=self.weighting, parent=self.parent, desc=self.desc,
                    weight=self.weight)

    # ---- text format of ORBvoc.txt: "k L scoring weighting" then one line per node (id order, root excluded):
    # "parent isLeaf d0 ... d31 weight"
    @classmethod
    def load_text(cls, path, device=0):
        with open(path) as f:
            k, L, n1, n2 = (int(t) for t in f.readline().split()[:4])
            if k < 0 or k > 20 or L < 1 or L > 10 or n1 < 0 or n1 > 5 or n2 < 0 or n2 > 3:
                raise ValueError("Vocabulary loading failure: This is not a correct text file!")
            parent, leaf, desc, weight = [0], [0], [np.zeros(32, np.uint8)], [0.0]
            for line in f:
                t = line.split()
                if len(t) < 35:
                    continue
                parent.append(int(t[0]))
                leaf.append(1 if int(t[1]) > 0 else 0)
                desc.append(np.array([int(x) for x in t[2:34]], np.uint8))
                weight.append(float(t[34]))
        return cls(k, L, n1, n2, parent, leaf, np.stack(desc), weight, device)

    def save_text(self, path):
        with open(path, "w") as f:
            f.write(f"{self.k} {self.L} {self.scoring} {self.weighting}\n")
            for i in range(1, len(self.parent)):
                d = " ".join(str(int(x)) for x in self.desc[i])
                f.write(f"{self.parent[i]} {int(self.is_leaf[i])} {d} {float(self.weight[i])!r}\n")

    @classmethod
    def random_tree(cls, k=10, L=4, seed=0, stop_fraction=0.0, early_leaf_fraction=0.0, device=0, scoring=0, weighting=0):
        """Synthetic vocabulary of the ORBvoc shape (the real file is not shipped with the reference): breadth-first ids,
        random node descriptors, idf-like positive weights; optionally stopped words (weight 0) and leaves above level L."""
        rng = np.random.RandomState(seed)
        parent, leaf, level = [0], [0], [0]
        frontier = [0]
        for lv in range(1, L + 1):
            nxt = []
            for p in frontier:
                for _ in range(k):
                    parent.append(p)
                    level.append(lv)
                    is_leaf = lv == L or (lv >= 2 and rng.rand() < early_leaf_fraction)
                    leaf.append(1 if is_leaf else 0)
                    if not is_leaf:
                        nxt.append(len(parent) - 1)
            frontier = nxt
        n = len(parent)
        desc = rng.randint(0, 256, (n, 32)).astype(np.uint8)
        weight = np.where(np.array(leaf) > 0, rng.uniform(0.5, 9.0, n), 0.0)
        weight[(np.array(leaf) > 0) & (rng.rand(n) < stop_fraction)] = 0.0
        return cls(k, L, scoring, weighting, parent, leaf, desc, weight, device)

    def transform(self, d_desc, d_counts, levelsup=4, stream=None):
        """d_desc: CUDA uint8 [n, cap, 32], d_counts int32 [n] -> dict of CUDA tensors (see plvi_bow_transform)."""
        import torch
        n, cap = d_desc.shape[0], d_desc.shape[1]
        dev = d_desc.device
        o = {
            "word_id": torch.empty((n, cap), dtype=torch.int32, device=dev),
            "word_weight": torch.empty((n, cap), dtype=torch.float64, device=dev),
            "node_id": torch.empty((n, cap), dtype=torch.int32, device=dev),
            "bow_count": torch.empty(n, dtype=torch.int32, device=dev),
            "bow_words": torch.empty((n, cap), dtype=torch.int32, device=dev),
            "bow_values": torch.empty((n, cap), dtype=torch.float64, device=dev),
            "fv_count": torch.empty(n, dtype=torch.int32, device=dev),
            "fv_nodes": torch.empty((n, cap), dtype=torch.int32, device=dev),
            "fv_start": torch.empty((n, cap + 1), dtype=torch.int32, device=dev),
            "fv_features": torch.empty((n, cap), dtype=torch.int32, device=dev),
        }
        sp = int(stream.cuda_stream) if stream is not None else 0
        check(lib().plvi_bow_transform(self._handle(), sp, ptr(d_desc), ptr(d_counts), n, cap, int(levelsup), ptr(o["word_id"]),
                                       ptr(o["word_weight"]), ptr(o["node_id"]), ptr(o["bow_count"]), ptr(o["bow_words"]),
                                       ptr(o["bow_values"]), ptr(o["fv_count"]), ptr(o["fv_nodes"]), ptr(o["fv_start"]),
                                       ptr(o["fv_features"])))
        return o
