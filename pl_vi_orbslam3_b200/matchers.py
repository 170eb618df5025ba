"""Host-side mirrors of ORBmatcher (include/ORBmatcher.h:40-89) and LineMatcher
(include/LineMatcher.h:87-107) on top of the C ABI.

The reference's searches take Frame / MapPoint objects; the geometry (projection,
frustum tests) is host code there and stays with the caller here: a search receives the
train Frame as `FrameView` (mvKeysUn + mDescriptors + grid bounds) and one `QUERY_DTYPE`
record per candidate point in the reference's iteration order.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from .capi import GRID_DTYPE, KEYPOINT_DTYPE, QUERY_DTYPE, check, lib, ptr

FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48


def frame_grid(min_x, max_x, min_y, max_y):
    """Frame::mfGridElementWidthInv/HeightInv (src/Frame.cc:163-170), float32 like the reference."""
    g = np.zeros(1, GRID_DTYPE)
    g["min_x"], g["min_y"] = np.float32(min_x), np.float32(min_y)
    g["inv_w"] = np.float32(FRAME_GRID_COLS) / (np.float32(max_x) - np.float32(min_x))
    g["inv_h"] = np.float32(FRAME_GRID_ROWS) / (np.float32(max_y) - np.float32(min_y))
    return g


@dataclass
class FrameView:
    keys: np.ndarray            # KEYPOINT_DTYPE [n]  (mvKeysUn)
    desc: np.ndarray            # u8 [n,32]          (mDescriptors)
    grid: np.ndarray            # GRID_DTYPE [1]
    blocked: np.ndarray = None  # u8 [n]: mvpMapPoints[i] && Observations()>0


class _Matcher:
    def __init__(self, max_pairs=1, max_train=8192, max_query=8192, device=0, stream=None):
        self._h = C.c_void_p()
        check(lib().plvi_matcher_create(C.byref(self._h), max_pairs, max_train, max_query, device,
                                        ptr(stream) if stream else None))
        self.max_pairs = max_pairs
        self.device = device

    def sync(self):
        import torch
        torch.cuda.synchronize(self.device)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().plvi_matcher_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return lib().plvi_matcher_stream(self._h)

    def _hamming(self, a, b, shift25):
        a = np.ascontiguousarray(np.atleast_2d(a), np.uint8)
        b = np.ascontiguousarray(np.atleast_2d(b), np.uint8)
        out = np.empty(len(a), np.int32)
        check(lib().plvi_hamming256(self._h, ptr(a), ptr(b), len(a), shift25, ptr(out), 0))
        return out


class ORBmatcher(_Matcher):
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30   # src/ORBmatcher.cc:36-38

    def __init__(self, nnratio=0.6, checkOri=True, **kw):
        super().__init__(**kw)
        self.mfNNratio, self.mbCheckOrientation = float(nnratio), bool(checkOri)

    def DescriptorDistance(self, a, b):
        """a, b: u8 [n,32] (or [32]) -> int32 [n] (src/ORBmatcher.cc:2350-2366)."""
        return self._hamming(a, b, 0)

    def search_batch(self, mode, frames, queries, qdescs, th):
        """Batched guided search: frames = [FrameView], queries = [QUERY_DTYPE array],
        qdescs = [u8 [nq,32]].  Returns (match_train list, match_query list, nmatches, queries out)."""
        P = len(frames)
        T = max(max(len(f.keys) for f in frames), 1)
        Q = max(max(len(q) for q in queries), 1)
        keys = np.zeros((P, T), KEYPOINT_DTYPE)
        desc = np.zeros((P, T, 32), np.uint8)
        blocked = np.zeros((P, T), np.uint8) if any(f.blocked is not None for f in frames) else None
        qs = np.zeros((P, Q), QUERY_DTYPE)
        qd = np.zeros((P, Q, 32), np.uint8)
        tc = np.array([len(f.keys) for f in frames], np.int32)
        qc = np.array([len(q) for q in queries], np.int32)
        for i, f in enumerate(frames):
            keys[i, :tc[i]] = f.keys
            desc[i, :tc[i]] = f.desc
            if blocked is not None and f.blocked is not None:
                blocked[i, :tc[i]] = f.blocked
            qs[i, :qc[i]] = queries[i]
            qd[i, :qc[i]] = qdescs[i]
        mt = np.empty((P, T), np.int32)
        mq = np.empty((P, Q), np.int32)
        nm = np.empty(P, np.int32)
        grid = np.ascontiguousarray(frames[0].grid)
        check(lib().plvi_search_by_projection(
            self._h, mode, P, ptr(keys), ptr(desc), ptr(blocked), ptr(tc), T, ptr(grid), ptr(qs), ptr(qd),
            ptr(qc), Q, th, self.mfNNratio, int(self.mbCheckOrientation), ptr(mt), ptr(mq), ptr(nm), 0))
        return ([mt[i, :tc[i]] for i in range(P)], [mq[i, :qc[i]] for i in range(P)], nm,
                [qs[i, :qc[i]] for i in range(P)])

    def SearchByProjection(self, CurrentFrame, queries, qdesc, mappoints=False):
        """Frame-to-frame (src/ORBmatcher.cc:1962) or, with mappoints=True, Frame vs local
        map points (src/ORBmatcher.cc:44).  Returns (nmatches, match_train, match_query)."""
        mode = 1 if mappoints else 0
        mt, mq, nm, _ = self.search_batch(mode, [CurrentFrame], [queries], [qdesc], self.TH_HIGH)
        return int(nm[0]), mt[0], mq[0]

    def SearchByProjection_KF(self, KF, queries, qdesc, ratioHamming=1.0):
        """Search of SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vpMatched, th, ratioHamming) and its
        vpPointsKFs overload (src/ORBmatcher.cc:473-704): the reference skips vpMatched[idx] != NULL and claims
        vpMatched[bestIdx] while it iterates, i.e. the sequential frame search without the rotation check.
        KF.blocked = vpMatched[i] != NULL on entry; queries: u, v, radius = th * mvScaleFactors[nPredictedLevel],
        levels nPredictedLevel-1 .. nPredictedLevel.  Returns (nmatches, match_train, match_query)."""
        ori, self.mbCheckOrientation = self.mbCheckOrientation, False
        try:
            mt, mq, nm, _ = self.search_batch(0, [KF], [queries], [qdesc], int(np.floor(self.TH_LOW * ratioHamming)))
        finally:
            self.mbCheckOrientation = ori
        return int(nm[0]), mt[0], mq[0]

    def SearchByBoW(self, F, group_items, queries, qdesc):
        """Descriptor part of SearchByBoW(KeyFrame*, Frame&, ...) (src/ORBmatcher.cc:269-471).
        group_items: int32 frame feature indices grouped by vocabulary node; queries: QUERY_DTYPE with
        min_level/max_level = [start, end) of the node group.  Returns (nmatches, match_train, match_query)."""
        n, nq = len(F.keys), len(queries)
        T, Q, I = max(n, 1), max(nq, 1), max(len(group_items), 1)
        keys = np.zeros(T, KEYPOINT_DTYPE); keys[:n] = F.keys
        desc = np.zeros((T, 32), np.uint8); desc[:n] = F.desc
        items = np.zeros(I, np.int32); items[:len(group_items)] = group_items
        qs = np.zeros(Q, QUERY_DTYPE); qs[:nq] = queries
        qd = np.zeros((Q, 32), np.uint8); qd[:nq] = qdesc
        tc, qc = np.array([n], np.int32), np.array([nq], np.int32)
        mt, mq, nm = np.empty(T, np.int32), np.empty(Q, np.int32), np.empty(1, np.int32)
        check(lib().plvi_search_by_bow(self._h, 1, ptr(keys), ptr(desc), ptr(tc), T, ptr(items), I, ptr(qs), ptr(qd),
                                       ptr(qc), Q, self.TH_LOW, self.mfNNratio, int(self.mbCheckOrientation),
                                       ptr(mt), ptr(mq), ptr(nm), 0))
        return int(nm[0]), mt[:n], mq[:nq]

    def SearchByBoW_KF(self, KF2, has_mappoint2, group_items, queries, qdesc):
        """Descriptor part of SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (src/ORBmatcher.cc:823-963): KF2 is the
        searched side, has_mappoint2[i] = vpMapPoints2[i] is a good map point.  Returns (nmatches, match_query) with
        match_query[q] = KF2 feature matched to query q or -1."""
        n, nq = len(KF2.keys), len(queries)
        T, Q, I = max(n, 1), max(nq, 1), max(len(group_items), 1)
        keys = np.zeros(T, KEYPOINT_DTYPE); keys[:n] = KF2.keys
        desc = np.zeros((T, 32), np.uint8); desc[:n] = KF2.desc
        blk = np.ones(T, np.uint8); blk[:n] = ~np.asarray(has_mappoint2, bool)
        items = np.zeros(I, np.int32); items[:len(group_items)] = group_items
        qs = np.zeros(Q, QUERY_DTYPE); qs[:nq] = queries
        qd = np.zeros((Q, 32), np.uint8); qd[:nq] = qdesc
        tc, qc = np.array([n], np.int32), np.array([nq], np.int32)
        mt, mq, nm = np.empty(T, np.int32), np.empty(Q, np.int32), np.empty(1, np.int32)
        check(lib().plvi_search_by_bow_kf(self._h, 1, ptr(keys), ptr(desc), ptr(blk), ptr(tc), T, ptr(items), I, ptr(qs),
                                          ptr(qd), ptr(qc), Q, self.TH_LOW, self.mfNNratio, int(self.mbCheckOrientation),
                                          ptr(mt), ptr(mq), ptr(nm), 0))
        return int(nm[0]), mq[:nq]

    def SearchForTriangulation(self, KF2, has_mappoint2, group_items, queries, qdesc, F12, ep, scale_factors2, level_sigma2_2,
                               coarse=False):
        """Descriptor + epipolar part of SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, false, bCoarse)
        (src/ORBmatcher.cc:965-1206, mono pinhole).  queries: one per KF1 feature without map point of a common node
        (u, v = kp1.pt).  Returns (nmatches, match_query)."""
        import torch
        from .capi import EPIPOLAR_DTYPE
        n, nq = len(KF2.keys), len(queries)
        T, Q, I = max(n, 1), max(nq, 1), max(len(group_items), 1)
        keys = np.zeros(T, KEYPOINT_DTYPE); keys[:n] = KF2.keys
        desc = np.zeros((T, 32), np.uint8); desc[:n] = KF2.desc
        blk = np.ones(T, np.uint8); blk[:n] = np.asarray(has_mappoint2, bool)
        items = np.zeros(I, np.int32); items[:len(group_items)] = group_items
        qs = np.zeros(Q, QUERY_DTYPE); qs[:nq] = queries
        qd = np.zeros((Q, 32), np.uint8); qd[:nq] = qdesc
        g = np.zeros(1, EPIPOLAR_DTYPE)
        g["F12"][0] = np.asarray(F12, np.float32).reshape(9)
        g["ep_x"], g["ep_y"] = np.float32(ep[0]), np.float32(ep[1])
        g["scale_factors"][0, :len(scale_factors2)] = scale_factors2
        g["level_sigma2"][0, :len(level_sigma2_2)] = level_sigma2_2
        g["coarse"], g["check_epipole"] = int(coarse), 1
        dev = torch.device("cuda", self.device)
        def up(a):
            return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
        t = [up(x) for x in (keys, desc, blk, np.array([n], np.int32), items, qs, qd, np.array([nq], np.int32), g)]
        mq = torch.empty(Q, dtype=torch.int32, device=dev)
        nm = torch.empty(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        check(lib().plvi_search_for_triangulation(self._h, 1, ptr(t[0]), ptr(t[1]), ptr(t[2]), ptr(t[3]), T, ptr(t[4]), I, ptr(t[5]),
                                                  ptr(t[6]), ptr(t[7]), Q, ptr(t[8]), self.TH_LOW, int(self.mbCheckOrientation),
                                                  ptr(mq), ptr(nm)))
        self.sync()
        return int(nm.cpu().numpy()[0]), mq.cpu().numpy()[:nq]

    def SearchInRadius(self, KF, queries, qdesc, inv_level_sigma2, chi2=5.99, th_dist=None):
        """The per-map-point search of Fuse(pKF, vpMapPoints, th) (src/ORBmatcher.cc:1399-1610; chi2 = 5.99, TH_LOW),
        Fuse(pKF, Scw, ...) (:1612; chi2 = 0), SearchBySim3 (:1736; chi2 = 0, TH_HIGH, once per direction)
        [SearchByProjection(pKF, Scw, ...) claims features while iterating: SearchByProjection_KF].  queries: one per projected map point
        (u, v, radius, min_level = nPredictedLevel - 1, max_level = nPredictedLevel, flags bit0 = skipped).
        Returns (nfound, best_idx, best_dist)."""
        import torch
        th_dist = self.TH_LOW if th_dist is None else int(th_dist)
        n, nq = len(KF.keys), len(queries)
        T, Q = max(n, 1), max(nq, 1)
        keys = np.zeros(T, KEYPOINT_DTYPE); keys[:n] = KF.keys
        desc = np.zeros((T, 32), np.uint8); desc[:n] = KF.desc
        qs = np.zeros(Q, QUERY_DTYPE); qs[:nq] = queries
        qd = np.zeros((Q, 32), np.uint8); qd[:nq] = qdesc
        s2 = np.zeros(16, np.float32); s2[:len(inv_level_sigma2)] = inv_level_sigma2
        grid = np.ascontiguousarray(KF.grid)
        dev = torch.device("cuda", self.device)
        def up(a):
            return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
        t = [up(x) for x in (keys, desc, np.array([n], np.int32), qs, qd, np.array([nq], np.int32))]
        bi = torch.empty(Q, dtype=torch.int32, device=dev)
        bd = torch.empty(Q, dtype=torch.int32, device=dev)
        nf = torch.empty(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        check(lib().plvi_search_in_radius(self._h, 1, ptr(t[0]), ptr(t[1]), ptr(t[2]), T, ptr(grid), ptr(t[3]), ptr(t[4]),
                                          ptr(t[5]), Q, ptr(s2), float(chi2), th_dist, ptr(bi), ptr(bd), ptr(nf)))
        self.sync()
        return int(nf.cpu().numpy()[0]), bi.cpu().numpy()[:nq], bd.cpu().numpy()[:nq]

    @staticmethod
    def init_queries(F1_keys, vbPrevMatched, windowSize):
        n1 = len(F1_keys)
        q = np.zeros(n1, QUERY_DTYPE)
        q["u"], q["v"] = vbPrevMatched[:, 0], vbPrevMatched[:, 1]
        q["radius"] = windowSize
        q["min_level"], q["max_level"] = 0, 0
        q["angle"] = F1_keys["angle"]
        q["flags"] = (F1_keys["octave"] > 0).astype(np.int32)
        return q

    def SearchForInitialization(self, F1_keys, F1_desc, F2, vbPrevMatched, windowSize=10):
        """src/ORBmatcher.cc:706-820.  Returns (nmatches, vnMatches12, updated vbPrevMatched)."""
        q = self.init_queries(F1_keys, vbPrevMatched, windowSize)
        mt, mq, nm, qs = self.search_batch(2, [F2], [q], [F1_desc], self.TH_LOW)
        prev = np.stack([qs[0]["u"], qs[0]["v"]], axis=1)
        return int(nm[0]), mq[0], prev


class LineMatcher(_Matcher):
    TH_HIGH, TH_LOW = 100, 50   # src/LineMatcher.cpp:37-38

    def FuseSearch(self, keylines, desc, queries, qdesc, flags=None):
        """The per-map-line search of LineMatcher::Fuse (src/LineMatcher.cpp:373-485): keylines / desc = the keyframe's
        mvKeys_Line / mDescriptors_l, queries [nq, 6] f32 = projected endpoints u1, v1, u2, v2, radius, predicted level.
        Returns (nfound, best_idx, best_dist)."""
        import torch
        from .capi import KEYLINE_DTYPE
        n, nq = len(keylines), len(queries)
        T, Q = max(n, 1), max(nq, 1)
        kl = np.zeros(T, KEYLINE_DTYPE); kl[:n] = keylines
        d = np.zeros((T, 32), np.uint8); d[:n] = desc
        q = np.zeros((Q, 6), np.float32); q[:nq] = np.asarray(queries, np.float32).reshape(-1, 6)
        qd = np.zeros((Q, 32), np.uint8); qd[:nq] = qdesc
        fl = np.zeros(Q, np.uint8)
        if flags is not None:
            fl[:nq] = flags
        dev = torch.device("cuda", self.device)
        def up(a):
            return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)
        t = [up(x) for x in (kl, d, np.array([n], np.int32), q, fl, qd, np.array([nq], np.int32))]
        bi = torch.empty(Q, dtype=torch.int32, device=dev)
        bd = torch.empty(Q, dtype=torch.int32, device=dev)
        nf = torch.empty(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        check(lib().plvi_line_fuse_search(self._h, 1, ptr(t[0]), ptr(t[1]), ptr(t[2]), T, ptr(t[3]), ptr(t[4]), ptr(t[5]),
                                          ptr(t[6]), Q, self.TH_LOW, ptr(bi), ptr(bd), ptr(nf)))
        self.sync()
        return int(nf.cpu().numpy()[0]), bi.cpu().numpy()[:nq], bd.cpu().numpy()[:nq]

    def match_batch(self, pairs, nnr, mutual=True):
        P = len(pairs)
        S1 = max(max(len(a) for a, _ in pairs), 1)
        S2 = max(max(len(b) for _, b in pairs), 1)
        d1 = np.zeros((P, S1, 32), np.uint8)
        d2 = np.zeros((P, S2, 32), np.uint8)
        n1 = np.array([len(a) for a, _ in pairs], np.int32)
        n2 = np.array([len(b) for _, b in pairs], np.int32)
        for i, (a, b) in enumerate(pairs):
            d1[i, :n1[i]] = a
            d2[i, :n2[i]] = b
        m12 = np.empty((P, S1), np.int32)
        nm = np.empty(P, np.int32)
        check(lib().plvi_line_match(self._h, P, ptr(d1), ptr(n1), S1, ptr(d2), ptr(n2), S2, float(nnr),
                                    int(mutual), ptr(m12), ptr(nm), 0))
        return [m12[i, :n1[i]] for i in range(P)], nm

    def match(self, desc1, desc2, nnr):
        """LineMatcher::match(desc1, desc2, nnr, matches_12) -> (count, matches_12)."""
        m, nm = self.match_batch([(desc1, desc2)], nnr, True)
        return int(nm[0]), m[0]

    def matchNNR(self, desc1, desc2, nnr):
        m, nm = self.match_batch([(desc1, desc2)], nnr, False)
        return int(nm[0]), m[0]

    def distance(self, a, b):
        return self._hamming(a, b, 0)

    def DescriptorDistance(self, a, b):
        """The reference's >>25 variant (src/LineMatcher.cpp:487-499), kept bit-compatible."""
        return self._hamming(a, b, 1)

    def _match_mad_batch(self, pairs, factor, masks=None):
        import torch
        P = len(pairs)
        S1 = max(max(len(a) for a, _ in pairs), 1)
        S2 = max(max(len(b) for _, b in pairs), 1)
        d1 = np.zeros((P, S1, 32), np.uint8)
        d2 = np.zeros((P, S2, 32), np.uint8)
        h1 = np.zeros((P, S1), np.uint8)
        h2 = np.zeros((P, S2), np.uint8)
        n1 = np.array([len(a) for a, _ in pairs], np.int32)
        n2 = np.array([len(b) for _, b in pairs], np.int32)
        for i, (a, b) in enumerate(pairs):
            d1[i, :n1[i]] = a
            d2[i, :n2[i]] = b
            if masks is not None:
                h1[i, :n1[i]], h2[i, :n2[i]] = masks[i]
        dev = torch.device("cuda", self.device)
        t = [torch.from_numpy(x).to(dev) for x in (d1, n1, d2, n2, h1, h2)]
        m12 = torch.empty((P, S1), dtype=torch.int32, device=dev)
        nm = torch.empty(P, dtype=torch.int32, device=dev)
        mad = torch.empty((P, 2), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        check(lib().plvi_line_match_mad(self._h, P, ptr(t[0]), ptr(t[1]), S1, ptr(t[2]), ptr(t[3]), S2,
                                        ptr(t[4]) if masks is not None else None, ptr(t[5]) if masks is not None else None,
                                        float(factor), ptr(m12), ptr(nm), ptr(mad)))
        self.sync()
        m12, nm, mad = m12.cpu().numpy(), nm.cpu().numpy(), mad.cpu().numpy()
        return [m12[i, :n1[i]] for i in range(P)], nm, mad

    def matchGrid_batch(self, pairs, inv_width, inv_height, grid_rows=48, grid_cols=64, window=(7, 0, 2, 2)):
        """The stereo line search of Frame::ComputeStereoMatches_Lines (src/Frame.cc:1421-1448): grid fill along
        LineIterator + LineMatcher::matchGrid (src/LineMatcher.cpp:191-272) for a batch of stereo pairs.
        pairs = [(seg_left [n1,4], desc_left [n1,32], seg_right [n2,4], desc_right [n2,32]), ...] with
        seg = (startPointX, startPointY, endPointX, endPointY); window = (left, right, up, down) grid cells.
        Returns ([matches_12 per pair], counts)."""
        import torch
        P = len(pairs)
        S1 = max(max(len(p[1]) for p in pairs), 1)
        S2 = max(max(len(p[3]) for p in pairs), 1)
        s1 = np.zeros((P, S1, 4), np.float32)
        s2 = np.zeros((P, S2, 4), np.float32)
        d1 = np.zeros((P, S1, 32), np.uint8)
        d2 = np.zeros((P, S2, 32), np.uint8)
        n1 = np.array([len(p[1]) for p in pairs], np.int32)
        n2 = np.array([len(p[3]) for p in pairs], np.int32)
        for i, (a, da, b, db) in enumerate(pairs):
            if n1[i]:
                s1[i, :n1[i]] = np.asarray(a, np.float32).reshape(-1, 4)
                d1[i, :n1[i]] = da
            if n2[i]:
                s2[i, :n2[i]] = np.asarray(b, np.float32).reshape(-1, 4)
                d2[i, :n2[i]] = db
        dev = torch.device("cuda", self.device)
        t = [torch.from_numpy(x).to(dev) for x in (s1, d1, n1, s2, d2, n2)]
        m12 = torch.empty((P, S1), dtype=torch.int32, device=dev)
        nm = torch.empty(P, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        wl, wr, wu, wd = (int(v) for v in window)
        check(lib().plvi_line_match_grid(self._h, P, ptr(t[0]), ptr(t[1]), ptr(t[2]), S1, ptr(t[3]), ptr(t[4]), ptr(t[5]), S2,
                                         float(inv_width), float(inv_height), int(grid_rows), int(grid_cols), wl, wr, wu, wd,
                                         ptr(m12), ptr(nm)))
        self.sync()
        m12, nm = m12.cpu().numpy(), nm.cpu().numpy()
        return [m12[i, :n1[i]] for i in range(P)], nm

    def matchGrid(self, seg_left, desc_left, seg_right, desc_right, inv_width, inv_height, **kw):
        """int LineMatcher::matchGrid(lines1, desc1, grid, desc2, directions2, w, matches_12) with the grid and the
        directions built from the right keylines as Frame::ComputeStereoMatches_Lines does -> (count, matches_12)."""
        m, nm = self.matchGrid_batch([(seg_left, desc_left, seg_right, desc_right)], inv_width, inv_height, **kw)
        return int(nm[0]), m[0]

    def matchGrid_host(self, seg_left, desc_left, seg_right, desc_right, inv_width, inv_height, grid_rows=48,
                       grid_cols=64, window=(7, 0, 2, 2)):
        """The same search through plvi_line_match_grid_host (numpy buffers in and out, one stereo pair): the call the
        C++ shim makes from Frame::ComputeStereoMatches_Lines -> (count, matches_12)."""
        s1 = np.ascontiguousarray(seg_left, np.float32).reshape(-1, 4)
        s2 = np.ascontiguousarray(seg_right, np.float32).reshape(-1, 4)
        d1 = np.ascontiguousarray(desc_left, np.uint8).reshape(-1, 32)
        d2 = np.ascontiguousarray(desc_right, np.uint8).reshape(-1, 32)
        m12 = np.full(max(len(d1), 1), -1, np.int32)
        nm = np.zeros(1, np.int32)
        check(lib().plvi_line_match_grid_host(self._h, ptr(s1) if len(d1) else None, ptr(d1) if len(d1) else None, len(d1),
                                              ptr(s2) if len(d2) else None, ptr(d2) if len(d2) else None, len(d2),
                                              float(inv_width), float(inv_height), int(grid_rows), int(grid_cols),
                                              *[int(v) for v in window], ptr(m12), ptr(nm)))
        return int(nm[0]), m12[:len(d1)]

    def stereo_depth_host(self, seg_left, seg_right, matches12, mbf, seg_left_un=None):
        """What follows the search in Frame::ComputeStereoMatches_Lines (src/Frame.cc:1453-1500) through
        plvi_line_stereo_depth_host -> (count, mvDisparity_l [n, 2], mvDepth_l [n, 2], mvle_l [n, 3])."""
        s1 = np.ascontiguousarray(seg_left, np.float32).reshape(-1, 4)
        s2 = np.ascontiguousarray(seg_right, np.float32).reshape(-1, 4)
        su = s1 if seg_left_un is None else np.ascontiguousarray(seg_left_un, np.float32).reshape(-1, 4)
        m12 = np.ascontiguousarray(matches12, np.int32)
        n1 = len(s1)
        disp, dep = np.zeros((max(n1, 1), 2), np.float32), np.zeros((max(n1, 1), 2), np.float32)
        le = np.zeros((max(n1, 1), 3), np.float64)
        nd = np.zeros(1, np.int32)
        check(lib().plvi_line_stereo_depth_host(self._h, ptr(s1) if n1 else None, n1, ptr(s2) if len(s2) else None, len(s2),
                                                ptr(m12) if n1 else None, ptr(su) if n1 else None, float(mbf), ptr(disp), ptr(dep),
                                                ptr(le), ptr(nd)))
        return int(nd[0]), disp[:n1], dep[:n1], le[:n1]

    def SerachForInitialize(self, desc_initial, desc_current):
        """int LineMatcher::SerachForInitialize(Frame&, Frame&, vector<pair<int,int>>&) (src/LineMatcher.cpp:113-141)
        on the two frames' line descriptors -> (count, [(qdx, tdx), ...])."""
        m, nm, _ = self._match_mad_batch([(desc_initial, desc_current)], 0.5)
        return int(nm[0]), [(i, int(t)) for i, t in enumerate(m[0]) if t >= 0]

    def SearchForTriangulation(self, desc_kf1, desc_kf2, has_line1, has_line2):
        """int LineMatcher::SearchForTriangulation(KeyFrame*, KeyFrame*, vector<pair<size_t,size_t>>&) (:143-171);
        has_lineN[i] = pKFN->GetMapLine(i) != NULL."""
        m, nm, _ = self._match_mad_batch([(desc_kf1, desc_kf2)], 0.1, [(has_line1, has_line2)])
        return int(nm[0]), [(i, int(t)) for i, t in enumerate(m[0]) if t >= 0]


def compute_distinctive_descriptors(d_desc, d_counts, stream=None):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:330-402) for a batch of map points.
    d_desc: CUDA uint8 [n_points, max_obs, 32], d_counts int32 [n_points] -> (best_idx int32 [n], mDescriptor uint8 [n, 32])."""
    import torch
    n, cap = d_desc.shape[0], d_desc.shape[1]
    idx = torch.empty(n, dtype=torch.int32, device=d_desc.device)
    best = torch.zeros((n, 32), dtype=torch.uint8, device=d_desc.device)
    sp = int(stream.cuda_stream) if stream is not None else 0
    check(lib().plvi_distinctive_descriptors(sp, ptr(d_desc), ptr(d_counts), n, cap, ptr(idx), ptr(best)))
    return idx, best
