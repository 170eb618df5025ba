"""CPU model of the schedule of k_search (csrc/match_kernels.cu): warps evaluate queries AHEAD of the in-order commit,
against whatever mixture of older and newer claim flags they happen to read; the commit re-evaluates a query only when
its best or second candidate is no longer eligible.  The reference's searches claim features greedily in query order
(src/ORBmatcher.cc: SearchByProjection, SearchForInitialization), so the schedule is exact iff that rule reproduces the
sequential result.  It does because eligibility only shrinks: a blocked feature never unblocks, a claimed distance only
decreases.  Checked here on random instances for both kinds of claims."""
import numpy as np


def _top2(cands, dist_row, ok):
    """two smallest (distance, position in the candidate list) among the eligible candidates"""
    best = second = (1 << 30, -1, -1)
    for pos, t in enumerate(cands):
        if not ok(t, dist_row[t]):
            continue
        key = (int(dist_row[t]), pos, int(t))
        if key < best:
            best, second = key, best
        elif key < second:
            second = key
    return best, second


def _run(rng, nq, nt, mode, early):
    D = rng.randint(0, 120, (nq, nt))
    cands = [rng.permutation(nt)[: rng.randint(0, 12)] for _ in range(nq)]
    th, ratio = 80, 0.9
    blk_hist = [np.zeros(nt, bool)]            # claim flags after every commit (mode "block")
    md_hist = [np.full(nt, 1 << 20)]           # claimed distances after every commit (mode "dist")
    owner = np.full(nt, -1)
    match = np.full(nq, -1)
    for q in range(nq):
        blk, md = blk_hist[-1].copy(), md_hist[-1].copy()
        now = (lambda t, d: not blk[t]) if mode == "block" else (lambda t, d: not (md[t] <= d))
        if early:
            # every candidate is judged against the state of some earlier moment (a different one per candidate)
            lag = {int(t): rng.randint(0, min(q, 25) + 1) for t in cands[q]}
            if mode == "block":
                seen = lambda t, d: not blk_hist[q - lag[int(t)]][t]
            else:
                seen = lambda t, d: not (md_hist[q - lag[int(t)]][t] <= d)
            best, second = _top2(cands[q], D[q], seen)
            stale = (best[2] >= 0 and not now(best[2], best[0])) or (second[2] >= 0 and not now(second[2], second[0]))
            if stale:
                best, second = _top2(cands[q], D[q], now)
        else:
            best, second = _top2(cands[q], D[q], now)
        if best[2] >= 0 and best[0] <= th:
            t = best[2]
            if mode == "block":
                match[q] = t
                blk[t] = True
            elif best[0] < ratio * (second[0] if second[2] >= 0 else (1 << 30)):
                if owner[t] >= 0:
                    match[owner[t]] = -1
                owner[t] = q
                match[q] = t
                md[t] = best[0]
        blk_hist.append(blk)
        md_hist.append(md)
    return match


def test_early_evaluation_with_stale_check_equals_the_sequential_search():
    for seed in range(30):
        for mode in ("block", "dist"):
            a = _run(np.random.RandomState(seed), 120, 60, mode, early=False)
            b = _run(np.random.RandomState(seed), 120, 60, mode, early=True)
            # (both runs build the distances and the candidate lists first, from the same seed: the same instance)
            assert np.array_equal(a, b), (seed, mode)
