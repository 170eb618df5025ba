"""Deterministic synthetic EuRoC-shaped frames (inputs C1-C5).

numpy only, so the same frames can be produced in the build container and on the
GPU box.  frame_euroc(seed) = smooth noise background + filled rectangles (60 large, 420 small) + thick
line segments + filled discs, 5x5 blur, N(0, 2^2) sensor noise.
"""
import numpy as np


def _blur5(img):
    k = np.array([1, 4, 6, 4, 1], np.float32) / 16.0
    p = np.pad(img, 2, mode="reflect")
    h = sum(k[i] * p[:, i:i + img.shape[1]] for i in range(5))
    v = sum(k[i] * h[i:i + img.shape[0], :] for i in range(5))
    return v


def frame_euroc(seed, w=752, h=480):
    rng = np.random.RandomState(seed)
    gw, gh = 47, 30
    coarse = rng.uniform(40, 200, (gh, gw)).astype(np.float32)
    xs = np.linspace(0, gw - 1, w, dtype=np.float32)
    ys = np.linspace(0, gh - 1, h, dtype=np.float32)
    x0 = np.minimum(xs.astype(np.int32), gw - 2)
    y0 = np.minimum(ys.astype(np.int32), gh - 2)
    fx = (xs - x0)[None, :]
    fy = (ys - y0)[:, None]
    c00 = coarse[y0][:, x0]
    c01 = coarse[y0][:, x0 + 1]
    c10 = coarse[y0 + 1][:, x0]
    c11 = coarse[y0 + 1][:, x0 + 1]
    img = (c00 * (1 - fx) + c01 * fx) * (1 - fy) + (c10 * (1 - fx) + c11 * fx) * fy
    for _ in range(60):
        rw, rh = rng.randint(20, 161, 2)
        x, y = rng.randint(-40, w - 10), rng.randint(-40, h - 10)
        g = rng.randint(0, 256)
        img[max(y, 0):max(y + rh, 0), max(x, 0):max(x + rw, 0)] = g
    for _ in range(420):  # small high-contrast blocks: corner-rich texture
        rw, rh = rng.randint(5, 31, 2)
        x, y = rng.randint(0, w - 5), rng.randint(0, h - 5)
        g = rng.randint(0, 256)
        img[y:y + rh, x:x + rw] = g
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for _ in range(40):
        x1, x2 = rng.uniform(0, w, 2)
        y1, y2 = rng.uniform(0, h, 2)
        t = rng.randint(1, 4)
        g = rng.randint(0, 256)
        dx, dy = x2 - x1, y2 - y1
        L2 = dx * dx + dy * dy + 1e-6
        xa, xb = int(max(min(x1, x2) - t, 0)), int(min(max(x1, x2) + t + 1, w))
        ya, yb = int(max(min(y1, y2) - t, 0)), int(min(max(y1, y2) + t + 1, h))
        sx, sy = xx[ya:yb, xa:xb], yy[ya:yb, xa:xb]
        u = np.clip(((sx - x1) * dx + (sy - y1) * dy) / L2, 0, 1)
        d2 = (sx - (x1 + u * dx)) ** 2 + (sy - (y1 + u * dy)) ** 2
        img[ya:yb, xa:xb][d2 <= (t * 0.5 + 0.25) ** 2] = g
    for _ in range(30):
        r = rng.randint(5, 41)
        cx, cy = rng.randint(0, w), rng.randint(0, h)
        g = rng.randint(0, 256)
        xa, xb = max(cx - r, 0), min(cx + r + 1, w)
        ya, yb = max(cy - r, 0), min(cy + r + 1, h)
        m = (xx[ya:yb, xa:xb] - cx) ** 2 + (yy[ya:yb, xa:xb] - cy) ** 2 <= r * r
        img[ya:yb, xa:xb][m] = g
    img = _blur5(img)
    img = img + rng.normal(0.0, 2.0, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def warp_pair(seed, w=752, h=480, angle_deg=1.5, tx=6.0, ty=-4.0, scale=1.01):
    """C3: (frame, affinely warped frame with fresh noise).  Returns (f1, f2, A) with
    A the 2x3 forward map p2 = A @ [p1,1] (bilinear sampling, replicate border)."""
    f1 = frame_euroc(seed, w, h)
    a = np.deg2rad(angle_deg)
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    R = scale * np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
    t = np.array([cx, cy]) - R @ np.array([cx, cy]) + np.array([tx, ty])
    A = np.hstack([R, t[:, None]])
    Ri = np.linalg.inv(R)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    sx = Ri[0, 0] * (xx - t[0]) + Ri[0, 1] * (yy - t[1])
    sy = Ri[1, 0] * (xx - t[0]) + Ri[1, 1] * (yy - t[1])
    sx = np.clip(sx, 0, w - 1)
    sy = np.clip(sy, 0, h - 1)
    x0 = np.minimum(sx.astype(np.int32), w - 2)
    y0 = np.minimum(sy.astype(np.int32), h - 2)
    fx, fy = sx - x0, sy - y0
    g = f1.astype(np.float64)
    v = (g[y0, x0] * (1 - fx) + g[y0, x0 + 1] * fx) * (1 - fy) + \
        (g[y0 + 1, x0] * (1 - fx) + g[y0 + 1, x0 + 1] * fx) * fy
    rng = np.random.RandomState(seed + 10 ** 6)
    v = v + rng.normal(0.0, 2.0, v.shape)
    f2 = np.clip(np.rint(v), 0, 255).astype(np.uint8)
    return f1, f2, A


def frame_batch(n, w=752, h=480, base_seed=0, distinct=16):
    """n frames for throughput runs: `distinct` generated frames, the rest are
    cyclic shifts of them (cheap, still all different images)."""
    base = [frame_euroc(base_seed + i, w, h) for i in range(min(distinct, n))]
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        b = base[i % len(base)]
        k = i // len(base)
        out[i] = np.roll(b, (3 * k, 5 * k), axis=(0, 1)) if k else b
    return out


# ---- whole benchmark batches: every frame generated from its own seed (seeds 0...4095), in
# ---- parallel worker processes, cached on local disk so that back-to-back bench runs on one box generate once ----------
def _gen_frame(a):
    return frame_euroc(*a)


def _gen_pair(a):
    f1, f2, _ = warp_pair(*a)
    return np.stack([f1, f2])


def warp_affine(w=752, h=480, angle_deg=1.5, tx=6.0, ty=-4.0, scale=1.01):
    """The 2x3 forward map of warp_pair (p2 = A @ [p1, 1]) without generating frames."""
    a = np.deg2rad(angle_deg)
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    R = scale * np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
    t = np.array([cx, cy]) - R @ np.array([cx, cy]) + np.array([tx, ty])
    return np.hstack([R, t[:, None]])


def _batch(kind, n, w, h, base_seed, workers, cache, budget_s=None):
    import os
    import tempfile
    from pathlib import Path
    unit = 2 if kind == "pairs" else 1
    nunits = (n + unit - 1) // unit
    workers = max(1, min(workers or (os.cpu_count() or 1), nunits))
    # generation budget: a frame costs ~0.1 s, a pair ~0.16 s of one core.  When the workers of this process cannot make
    # every unit within budget_s, the first `distinct` units are generated from their own seeds and the rest are cyclic
    # shifts of them (both frames of a pair by the same shift, <= 1 px off the pair's affine map for the shifts used)
    distinct = nunits
    if budget_s:
        distinct = max(1, min(nunits, int(workers * budget_s / (0.16 if unit == 2 else 0.1))))
    path = None
    if cache:
        d = Path(os.environ.get("PLVI_SYNTH_CACHE", Path(tempfile.gettempdir()) / "plvi_synth_cache"))
        path = d / f"{kind}_{w}x{h}_s{base_seed}_n{n}_d{distinct}.npy"
        try:
            if path.exists():
                a = np.load(path)
                if a.shape == (n, h, w) and a.dtype == np.uint8:
                    return a, distinct * unit
        except Exception:
            pass
    if kind == "pairs":
        jobs, fn = [(base_seed + p, w, h) for p in range(distinct)], _gen_pair
    else:
        jobs, fn = [(base_seed + i, w, h) for i in range(distinct)], _gen_frame
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:     # call before CUDA is initialised in this process
            parts = pool.map(fn, jobs, chunksize=max(1, len(jobs) // (8 * workers)))
    else:
        parts = [fn(j) for j in jobs]
    base = np.concatenate(parts) if kind == "pairs" else np.stack(parts)
    if distinct < nunits:
        out = np.empty((nunits * unit, h, w), np.uint8)
        nb = len(base)
        for u in range(nunits):
            k = u // distinct
            src = base[(u % distinct) * unit:(u % distinct) * unit + unit]
            out[u * unit:(u + 1) * unit] = np.roll(src, (3 * k, 5 * k), axis=(1, 2)) if k else src
        base = out
    out = base[:n]
    if path is not None:
        try:
            path.parent.mkdir(parents=True, exist_ok=True)
            tmp = path.with_suffix(f".{os.getpid()}.tmp.npy")
            np.save(tmp, out)
            os.replace(tmp, path)
        except Exception:
            pass
    return out, distinct * unit


def seq_batch(n, w=752, h=480, base_seed=0, workers=0, cache=True, budget_s=None, with_info=False):
    """C1 / C4 / C5: frames frame_euroc(base_seed) ... frame_euroc(base_seed + n - 1), [n, h, w] u8."""
    out, distinct = _batch("seq", n, w, h, base_seed, workers, cache, budget_s)
    return (out, distinct) if with_info else out


def pair_batch(n, w=752, h=480, base_seed=0, workers=0, cache=True, budget_s=None, with_info=False):
    """C3: n frames = n / 2 pairs; frame 2p = frame_euroc(base_seed + p), frame 2p + 1 = its warp (warp_pair, map
    warp_affine()), [n, h, w] u8."""
    out, distinct = _batch("pairs", n, w, h, base_seed, workers, cache, budget_s)
    return (out, distinct) if with_info else out
