"""TEST INFRASTRUCTURE: ctypes front for the CPU oracle (oracle/*.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  Parity pinning status: see oracle/oracle_common.h.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = _DIR / "libplvi_oracle.so"

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)
assert KEYPOINT_DTYPE.itemsize == 28

KEYLINE_DTYPE = np.dtype(
    [("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"), ("pt_y", "<f4"),
     ("response", "<f4"), ("size", "<f4"), ("startPointX", "<f4"), ("startPointY", "<f4"),
     ("endPointX", "<f4"), ("endPointY", "<f4"), ("sPointInOctaveX", "<f4"),
     ("sPointInOctaveY", "<f4"), ("ePointInOctaveX", "<f4"), ("ePointInOctaveY", "<f4"),
     ("lineLength", "<f4"), ("numOfPixels", "<i4")]
)
assert KEYLINE_DTYPE.itemsize == 68


def build(force: bool = False) -> Path:
    srcs = list(_DIR.glob("oracle_*.cpp")) + list(_DIR.glob("*.h"))
    stale = (not _LIB.exists()) or any(s.stat().st_mtime > _LIB.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(_DIR), "-B", "libplvi_oracle.so"], check=True,
                       capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB))
        _lib.plvio_fast_atan2.restype = C.c_float
        _lib.plvio_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.plvio_ic_angle.restype = C.c_float
        _lib.plvio_ic_angle.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        _lib.plvio_orb_descriptor.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float,
                                              C.c_void_p]
        _lib.plvio_orb_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 5
        _lib.plvio_orb_extract.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u8(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img


# ---- OpenCV-resident primitives ------------------------------------------------
def fast_atan2(y, x):
    return float(lib().plvio_fast_atan2(float(y), float(x)))


def resize_linear(img, dw, dh):
    img = _u8(img)
    out = np.empty((dh, dw), np.uint8)
    lib().plvio_resize_linear_u8(_p(img), img.strides[0], img.shape[1], img.shape[0], _p(out), int(dw), int(dw), int(dh))
    return out


def gaussian_blur7(img):
    img = _u8(img)
    out = np.empty_like(img)
    lib().plvio_gaussian_blur7_u8(_p(img), img.strides[0], img.shape[1], img.shape[0], _p(out), img.shape[1])
    return out


def gaussian_blur5(img):
    img = _u8(img)
    out = np.empty_like(img)
    lib().plvio_gaussian_blur5_u8(_p(img), img.strides[0], img.shape[1], img.shape[0], _p(out), img.shape[1])
    return out


def fast_score_map(img):
    img = _u8(img)
    out = np.empty(img.shape, np.int32)
    lib().plvio_fast_score_map(_p(img), img.strides[0], img.shape[1], img.shape[0], _p(out))
    return out


def fast_roi(img, th):
    """cv::FAST(img, th, nms=True) -> (n,3) float32 [x,y,response] in row-major order."""
    img = _u8(img)
    cap = img.size
    out = np.empty((cap, 3), np.float32)
    n = lib().plvio_fast_roi(_p(img), img.strides[0], img.shape[1], img.shape[0], int(th), _p(out), cap)
    return out[:n].copy()


def grid_fast(img, ini_th=20, min_th=7):
    """Candidates of ComputeKeyPointsOctTree for one level, relative to (16,16)."""
    img = _u8(img)
    cap = img.size
    out = np.empty((cap, 3), np.float32)
    n = lib().plvio_grid_fast(_p(img), img.strides[0], img.shape[1], img.shape[0], ini_th, min_th, _p(out), cap)
    return out[:n].copy()


def distribute_octree(cands, min_x, max_x, min_y, max_y, n_target):
    cands = np.ascontiguousarray(cands, np.float32)
    out = np.empty(len(cands) + 8, np.int32)
    n = lib().plvio_distribute_octree(_p(cands), len(cands), min_x, max_x, min_y, max_y, n_target,
                                      _p(out), len(out))
    return out[:n].copy()


def ic_angle(img, x, y):
    img = _u8(img)
    return float(lib().plvio_ic_angle(_p(img), img.strides[0], float(x), float(y)))


def orb_descriptor(img, x, y, angle):
    img = _u8(img)
    d = np.empty(32, np.uint8)
    lib().plvio_orb_descriptor(_p(img), img.strides[0], float(x), float(y), float(angle), _p(d))
    return d


# ---- ORBextractor ---------------------------------------------------------------
def orb_plan(w, h, nfeatures=1000, scale_factor=1.2, nlevels=8):
    lw = np.empty(nlevels, np.int32)
    lh = np.empty(nlevels, np.int32)
    sc = np.empty(nlevels, np.float32)
    q = np.empty(nlevels, np.int32)
    um = np.empty(16, np.int32)
    lib().plvio_orb_plan(w, h, nfeatures, scale_factor, nlevels, _p(lw), _p(lh), _p(sc), _p(q), _p(um))
    return {"w": lw, "h": lh, "scale": sc, "quota": q, "umax": um}


def orb_extract(img, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7,
                lapping=(0, 0), debug=False):
    """ORBextractor::operator().  Returns dict(keypoints, descriptors, mono_index[, pyramid, blurred, level_counts])."""
    img = _u8(img)
    h, w = img.shape
    plan = orb_plan(w, h, nfeatures, scale_factor, nlevels)
    cap = nfeatures + 3 * nlevels + 64
    kps = np.zeros(cap, KEYPOINT_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    mono = C.c_int(0)
    npx = int((plan["w"].astype(np.int64) * plan["h"]).sum())
    pyr = np.empty(npx, np.uint8) if debug else None
    blur = np.empty(npx, np.uint8) if debug else None
    lc = np.zeros(nlevels, np.int32)
    n = lib().plvio_orb_extract(_p(img), w, h, img.strides[0], nfeatures, scale_factor, nlevels,
                                ini_th, min_th, int(lapping[0]), int(lapping[1]), _p(kps), _p(desc),
                                cap, C.byref(mono), _p(pyr), _p(blur), _p(lc))
    if n < 0:
        raise RuntimeError(f"oracle orb_extract failed: {n}")
    out = {"keypoints": kps[:n].copy(), "descriptors": desc[:n].copy(), "mono_index": mono.value,
           "level_counts": lc, "plan": plan}
    if debug:
        out["pyramid"] = split_levels(pyr, plan)
        out["blurred"] = split_levels(blur, plan)
    return out


def split_levels(flat, plan):
    res, o = [], 0
    for w, h in zip(plan["w"], plan["h"]):
        res.append(flat[o:o + int(w) * int(h)].reshape(int(h), int(w)))
        o += int(w) * int(h)
    return res


# ---- matchers ----------------------------------------------------------------------
QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("min_level", "<i4"),
                        ("max_level", "<i4"), ("angle", "<f4"), ("flags", "<i4")])


def hamming256(a, b, shift25=False):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    f = lib().plvio_hamming256_shift25 if shift25 else lib().plvio_hamming256
    return int(f(_p(a), _p(b)))


def _grid_args(grid):
    g = np.asarray(grid).reshape(-1)[0]
    return [C.c_float(float(g["min_x"])), C.c_float(float(g["min_y"])), C.c_float(float(g["inv_w"])),
            C.c_float(float(g["inv_h"]))]


def _grid_floats(grid):
    g = np.asarray(grid).reshape(-1)[0]
    return [float(g["min_x"]), float(g["min_y"]), float(g["inv_w"]), float(g["inv_h"])]


def search_frame(keys, desc, grid, queries, qdesc, th=100, check_ori=True, blocked=None):
    keys = np.ascontiguousarray(keys)
    desc = np.ascontiguousarray(desc, np.uint8)
    queries = np.ascontiguousarray(queries)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    blocked = None if blocked is None else np.ascontiguousarray(blocked, np.uint8)
    mt = np.empty(max(len(keys), 1), np.int32)
    n = lib().plvio_search_frame(_p(keys), _p(desc), len(keys), _p(blocked), *_grid_args(grid), _p(queries),
                                 _p(qdesc), len(queries), int(th), int(check_ori), _p(mt))
    return n, mt[:len(keys)]


def search_in_radius(keys, desc, grid, queries, qdesc, inv_level_sigma2, chi2=5.99, th=50):
    """Per-map-point search of Fuse / SearchBySim3 / SearchByProjection(KF, Scw): (found, best_idx, best_dist)."""
    keys = np.ascontiguousarray(keys)
    desc = np.ascontiguousarray(desc, np.uint8)
    queries = np.ascontiguousarray(queries)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    s2 = np.zeros(16, np.float32)
    s2[:len(inv_level_sigma2)] = inv_level_sigma2
    bi = np.empty(max(len(queries), 1), np.int32)
    bd = np.empty(max(len(queries), 1), np.int32)
    f = lib().plvio_search_in_radius
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                  C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    n = f(_p(keys), _p(desc), len(keys), *_grid_args(grid), _p(queries), _p(qdesc), len(queries), _p(s2),
          float(chi2), int(th), _p(bi), _p(bd))
    return n, bi[:len(queries)], bd[:len(queries)]


def line_fuse_search(keylines, desc, queries, qdesc, flags=None, th_low=50):
    """Per-map-line search of LineMatcher::Fuse: queries [nq, 6] f32 = u1, v1, u2, v2, radius, predicted level."""
    kl = np.ascontiguousarray(keylines)
    desc = np.ascontiguousarray(desc, np.uint8)
    q = np.ascontiguousarray(queries, np.float32).reshape(-1, 6)
    qd = np.ascontiguousarray(qdesc, np.uint8)
    fl = None if flags is None else np.ascontiguousarray(flags, np.uint8)
    bi = np.empty(max(len(q), 1), np.int32)
    bd = np.empty(max(len(q), 1), np.int32)
    f = lib().plvio_line_fuse_search
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n = f(_p(kl), _p(desc), len(kl), _p(q), _p(fl), _p(qd), len(q), int(th_low), _p(bi), _p(bd))
    return n, bi[:len(q)], bd[:len(q)]


def search_mappoints(keys, desc, grid, queries, qdesc, th=100, nnratio=0.8, blocked=None):
    keys = np.ascontiguousarray(keys)
    desc = np.ascontiguousarray(desc, np.uint8)
    queries = np.ascontiguousarray(queries)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    blocked = None if blocked is None else np.ascontiguousarray(blocked, np.uint8)
    mt = np.empty(max(len(keys), 1), np.int32)
    n = lib().plvio_search_mappoints(_p(keys), _p(desc), len(keys), _p(blocked), *_grid_args(grid),
                                     _p(queries), _p(qdesc), len(queries), int(th), C.c_float(nnratio), _p(mt))
    return n, mt[:len(keys)]


def search_init(keys2, desc2, grid, queries, desc1, th=50, nnratio=0.9, check_ori=True):
    keys2 = np.ascontiguousarray(keys2)
    desc2 = np.ascontiguousarray(desc2, np.uint8)
    queries = np.ascontiguousarray(queries).copy()
    desc1 = np.ascontiguousarray(desc1, np.uint8)
    m12 = np.empty(max(len(queries), 1), np.int32)
    n = lib().plvio_search_init(_p(keys2), _p(desc2), len(keys2), *_grid_args(grid), _p(queries), _p(desc1),
                                len(queries), int(th), C.c_float(nnratio), int(check_ori), _p(m12))
    return n, m12[:len(queries)], queries


def search_bow(keys, desc, items, queries, qdesc, th=50, nnratio=0.7, check_ori=True):
    keys = np.ascontiguousarray(keys)
    desc = np.ascontiguousarray(desc, np.uint8)
    items = np.ascontiguousarray(items, np.int32)
    queries = np.ascontiguousarray(queries)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    mt = np.empty(max(len(keys), 1), np.int32)
    n = lib().plvio_search_bow(_p(keys), _p(desc), len(keys), _p(items), _p(queries), _p(qdesc), len(queries),
                               int(th), C.c_float(nnratio), int(check_ori), _p(mt))
    return n, mt[:len(keys)]


def match_nnr(d1, d2, nnr):
    d1 = np.ascontiguousarray(d1, np.uint8)
    d2 = np.ascontiguousarray(d2, np.uint8)
    m = np.empty(max(len(d1), 1), np.int32)
    n = lib().plvio_match_nnr(_p(d1), len(d1), _p(d2), len(d2), C.c_float(nnr), _p(m))
    return n, m[:len(d1)]


def line_match(d1, d2, nnr):
    d1 = np.ascontiguousarray(d1, np.uint8)
    d2 = np.ascontiguousarray(d2, np.uint8)
    m = np.empty(max(len(d1), 1), np.int32)
    n = lib().plvio_line_match(_p(d1), len(d1), _p(d2), len(d2), C.c_float(nnr), _p(m))
    return n, m[:len(d1)]


# ---- lines -------------------------------------------------------------------------
def gaussian_kernel_f64(n, sigma):
    k = np.empty(n, np.float64)
    lib().plvio_gaussian_kernel_f64(int(n), C.c_double(sigma), _p(k))
    return k


def gaussian_blur_f64(img, k):
    img = np.ascontiguousarray(img, np.float64)
    k = np.ascontiguousarray(k, np.float64)
    out = np.empty_like(img)
    lib().plvio_gaussian_blur_f64(_p(img), img.shape[1], img.shape[0], _p(out), _p(k), len(k))
    return out


def resize_linear_f64(img, dw, dh, fx, fy):
    img = np.ascontiguousarray(img, np.float64)
    out = np.empty((dh, dw), np.float64)
    lib().plvio_resize_linear_f64(_p(img), img.shape[1], img.shape[0], _p(out), int(dw), int(dh), C.c_double(fx), C.c_double(fy))
    return out


def pyr_down(img):
    img = _u8(img)
    h, w = img.shape
    out = np.empty((h // 2, w // 2), np.uint8)
    lib().plvio_pyr_down_u8(_p(img), w, h, _p(out), w // 2, h // 2)
    return out


def sobel3(img):
    img = _u8(img)
    dx = np.empty(img.shape, np.int16)
    dy = np.empty(img.shape, np.int16)
    lib().plvio_sobel3_s16(_p(img), img.shape[1], img.shape[0], _p(dx), _p(dy))
    return dx, dy


def lsd(img, lsd_scale=0.8, debug=False, refine=0):
    """LineSegmentDetectorImpl::detect (refine 0 / 1 / 2) on one u8 image -> (segments [n,4] f32[, dict])."""
    img = _u8(img)
    h, w = img.shape
    sw, sh = C.c_int(0), C.c_int(0)
    cap = 1 << 16
    segs = np.empty((cap, 4), np.float32)
    rs = np.empty(cap, np.int32)
    W, H = int(round(w * float(np.float32(lsd_scale)))), int(round(h * float(np.float32(lsd_scale))))
    big = (W + 2) * (H + 2)
    scaled = np.empty(big, np.float64) if debug else None
    ang = np.empty(big, np.float64) if debug else None
    mg = np.empty(big, np.float64) if debug else None
    n = lib().plvio_lsd(_p(img), img.strides[0], w, h, C.c_float(lsd_scale), C.byref(sw), C.byref(sh),
                        _p(scaled), _p(ang), _p(mg), _p(segs), cap, _p(rs), int(refine))
    out = segs[:n].copy()
    if not debug:
        return out
    m = sw.value * sh.value
    return out, {"w": sw.value, "h": sh.value, "scaled": scaled[:m].reshape(sh.value, sw.value),
                 "angles": ang[:m].reshape(sh.value, sw.value), "modgrad": mg[:m].reshape(sh.value, sw.value),
                 "region_sizes": rs[:n].copy()}


def lbd(keyline, dx, dy):
    kl = np.ascontiguousarray(np.asarray(keyline, KEYLINE_DTYPE).reshape(1))
    dx = np.ascontiguousarray(dx, np.int16)
    dy = np.ascontiguousarray(dy, np.int16)
    des = np.empty(72, np.float32)
    b = np.empty(32, np.uint8)
    lib().plvio_lbd(_p(kl), _p(dx), _p(dy), dx.shape[1], dx.shape[0], _p(des), _p(b))
    return des, b


def line_extract(img, lsd_nfeatures=200, lsd_refine=0, lsd_scale=0.8, nlevels=2, scale=2.0):
    """Lineextractor::operator() -> dict(keylines, descriptors, line_eq, raw_counts)."""
    img = _u8(img)
    h, w = img.shape
    cap = 1 << 15
    kl = np.zeros(cap, KEYLINE_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    eq = np.zeros((cap, 3), np.float64)
    raw = np.zeros(max(nlevels, 1), np.int32)
    n = lib().plvio_line_extract(_p(img), w, h, img.strides[0], int(lsd_nfeatures), int(lsd_refine),
                                 C.c_float(lsd_scale), int(nlevels), C.c_float(scale), _p(kl), _p(desc), _p(eq),
                                 cap, _p(raw))
    if n < 0:
        raise RuntimeError(f"oracle line_extract failed: {n}")
    return {"keylines": kl[:n].copy(), "descriptors": desc[:n].copy(), "line_eq": eq[:n].copy(), "raw_counts": raw}


# ---- Frame steps after extraction (oracle_frame.cpp)
EUROC_CAMERA = dict(fx=458.654, fy=457.296, cx=367.215, cy=248.375,
                    dist=(-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05))   # Examples/Monocular-Inertial/EuRoC.yaml


def stereo_matches(kps_l, desc_l, kps_r, desc_r, pyr_l, pyr_r, scale_factors, mb, mbf):
    """Frame::ComputeStereoMatches (src/Frame.cc:1228-1406).  pyr_l / pyr_r: lists of level images (orb_extract(...,
    debug=True)["pyramid"]).  Returns (mvuRight, mvDepth, number of stereo points kept)."""
    kl, kr = np.ascontiguousarray(kps_l), np.ascontiguousarray(kps_r)
    dl, dr = np.ascontiguousarray(desc_l, np.uint8), np.ascontiguousarray(desc_r, np.uint8)
    lw = np.array([p.shape[1] for p in pyr_l], np.int32)
    lh = np.array([p.shape[0] for p in pyr_l], np.int32)
    pl = np.concatenate([np.ascontiguousarray(p, np.uint8).reshape(-1) for p in pyr_l])
    pr = np.concatenate([np.ascontiguousarray(p, np.uint8).reshape(-1) for p in pyr_r])
    sf = np.ascontiguousarray(scale_factors, np.float32)
    isf = (np.float32(1.0) / sf).astype(np.float32)
    ur = np.empty(max(len(kl), 1), np.float32)
    dp = np.empty(max(len(kl), 1), np.float32)
    f = lib().plvio_stereo_matches
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    n = f(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), _p(pl), _p(pr), _p(lw), _p(lh), len(lw), _p(sf), _p(isf),
          float(mb), float(mbf), _p(ur), _p(dp))
    return ur[:len(kl)], dp[:len(kl)], n


def camera_arrays(cam):
    """(K[4], dist[14], P[4]) float64 arrays; K and dist go through float32 like the reference's cv::Mat(CV_32F)."""
    K = np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float32).astype(np.float64)
    k = np.zeros(14, np.float64)
    d = np.asarray(cam["dist"], np.float32).astype(np.float64)
    k[:len(d)] = d
    P = np.array([cam.get("new_fx", K[0]), cam.get("new_fy", K[1]), cam.get("new_cx", K[2]), cam.get("new_cy", K[3])], np.float64)
    return K, k, P


def undistort_points(xy, cam=EUROC_CAMERA, iters=5):
    """cv::undistortPoints(xy, K, dist, P=K) as Frame::UndistortKeyPoints calls it: [n,2] f32 -> [n,2] f32."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    K, k, P = camera_arrays(cam)
    out = np.empty_like(xy)
    lib().plvio_undistort_points(_p(xy), C.c_int(len(xy)), _p(K), _p(k), _p(P), C.c_int(iters), _p(out))
    return out


def assign_grid(xy, grid):
    """Frame::AssignFeaturesToGrid: (cell_start[3073], items[n_in_grid]) of the 64x48 grid, cell = ix*48+iy."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    start = np.zeros(64 * 48 + 1, np.int32)
    items = np.zeros(max(len(xy), 1), np.int32)
    g = np.asarray(grid).reshape(-1)[0]
    lib().plvio_assign_grid(_p(xy), C.c_int(len(xy)), C.c_float(g["min_x"]), C.c_float(g["min_y"]), C.c_float(g["inv_w"]),
                            C.c_float(g["inv_h"]), _p(start), _p(items))
    return start, items[:start[-1]]


# ---- DBoW2 transform (oracle_bow.cpp)
def bow_transform(vocab, desc, levelsup=4):
    """vocab: dict(k, L, scoring, weighting, parent[int32 n], desc[u8 n,32], weight[f64 n]); desc u8 [m,32].
    Returns dict(word_id, word_weight, node_id, bow=(words, values), fv=(nodes, start, features))."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    m = len(desc)
    parent = np.ascontiguousarray(vocab["parent"], np.int32)
    nd = np.ascontiguousarray(vocab["desc"], np.uint8)
    nw = np.ascontiguousarray(vocab["weight"], np.float64)
    wid, ww, nid = np.zeros(m, np.int32), np.zeros(m, np.float64), np.zeros(m, np.int32)
    bc, fc = C.c_int(0), C.c_int(0)
    bw, bv = np.zeros(max(m, 1), np.int32), np.zeros(max(m, 1), np.float64)
    fn, fs, ff = np.zeros(max(m, 1), np.int32), np.zeros(m + 1, np.int32), np.zeros(max(m, 1), np.int32)
    lib().plvio_bow_transform(C.c_int(vocab["L"]), C.c_int(vocab["scoring"]), C.c_int(vocab["weighting"]), C.c_int(len(parent)),
                              _p(parent), _p(nd), _p(nw), _p(desc), C.c_int(m), C.c_int(levelsup), _p(wid), _p(ww), _p(nid),
                              C.byref(bc), _p(bw), _p(bv), C.byref(fc), _p(fn), _p(fs), _p(ff))
    return {"word_id": wid, "word_weight": ww, "node_id": nid, "bow": (bw[:bc.value].copy(), bv[:bc.value].copy()),
            "fv": (fn[:fc.value].copy(), fs[:fc.value + 1].copy(), ff[:fs[fc.value]].copy())}


def line_match_mad(d1, d2, factor, has1=None, has2=None):
    """LineMatcher::SerachForInitialize (factor 0.5) / SearchForTriangulation (factor 0.1): (n, matches12, (nn_mad, nn12_mad))."""
    d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32)
    d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
    m = np.full(max(len(d1), 1), -1, np.int32)
    mad = np.zeros(2, np.float64)
    h1 = None if has1 is None else np.ascontiguousarray(has1, np.uint8)
    h2 = None if has2 is None else np.ascontiguousarray(has2, np.uint8)
    n = lib().plvio_line_match_mad(_p(d1), C.c_int(len(d1)), _p(d2), C.c_int(len(d2)), _p(h1), _p(h2), C.c_double(factor), _p(m), _p(mad))
    return n, m[:len(d1)], mad


def distinctive_descriptor(desc):
    """MapPoint::ComputeDistinctiveDescriptors: index of the descriptor with the least median distance."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    return int(lib().plvio_distinctive_descriptor(_p(desc), C.c_int(len(desc))))


def search_bow_kfkf(keys1, desc1, mp1, fv1, keys2, desc2, mp2, fv2, nnratio=0.8, check_ori=True):
    """ORBmatcher::SearchByBoW(KF1, KF2, vpMatches12): fvN = (nodes, start, features) FeatureVector CSR; mpN = has-map-point
    flags.  Returns (nmatches, matches12[n1]) with matches12[idx1] = idx2 or -1."""
    keys1, keys2 = np.ascontiguousarray(keys1, KEYPOINT_DTYPE), np.ascontiguousarray(keys2, KEYPOINT_DTYPE)
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    mp1, mp2 = np.ascontiguousarray(mp1, np.uint8), np.ascontiguousarray(mp2, np.uint8)
    a = [np.ascontiguousarray(x, np.int32) for x in fv1]
    b = [np.ascontiguousarray(x, np.int32) for x in fv2]
    m = np.full(max(len(keys1), 1), -1, np.int32)
    n = lib().plvio_search_bow_kfkf(_p(keys1), _p(desc1), _p(mp1), C.c_int(len(keys1)), _p(a[0]), _p(a[1]), _p(a[2]), C.c_int(len(a[0])),
                                    _p(keys2), _p(desc2), _p(mp2), C.c_int(len(keys2)), _p(b[0]), _p(b[1]), _p(b[2]), C.c_int(len(b[0])),
                                    C.c_float(nnratio), C.c_int(int(check_ori)), _p(m))
    return n, m[:len(keys1)]


def search_triangulation(keys1, desc1, mp1, fv1, keys2, desc2, mp2, fv2, F12, ep, sf2, sigma2_2, coarse=False, check_ori=True):
    """ORBmatcher::SearchForTriangulation (mono pinhole): (nmatches, matches12[n1]); vMatchedPairs = [(i, m[i]) for m[i] >= 0]."""
    keys1, keys2 = np.ascontiguousarray(keys1, KEYPOINT_DTYPE), np.ascontiguousarray(keys2, KEYPOINT_DTYPE)
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    mp1, mp2 = np.ascontiguousarray(mp1, np.uint8), np.ascontiguousarray(mp2, np.uint8)
    a = [np.ascontiguousarray(x, np.int32) for x in fv1]
    b = [np.ascontiguousarray(x, np.int32) for x in fv2]
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    sf = np.ascontiguousarray(sf2, np.float32)
    sg = np.ascontiguousarray(sigma2_2, np.float32)
    m = np.full(max(len(keys1), 1), -1, np.int32)
    n = lib().plvio_search_triangulation(_p(keys1), _p(desc1), _p(mp1), C.c_int(len(keys1)), _p(a[0]), _p(a[1]), _p(a[2]), C.c_int(len(a[0])),
                                         _p(keys2), _p(desc2), _p(mp2), C.c_int(len(keys2)), _p(b[0]), _p(b[1]), _p(b[2]), C.c_int(len(b[0])),
                                         _p(F), C.c_float(ep[0]), C.c_float(ep[1]), _p(sf), _p(sg), C.c_int(int(coarse)),
                                         C.c_int(int(check_ori)), _p(m))
    return n, m[:len(keys1)]


# ---- oracle/_ref: the reference's own sources compiled against the cvmini stand-in ----------------
# (oracle/Makefile.ref; built where /root/reference exists, shipped prebuilt to the GPU box)
_REF_LIB = _DIR / "_ref" / "libplvi_ref.so"
_REF_ORB_LIB = _DIR / "_ref" / "libplvi_ref_orbmatcher.so"   # ORBmatcher.cc (stand-in classes of its own, see Makefile.ref)
_REF_MP_LIB = _DIR / "_ref" / "libplvi_ref_mappoint.so"     # MapPoint.cc + MapPoint.h over stand-in KeyFrame / Frame / Map
_REF_FR_LIB = _DIR / "_ref" / "libplvi_ref_frame.so"        # Frame.cc + Frame.h over stand-in MapPoint / KeyFrame / cameras / IMU types
_REF_PH_LIB = _DIR / "_ref" / "libplvi_ref_pinhole.so"      # Pinhole.cpp + Pinhole.h + GeometricCamera.h
_ref_orb = None
_ref_mp = None
_ref_fr = None
_ref_ph = None
REFERENCE_ROOT = Path(os.environ.get("PLVI_REFERENCE_ROOT", "/root/reference"))
_ref = None

# DROP-IN PROOF (Makefile.ref): the same glue and the reference's own Frame.cc / KeyFrame.cc compiled against the
# PRODUCT's drop-in headers (pl_vi_orbslam3_b200/shim) and linked with libplvi_cuda.so.  Inside `with dropin():` every
# ref_* wrapper of this module runs on those builds, so a test calls the same function twice and compares.
_DROPIN_LIBS = {"libplvi_ref.so": "libplvi_dropin.so", "libplvi_ref_orbmatcher.so": "libplvi_dropin_orbmatcher.so",
                "libplvi_ref_frame.so": "libplvi_dropin_frame.so"}
_dropin_on = False
_dropin_cache = {}


def dropin_available() -> bool:
    return all((_DIR / "_ref" / n).exists() for n in _DROPIN_LIBS.values())


class dropin:
    """Context manager: route ref_lib() / ref_orbmatcher_lib() / ref_frame_lib() to the drop-in builds."""
    def __enter__(self):
        global _dropin_on
        if not dropin_available():
            raise RuntimeError("oracle/_ref/libplvi_dropin*.so are not built")
        self._prev, _dropin_on = _dropin_on, True
        return self

    def __exit__(self, *exc):
        global _dropin_on
        _dropin_on = self._prev
        return False


def _dropin_lib(ref_path):
    name = _DROPIN_LIBS[Path(ref_path).name]
    if name not in _dropin_cache:
        lib()
        # the product library first, by absolute path (the drop-in builds name it as a dependency)
        C.CDLL(str(_DIR.parent / "pl_vi_orbslam3_b200" / "libplvi_cuda.so"), mode=C.RTLD_GLOBAL)
        l = C.CDLL(str(_DIR / "_ref" / name))
        if name == "libplvi_dropin.so":
            _ref_argtypes(l)
        _dropin_cache[name] = l
    return _dropin_cache[name]


def _ref_argtypes(l):
    l.plviref_orb_extract.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p]
    if hasattr(l, "plviref_lsd"):
        l.plviref_lsd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int]
    l.plviref_line_extract.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                       C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]


def ref_build(force: bool = False):
    """Build oracle/_ref/libplvi_ref.so if the reference tree is present.  Returns the path or None."""
    if not (REFERENCE_ROOT / "src" / "ORBextractor.cc").exists():
        return _REF_LIB if _REF_LIB.exists() else None
    build()
    srcs = [_DIR / "ref_glue.cpp", _DIR / "ref_glue_linematcher.cpp", _DIR / "ref_glue_orbmatcher.cpp", _DIR / "ref_glue_mappoint.cpp", _DIR / "ref_glue_frame.cpp", _DIR / "ref_glue_pinhole.cpp",
            _DIR / "cvmini" / "slam_mock_pinhole.h",
            _DIR / "cvmini" / "slam_mock_frame.h", _DIR / "cvmini" / "slam_mock_keyframe.h",
            _DIR / "Makefile.ref",
            _DIR / "cvmini" / "cvmini.hpp", _DIR / "cvmini" / "eigenmini.hpp", _DIR / "cvmini" / "slam_mock.h",
            _DIR / "cvmini" / "slam_mock_orb.h", _LIB]
    shim = _DIR.parent / "pl_vi_orbslam3_b200" / "shim"
    srcs += sorted(shim.glob("include/*.h")) + sorted(shim.glob("src/*")) + [_DIR.parent / "include" / "plvi.h"]
    targets = [_REF_LIB, _REF_ORB_LIB, _REF_MP_LIB, _REF_FR_LIB, _REF_PH_LIB] + [_DIR / "_ref" / n for n in _DROPIN_LIBS.values()]
    stale = any((not t.exists()) or any(s.stat().st_mtime > t.stat().st_mtime for s in srcs) for t in targets)
    if force or stale:
        subprocess.run(["make", "-C", str(_DIR), "-f", "Makefile.ref", f"REF={REFERENCE_ROOT}"] + (["-B"] if force else []),
                       check=True, capture_output=True)
    return _REF_LIB


def ref_available() -> bool:
    return ref_build() is not None


def ref_lib():
    global _ref
    if _dropin_on:
        return _dropin_lib(_REF_LIB)
    if _ref is None:
        p = ref_build()
        if p is None:
            raise RuntimeError("oracle/_ref/libplvi_ref.so is not built and /root/reference is absent")
        lib()   # libplvi_oracle.so first (the stand-in's primitives resolve into it)
        _ref = C.CDLL(str(p))
        _ref_argtypes(_ref)
    return _ref


def ref_set_monotone(on: bool):
    """Heap addresses monotone inside the compiled reference (default, makes DistributeOctTree's pointer-ordered
    ties deterministic = the oracle's definition) or plain malloc (timing runs)."""
    ref_lib().plviref_set_monotone(int(bool(on)))


def ref_orb_extract(img, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, lapping=(0, 0),
                    debug=False):
    """The reference's ORBextractor::operator() itself (src/ORBextractor.cc), same outputs as orb_extract."""
    img = _u8(img)
    h, w = img.shape
    plan = orb_plan(w, h, nfeatures, scale_factor, nlevels)
    cap = 4 * nfeatures + 64
    kps = np.zeros(cap, KEYPOINT_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    mono = C.c_int(0)
    npx = int((plan["w"].astype(np.int64) * plan["h"]).sum())
    pyr = np.empty(npx, np.uint8) if debug else None
    n = ref_lib().plviref_orb_extract(_p(img), w, h, img.strides[0], nfeatures, scale_factor, nlevels, ini_th,
                                      min_th, int(lapping[0]), int(lapping[1]), _p(kps), _p(desc), cap,
                                      C.byref(mono), _p(pyr))
    if n < 0:
        raise RuntimeError(f"reference orb_extract failed: {n}")
    out = {"keypoints": kps[:n].copy(), "descriptors": desc[:n].copy(), "mono_index": mono.value}
    if debug:
        out["pyramid"] = split_levels(pyr, plan)
    return out


def ref_bow_transform(voc_text_path, desc, levelsup=4):
    """The reference's own DBoW2 (TemplatedVocabulary::loadFromTextFile + transform) on a vocabulary text file.
    Returns dict(words=vocabulary size, bow=(words, values), fv=(nodes, start, features))."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    m = len(desc)
    bc, fc = C.c_int(0), C.c_int(0)
    bw, bv = np.zeros(max(m, 1), np.int32), np.zeros(max(m, 1), np.float64)
    fn, fs, ff = np.zeros(max(m, 1), np.int32), np.zeros(m + 1, np.int32), np.zeros(max(m, 1), np.int32)
    f = ref_lib().plviref_bow_transform
    f.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7
    n = f(str(voc_text_path).encode(), _p(desc), m, int(levelsup), C.byref(bc), _p(bw), _p(bv), C.byref(fc), _p(fn), _p(fs), _p(ff))
    if n < 0:
        raise RuntimeError("reference DBoW2 could not load the vocabulary text file")
    return {"words": n, "bow": (bw[:bc.value].copy(), bv[:bc.value].copy()),
            "fv": (fn[:fc.value].copy(), fs[:fc.value + 1].copy(), ff[:fs[fc.value]].copy())}


def ref_lsd(img, lsd_scale=0.8, refine=0):
    """The reference's LineSegmentDetectorImpl::detect itself (src/LSD/lsd.cpp) on one u8 image."""
    img = _u8(img)
    h, w = img.shape
    cap = 1 << 16
    segs = np.empty((cap, 4), np.float32)
    n = ref_lib().plviref_lsd(_p(img), img.strides[0], w, h, int(refine), C.c_float(lsd_scale), _p(segs), cap)
    return segs[:n].copy()


def ref_line_extract(img, lsd_nfeatures=200, lsd_refine=0, lsd_scale=0.8, nlevels=2, scale=2.0):
    """The reference's Lineextractor::operator() itself (LSD + LBD), same outputs as line_extract."""
    img = _u8(img)
    h, w = img.shape
    cap = 1 << 15
    kl = np.zeros(cap, KEYLINE_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    eq = np.zeros((cap, 3), np.float64)
    n = ref_lib().plviref_line_extract(_p(img), w, h, img.strides[0], int(lsd_nfeatures), int(lsd_refine),
                                       C.c_float(lsd_scale), int(nlevels), C.c_float(scale), _p(kl), _p(desc),
                                       _p(eq), cap)
    if n < 0:
        raise RuntimeError(f"reference line_extract failed: {n}")
    return {"keylines": kl[:n].copy(), "descriptors": desc[:n].copy(), "line_eq": eq[:n].copy()}


def _ref_descs(d1, d2):
    d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32)
    d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
    return d1, d2


def ref_line_match(d1, d2, nnr, variant="match"):
    """The reference's LineMatcher::matchNNR ("nnr"), match(desc1, desc2) ("match") or match(vpLocalMapLines, Frame)
    ("maplines") itself (src/LineMatcher.cpp:40-111).  Needs >= 2 rows on both sides (the reference reads
    matches_[idx][1] unconditionally).  Returns (nmatches, matches12)."""
    d1, d2 = _ref_descs(d1, d2)
    if len(d1) < 2 or len(d2) < 2:
        raise ValueError("the reference's matchNNR is undefined for fewer than 2 descriptors on a side")
    f = getattr(ref_lib(), {"nnr": "plviref_line_match_nnr", "match": "plviref_line_match",
                            "maplines": "plviref_line_match_maplines"}[variant])
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
    m = np.full(len(d1), -1, np.int32)
    n = f(_p(d1), len(d1), _p(d2), len(d2), C.c_float(nnr), _p(m))
    return n, m


def ref_line_match_mad(d1, d2, factor, has1=None, has2=None):
    """The reference's LineMatcher::SerachForInitialize (factor 0.5, no has-flags) or SearchForTriangulation(KF, KF)
    (factor 0.1) itself (src/LineMatcher.cpp:113-171); lineDescriptorMAD is the oracle's restatement of
    src/Frame.cc:1089-1113 (Frame.cc itself needs the whole SLAM object graph).  Returns (nmatches, matches12) in the
    layout of line_match_mad."""
    d1, d2 = _ref_descs(d1, d2)
    if len(d1) < 1 or len(d2) < 2:
        raise ValueError("the reference's knn-2 needs >= 2 train descriptors")
    pairs = np.zeros((len(d1), 2), np.int32)
    if factor == 0.5:
        assert has1 is None and has2 is None
        f = ref_lib().plviref_line_search_for_initialize
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        n = f(_p(d1), len(d1), _p(d2), len(d2), _p(pairs))
    elif factor == 0.1:
        h1 = None if has1 is None else np.ascontiguousarray(has1, np.uint8)
        h2 = None if has2 is None else np.ascontiguousarray(has2, np.uint8)
        f = ref_lib().plviref_line_search_for_triangulation
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        n = f(_p(d1), len(d1), _p(d2), len(d2), _p(h1), _p(h2), _p(pairs))
    else:
        raise ValueError("the reference hard-codes the factors 0.5 (initialize) and 0.1 (triangulation)")
    m = np.full(len(d1), -1, np.int32)
    m[pairs[:n, 0]] = pairs[:n, 1]
    assert np.all(np.diff(pairs[:n, 0]) > 0)   # query order
    return n, m


def ref_line_distance(a, b, which=0):
    """LineMatcher::distance (which=0) / DescriptorDistance (which=1) of the reference on two 32-byte descriptors."""
    a, b = np.ascontiguousarray(a, np.uint8).reshape(32), np.ascontiguousarray(b, np.uint8).reshape(32)
    f = ref_lib().plviref_line_distance
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    return int(f(_p(a), _p(b), int(which)))


def ref_orbmatcher_lib():
    """oracle/_ref/libplvi_ref_orbmatcher.so: the reference's src/ORBmatcher.cc compiled unmodified against the stand-in
    Frame / KeyFrame / MapPoint of cvmini/slam_mock_orb.h."""
    global _ref_orb
    if _dropin_on:
        return _dropin_lib(_REF_ORB_LIB)
    if _ref_orb is None:
        if ref_build() is None or not _REF_ORB_LIB.exists():
            raise RuntimeError("oracle/_ref/libplvi_ref_orbmatcher.so is not built and /root/reference is absent")
        lib()
        _ref_orb = C.CDLL(str(_REF_ORB_LIB))
    return _ref_orb


def _fv_args(fv):
    a = [np.ascontiguousarray(x, np.int32) for x in fv]
    return a, [_p(a[0]), _p(a[1]), _p(a[2]), C.c_int(len(a[0]))]


def ref_search_mappoints(keys, desc, grid, scale_factors, proj, viewcos, level, flags, qdesc, th=1.0, nnratio=0.8, blocked=None,
                         real_frame_bounds=None):
    """The reference's ORBmatcher::SearchByProjection(F, vpMapPoints, th) itself (src/ORBmatcher.cc:44-214, monocular
    frame).  flags: bit0 not in view, bit1 Observations() == 0, bit2 isBad().  Returns (nmatches, match_train).
    real_frame_bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY): run it on the reference's own Frame class (libplvi_ref_frame.so:
    its AssignFeaturesToGrid / GetFeaturesInArea underneath) instead of the stand-in Frame."""
    keys = np.ascontiguousarray(keys, KEYPOINT_DTYPE)
    desc = np.ascontiguousarray(desc, np.uint8)
    g = np.array(_grid_floats(grid), np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    proj = np.ascontiguousarray(proj, np.float32).reshape(-1, 2)
    vc, lv, fl = np.ascontiguousarray(viewcos, np.float32), np.ascontiguousarray(level, np.int32), np.ascontiguousarray(flags, np.int32)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    blk = None if blocked is None else np.ascontiguousarray(blocked, np.uint8)
    mt = np.full(max(len(keys), 1), -1, np.int32)
    if real_frame_bounds is not None:
        g = np.array(real_frame_bounds, np.float32)
        f = ref_frame_lib().plviref_orb_search_by_projection_mappoints_realframe
    else:
        f = ref_orbmatcher_lib().plviref_orb_search_by_projection_mappoints
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
    n = f(_p(keys), _p(desc), len(keys), _p(blk), _p(g), _p(sf), len(sf), _p(proj), _p(vc), _p(lv), _p(fl), _p(qdesc), len(proj),
          C.c_float(th), C.c_float(nnratio), _p(mt))
    return n, mt[:len(keys)]


def ref_search_init(keys1, desc1, keys2, desc2, grid2, prev_matched, window=100, nnratio=0.9, check_ori=True, real_frame_bounds=None):
    """The reference's ORBmatcher::SearchForInitialization itself (src/ORBmatcher.cc:706-821).
    Returns (nmatches, matches12, prev_matched updated)."""
    keys1, keys2 = np.ascontiguousarray(keys1, KEYPOINT_DTYPE), np.ascontiguousarray(keys2, KEYPOINT_DTYPE)
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    g = np.array(_grid_floats(grid2), np.float32)
    pm = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    m12 = np.full(max(len(keys1), 1), -1, np.int32)
    if real_frame_bounds is not None:   # the reference's own Frame class underneath (see ref_search_mappoints)
        g = np.array(real_frame_bounds, np.float32)
        f = ref_frame_lib().plviref_orb_search_for_initialization_realframe
    else:
        f = ref_orbmatcher_lib().plviref_orb_search_for_initialization
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                  C.c_int, C.c_void_p]
    n = f(_p(keys1), _p(desc1), len(keys1), _p(keys2), _p(desc2), len(keys2), _p(g), _p(pm), int(window), C.c_float(nnratio),
          int(check_ori), _p(m12))
    return n, m12[:len(keys1)], pm


def ref_search_bow_kf_f(keys1, desc1, mp1, fv1, keys2, desc2, fv2, nnratio=0.7, check_ori=True):
    """The reference's ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) itself (src/ORBmatcher.cc:269-471, monocular).
    mp1: 0 none / 1 good / 2 bad map point.  Returns (nmatches, match_train[n2] = keyframe feature or -1)."""
    keys1, keys2 = np.ascontiguousarray(keys1, KEYPOINT_DTYPE), np.ascontiguousarray(keys2, KEYPOINT_DTYPE)
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    mp1 = np.ascontiguousarray(mp1, np.uint8)
    a, fa = _fv_args(fv1)
    b, fb = _fv_args(fv2)
    mt = np.full(max(len(keys2), 1), -1, np.int32)
    f = ref_orbmatcher_lib().plviref_orb_search_by_bow_kf_f
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    n = f(_p(keys1), _p(desc1), _p(mp1), len(keys1), *fa, _p(keys2), _p(desc2), len(keys2), *fb, C.c_float(nnratio), int(check_ori),
          _p(mt))
    return n, mt[:len(keys2)]


def ref_search_bow_kfkf(keys1, desc1, mp1, fv1, keys2, desc2, mp2, fv2, nnratio=0.8, check_ori=True):
    """The reference's ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) itself (src/ORBmatcher.cc:823-963); arguments and
    result as search_bow_kfkf."""
    keys1, keys2 = np.ascontiguousarray(keys1, KEYPOINT_DTYPE), np.ascontiguousarray(keys2, KEYPOINT_DTYPE)
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    mp1, mp2 = np.ascontiguousarray(mp1, np.uint8), np.ascontiguousarray(mp2, np.uint8)
    a, fa = _fv_args(fv1)
    b, fb = _fv_args(fv2)
    m = np.full(max(len(keys1), 1), -1, np.int32)
    f = ref_orbmatcher_lib().plviref_orb_search_by_bow_kf_kf
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int,
                  C.c_void_p]
    n = f(_p(keys1), _p(desc1), _p(mp1), len(keys1), *fa, _p(keys2), _p(desc2), _p(mp2), len(keys2), *fb, C.c_float(nnratio),
          int(check_ori), _p(m))
    return n, m[:len(keys1)]


def ref_orb_descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint8).reshape(32), np.ascontiguousarray(b, np.uint8).reshape(32)
    f = ref_orbmatcher_lib().plviref_orb_descriptor_distance
    f.argtypes = [C.c_void_p, C.c_void_p]
    return int(f(_p(a), _p(b)))


def ref_search_frame(keys2, desc2, grid, bounds, scale_factors, keys1, uv, flags, qdesc, th, check_ori=True, blocked=None,
                     real_frame=False):
    """The reference's ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono=true) itself (src/ORBmatcher.cc:
    1962-2178) with identity poses and map point i at (uv[i], 1) before a unit pinhole camera (its projection code then
    yields uv[i] exactly).  bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY); flags bit0: no map point / outlier, bit1:
    Observations() == 0.  Returns (nmatches, match_train)."""
    keys1, keys2 = np.ascontiguousarray(keys1, KEYPOINT_DTYPE), np.ascontiguousarray(keys2, KEYPOINT_DTYPE)
    desc2, qdesc = np.ascontiguousarray(desc2, np.uint8), np.ascontiguousarray(qdesc, np.uint8)
    g, b = np.array(_grid_floats(grid), np.float32), np.array(bounds, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
    fl = np.ascontiguousarray(flags, np.int32)
    blk = None if blocked is None else np.ascontiguousarray(blocked, np.uint8)
    mt = np.full(max(len(keys2), 1), -1, np.int32)
    if real_frame:   # the reference's own Frame class underneath (see ref_search_mappoints); the grid follows from the bounds
        f = ref_frame_lib().plviref_orb_search_by_projection_frame_realframe
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        n = f(_p(keys2), _p(desc2), len(keys2), _p(blk), _p(b), _p(sf), len(sf), _p(keys1), len(keys1), _p(uv), _p(fl), _p(qdesc),
              C.c_float(th), int(check_ori), _p(mt))
        return n, mt[:len(keys2)]
    f = ref_orbmatcher_lib().plviref_orb_search_by_projection_frame
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
    n = f(_p(keys2), _p(desc2), len(keys2), _p(blk), _p(g), _p(b), _p(sf), len(sf), _p(keys1), len(keys1), _p(uv), _p(fl), _p(qdesc),
          C.c_float(th), int(check_ori), _p(mt))
    return n, mt[:len(keys2)]


def ref_search_triangulation(keys1, desc1, mp1, fv1, keys2, desc2, mp2, fv2, F12, ep, sf2, sigma2_1, sigma2_2, coarse=False,
                             check_ori=True):
    """The reference's ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, false, bCoarse) itself
    (src/ORBmatcher.cc:965-1206), monocular; epipole through its own pose arithmetic (identity rotations), F12 handed to
    the stand-in camera's epipolar-line test.  Arguments / result as search_triangulation."""
    keys1, keys2 = np.ascontiguousarray(keys1, KEYPOINT_DTYPE), np.ascontiguousarray(keys2, KEYPOINT_DTYPE)
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    mp1, mp2 = np.ascontiguousarray(mp1, np.uint8), np.ascontiguousarray(mp2, np.uint8)
    a, fa = _fv_args(fv1)
    b, fb = _fv_args(fv2)
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    sf, s1, s2 = (np.ascontiguousarray(x, np.float32) for x in (sf2, sigma2_1, sigma2_2))
    m = np.full(max(len(keys1), 1), -1, np.int32)
    f = ref_orbmatcher_lib().plviref_orb_search_for_triangulation
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    n = f(_p(keys1), _p(desc1), _p(mp1), len(keys1), *fa, _p(keys2), _p(desc2), _p(mp2), len(keys2), *fb, _p(F), C.c_float(ep[0]),
          C.c_float(ep[1]), _p(sf), _p(s1), _p(s2), len(sf), int(coarse), int(check_ori), _p(m))
    return n, m[:len(keys1)]


def _ref_projection_args(keys, desc, grid, bounds, scale_factors, uv, level, flags, qdesc):
    keys = np.ascontiguousarray(keys, KEYPOINT_DTYPE)
    desc, qdesc = np.ascontiguousarray(desc, np.uint8), np.ascontiguousarray(qdesc, np.uint8)
    g, b = np.array(_grid_floats(grid), np.float32), np.array(bounds, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
    lv, fl = np.ascontiguousarray(level, np.int32), np.ascontiguousarray(flags, np.int32)
    return keys, desc, g, b, sf, uv, lv, fl, qdesc


def ref_fuse(keys, desc, grid, bounds, scale_factors, inv_level_sigma2, uv, level, flags, qdesc, th=3.0, sim3=False):
    """The reference's ORBmatcher::Fuse(pKF, vpMapPoints, th) (src/ORBmatcher.cc:1399-1610) or, sim3=True,
    Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (:1612-1734) itself: identity pose, map point i at (uv[i], 1), predicted
    level[i]; flags bit0 = isBad().  Returns (nFused, best_idx[nq])."""
    keys, desc, g, b, sf, uv, lv, fl, qdesc = _ref_projection_args(keys, desc, grid, bounds, scale_factors, uv, level, flags, qdesc)
    inv = np.ascontiguousarray(inv_level_sigma2, np.float32)
    bi = np.full(max(len(uv), 1), -1, np.int32)
    f = ref_orbmatcher_lib().plviref_orb_fuse_sim3 if sim3 else ref_orbmatcher_lib().plviref_orb_fuse
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
    n = f(_p(keys), _p(desc), len(keys), _p(g), _p(b), _p(sf), _p(inv), len(sf), _p(uv), _p(lv), _p(fl), _p(qdesc), len(uv),
          C.c_float(th), _p(bi))
    return n, bi[:len(uv)]


def ref_search_by_projection_kf(keys, desc, grid, bounds, scale_factors, uv, level, flags, qdesc, th, ratio_hamming=1.0,
                                matched_in=None):
    """The reference's ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) itself
    (src/ORBmatcher.cc:473-586), Scw = identity.  Returns (nmatches, match_train[n])."""
    keys, desc, g, b, sf, uv, lv, fl, qdesc = _ref_projection_args(keys, desc, grid, bounds, scale_factors, uv, level, flags, qdesc)
    mi = None if matched_in is None else np.ascontiguousarray(matched_in, np.uint8)
    mt = np.full(max(len(keys), 1), -1, np.int32)
    f = ref_orbmatcher_lib().plviref_orb_search_by_projection_kf
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
    n = f(_p(keys), _p(desc), len(keys), _p(mi), _p(g), _p(b), _p(sf), len(sf), _p(uv), _p(lv), _p(fl), _p(qdesc), len(uv), int(th),
          C.c_float(ratio_hamming), _p(mt))
    return n, mt[:len(keys)]


def ref_line_fuse(keylines, desc, bounds, scale_factors, queries, qdesc, flags=None, th=3.0):
    """The reference's LineMatcher::Fuse(pKF, vpMapLines, th) itself (src/LineMatcher.cpp:373-485): identity pose, unit
    pinhole, map line i with endpoints (u1, v1, 1) / (u2, v2, 1) and predicted level queries[i, 5] (queries[i, 4] is not
    read: the reference computes radius = th * mvScaleFactors[level]); flags[i] != 0 = isBad().
    Returns (nFused, best_idx[nq])."""
    kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE)
    desc, qd = np.ascontiguousarray(desc, np.uint8), np.ascontiguousarray(qdesc, np.uint8)
    q = np.ascontiguousarray(queries, np.float32).reshape(-1, 6)
    b, sf = np.array(bounds, np.float32), np.ascontiguousarray(scale_factors, np.float32)
    fl = None if flags is None else np.ascontiguousarray(flags, np.uint8)
    bi = np.full(max(len(q), 1), -1, np.int32)
    f = ref_lib().plviref_line_fuse
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_float, C.c_void_p]
    n = f(_p(kl), _p(desc), len(kl), _p(b), _p(sf), len(sf), _p(q), _p(fl), _p(qd), len(q), C.c_float(th), _p(bi))
    return n, bi[:len(q)]


def ref_distinctive_descriptor(desc, bad=None):
    """The reference's MapPoint::ComputeDistinctiveDescriptors itself (src/MapPoint.cc:330-402; MapPoint.cc and MapPoint.h
    compiled unmodified over stand-in KeyFrame / Frame / Map): observation i = feature 0 of keyframe i with descriptor
    desc[i]; bad[i] = pKF->isBad().  Returns the chosen 32-byte descriptor, or None when the reference returns early."""
    global _ref_mp
    if _ref_mp is None:
        if ref_build() is None or not _REF_MP_LIB.exists():
            raise RuntimeError("oracle/_ref/libplvi_ref_mappoint.so is not built and /root/reference is absent")
        lib()
        _ref_mp = C.CDLL(str(_REF_MP_LIB))
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    b = None if bad is None else np.ascontiguousarray(bad, np.uint8)
    out = np.zeros(32, np.uint8)
    f = _ref_mp.plviref_distinctive_descriptor
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return out if f(_p(desc), _p(b), len(desc), _p(out)) else None


def ref_mapline_distinctive_descriptor(desc, bad=None):
    """The reference's MapLine::ComputeDistinctiveDescriptors itself (src/MapLine.cc:264-329; MapLine.cc and MapLine.h
    compiled unmodified over stand-in KeyFrame / Frame / Map): observation i = line 0 of keyframe i with LBD descriptor
    desc[i]; bad[i] = pKF->isBad().  Returns the chosen 32-byte descriptor, or None when the reference returns early."""
    ref_distinctive_descriptor(np.zeros((1, 32), np.uint8))   # loads the library
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    b = None if bad is None else np.ascontiguousarray(bad, np.uint8)
    out = np.zeros(32, np.uint8)
    f = _ref_mp.plviref_mapline_distinctive_descriptor
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return out if f(_p(desc), _p(b), len(desc), _p(out)) else None


def ref_search_by_sim3(keys1, desc1, uv1, level1, flags1, keys2, desc2, uv2, level2, flags2, grid, bounds, scale_factors, th):
    """The reference's ORBmatcher::SearchBySim3 itself (src/ORBmatcher.cc:1736-1960) with s12 = 1, R12 = I, t12 = 0 and
    identity keyframe poses: the map point of feature i of keyframe A sits at (uvA[i], 1) and projects to uvA[i] in the
    other keyframe, predicted level levelA[i].  flags bit0: no map point, bit1: isBad(), bit2 (keyframe 1): already in
    vpMatches12.  Returns (nFound, matches12[n1])."""
    keys1, keys2 = np.ascontiguousarray(keys1, KEYPOINT_DTYPE), np.ascontiguousarray(keys2, KEYPOINT_DTYPE)
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    uv1, uv2 = (np.ascontiguousarray(x, np.float32).reshape(-1, 2) for x in (uv1, uv2))
    l1, l2, f1, f2 = (np.ascontiguousarray(x, np.int32) for x in (level1, level2, flags1, flags2))
    g, b = np.array(_grid_floats(grid), np.float32), np.array(bounds, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    m = np.full(max(len(keys1), 1), -1, np.int32)
    f = ref_orbmatcher_lib().plviref_orb_search_by_sim3
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
    n = f(_p(keys1), _p(desc1), len(keys1), _p(uv1), _p(l1), _p(f1), _p(keys2), _p(desc2), len(keys2), _p(uv2), _p(l2), _p(f2), _p(g),
          _p(b), _p(sf), len(sf), C.c_float(th), _p(m))
    return n, m[:len(keys1)]


def ref_search_reloc(keys2, desc2, grid, bounds, scale_factors, keys1, uv, level, flags, qdesc, th, orb_dist, check_ori=True, blocked=None):
    """The reference's ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) itself
    (src/ORBmatcher.cc:2180-2302): identity pose, keyframe map point i at (uv[i], 1), predicted level[i]; flags bit0: no map
    point, bit1: isBad(), bit2: in sAlreadyFound; blocked: frame features that already hold a map point.
    Returns (nmatches, match_train[n2])."""
    keys1, keys2 = np.ascontiguousarray(keys1, KEYPOINT_DTYPE), np.ascontiguousarray(keys2, KEYPOINT_DTYPE)
    desc2, qdesc = np.ascontiguousarray(desc2, np.uint8), np.ascontiguousarray(qdesc, np.uint8)
    g, b = np.array(_grid_floats(grid), np.float32), np.array(bounds, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
    lv, fl = np.ascontiguousarray(level, np.int32), np.ascontiguousarray(flags, np.int32)
    blk = None if blocked is None else np.ascontiguousarray(blocked, np.uint8)
    mt = np.full(max(len(keys2), 1), -1, np.int32)
    f = ref_orbmatcher_lib().plviref_orb_search_by_projection_reloc
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]
    n = f(_p(keys2), _p(desc2), len(keys2), _p(blk), _p(g), _p(b), _p(sf), len(sf), _p(keys1), len(keys1), _p(uv), _p(lv), _p(fl),
          _p(qdesc), C.c_float(th), int(orb_dist), int(check_ori), _p(mt))
    return n, mt[:len(keys2)]


def ref_frame_lib():
    """oracle/_ref/libplvi_ref_frame.so: the reference's src/Frame.cc + include/Frame.h compiled unmodified over the stand-ins
    of cvmini/slam_mock_frame.h."""
    global _ref_fr
    if _dropin_on:
        return _dropin_lib(_REF_FR_LIB)
    if _ref_fr is None:
        if ref_build() is None or not _REF_FR_LIB.exists():
            raise RuntimeError("oracle/_ref/libplvi_ref_frame.so is not built and /root/reference is absent")
        lib()
        _ref_fr = C.CDLL(str(_REF_FR_LIB))
    return _ref_fr


def ref_assign_grid(keys, bounds):
    """The reference's Frame::AssignFeaturesToGrid / PosInGrid itself: (cell_start[3073], items) like assign_grid."""
    keys = np.ascontiguousarray(keys, KEYPOINT_DTYPE)
    b = np.array(bounds, np.float32)
    start = np.zeros(64 * 48 + 1, np.int32)
    items = np.zeros(max(len(keys), 1), np.int32)
    f = ref_frame_lib().plviref_frame_assign_grid
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    f(_p(keys), len(keys), _p(b), _p(start), _p(items))
    return start, items[:start[-1]]


def ref_features_in_area(keys, bounds, xyr, levels):
    """The reference's Frame::GetFeaturesInArea itself for queries xyr [nq,3] (x, y, r), levels [nq,2] (min, max):
    list of index arrays in the reference's output order."""
    keys = np.ascontiguousarray(keys, KEYPOINT_DTYPE)
    b = np.array(bounds, np.float32)
    xyr = np.ascontiguousarray(xyr, np.float32).reshape(-1, 3)
    lv = np.ascontiguousarray(levels, np.int32).reshape(-1, 2)
    cap = max(len(keys), 1) * len(xyr)
    start, out = np.zeros(len(xyr) + 1, np.int32), np.zeros(max(cap, 1), np.int32)
    f = ref_frame_lib().plviref_frame_features_in_area
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    f(_p(keys), len(keys), _p(b), _p(xyr), _p(lv), len(xyr), _p(start), _p(out), cap)
    return [out[start[i]:start[i + 1]].copy() for i in range(len(xyr))]


def features_in_area(keys, grid, xyr, levels):
    """The oracle's restatement of Frame::GetFeaturesInArea (plvio_grid_*), same arguments / result as ref_features_in_area."""
    keys = np.ascontiguousarray(keys, KEYPOINT_DTYPE)
    g = _grid_floats(grid)
    L = lib()
    L.plvio_grid_create.restype = C.c_void_p
    L.plvio_grid_create.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
    L.plvio_grid_destroy.argtypes = [C.c_void_p]
    L.plvio_grid_features_in_area.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
    h = L.plvio_grid_create(_p(keys), len(keys), *g)
    out = []
    tmp = np.zeros(max(len(keys), 1), np.int32)
    xyr = np.ascontiguousarray(xyr, np.float32).reshape(-1, 3)
    lv = np.ascontiguousarray(levels, np.int32).reshape(-1, 2)
    for (x, y, r), (a, b) in zip(xyr, lv):
        k = L.plvio_grid_features_in_area(h, C.c_float(x), C.c_float(y), C.c_float(r), int(a), int(b), _p(tmp), len(tmp))
        out.append(tmp[:k].copy())
    L.plvio_grid_destroy(h)
    return out


def ref_line_descriptor_mad(d0, d1):
    """The reference's Frame::lineDescriptorMAD itself on kNN-2 distances: (nn_mad, nn12_mad)."""
    d0, d1 = np.ascontiguousarray(d0, np.int32), np.ascontiguousarray(d1, np.int32)
    a, b = C.c_double(0), C.c_double(0)
    f = ref_frame_lib().plviref_frame_line_descriptor_mad
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    f(_p(d0), _p(d1), len(d0), C.byref(a), C.byref(b))
    return a.value, b.value


def line_descriptor_mad(d0, d1):
    d0, d1 = np.ascontiguousarray(d0, np.int32), np.ascontiguousarray(d1, np.int32)
    a, b = C.c_double(0), C.c_double(0)
    f = lib().plvio_line_descriptor_mad
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    f(_p(d0), _p(d1), len(d0), C.byref(a), C.byref(b))
    return a.value, b.value


def _cam_f32(cam):
    K = np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float32)
    return K, np.asarray(cam["dist"], np.float32)


def ref_undistort_keypoints(keys, cam=None):
    """The reference's Frame::UndistortKeyPoints itself (cv::undistortPoints = the oracle's cv2-pinned restatement)."""
    keys = np.ascontiguousarray(keys, KEYPOINT_DTYPE)
    K, d = _cam_f32(cam or EUROC_CAMERA)
    out = np.zeros(max(len(keys), 1), KEYPOINT_DTYPE)
    f = ref_frame_lib().plviref_frame_undistort_keypoints
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    f(_p(keys), len(keys), _p(K), _p(d), len(d), _p(out))
    return out[:len(keys)]


def ref_undistort_keylines(keylines, cam=None):
    """The reference's Frame::UndistortKeyLines itself."""
    kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE)
    K, d = _cam_f32(cam or EUROC_CAMERA)
    out = np.zeros(max(len(kl), 1), KEYLINE_DTYPE)
    f = ref_frame_lib().plviref_frame_undistort_keylines
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    f(_p(kl), len(kl), _p(K), _p(d), len(d), _p(out))
    return out[:len(kl)]


def ref_stereo_matches(kps_l, desc_l, kps_r, desc_r, pyr_l, pyr_r, scale_factors, mb, mbf):
    """The reference's Frame::ComputeStereoMatches itself; arguments / result as stereo_matches."""
    kl, kr = np.ascontiguousarray(kps_l, KEYPOINT_DTYPE), np.ascontiguousarray(kps_r, KEYPOINT_DTYPE)
    dl, dr = np.ascontiguousarray(desc_l, np.uint8), np.ascontiguousarray(desc_r, np.uint8)
    lw = np.array([p.shape[1] for p in pyr_l], np.int32)
    lh = np.array([p.shape[0] for p in pyr_l], np.int32)
    pl = [np.ascontiguousarray(p, np.uint8) for p in pyr_l]
    pr = [np.ascontiguousarray(p, np.uint8) for p in pyr_r]
    ptr_l = (C.c_void_p * len(pl))(*[p.ctypes.data for p in pl])
    ptr_r = (C.c_void_p * len(pr))(*[p.ctypes.data for p in pr])
    sf = np.ascontiguousarray(scale_factors, np.float32)
    ur, dp = np.empty(max(len(kl), 1), np.float32), np.empty(max(len(kl), 1), np.float32)
    f = ref_frame_lib().plviref_frame_compute_stereo_matches
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    n = f(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), ptr_l, ptr_r, _p(lw), _p(lh), len(lw), _p(sf), C.c_float(mb), C.c_float(mbf),
          _p(ur), _p(dp))
    return ur[:len(kl)], dp[:len(kl)], n


def _csr_lists(start, out):
    return [out[start[i]:start[i + 1]].copy() for i in range(len(start) - 1)]


def ref_keyframe_features_in_area(keys, bounds, xyr):
    """The reference's KeyFrame::GetFeaturesInArea itself (src/KeyFrame.cc:1200-1244; the keyframe is built by the reference's
    own KeyFrame(Frame&, Map*, KeyFrameDatabase*) from a frame holding `keys`): list of index arrays."""
    keys = np.ascontiguousarray(keys, KEYPOINT_DTYPE)
    b = np.array(bounds, np.float32)
    xyr = np.ascontiguousarray(xyr, np.float32).reshape(-1, 3)
    cap = max(len(keys), 1) * len(xyr)
    start, out = np.zeros(len(xyr) + 1, np.int32), np.zeros(max(cap, 1), np.int32)
    f = ref_frame_lib().plviref_keyframe_features_in_area
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    f(_p(keys), len(keys), _p(b), _p(xyr), len(xyr), _p(start), _p(out), cap)
    return _csr_lists(start, out)


def ref_keyframe_lines_in_area(keylines, bounds, q5):
    """The reference's KeyFrame::GetLinesInArea(x1, y1, x2, y2, r) itself (src/KeyFrame.cc:1170-1198): list of index arrays."""
    kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE)
    b = np.array(bounds, np.float32)
    q5 = np.ascontiguousarray(q5, np.float32).reshape(-1, 5)
    cap = max(len(kl), 1) * len(q5)
    start, out = np.zeros(len(q5) + 1, np.int32), np.zeros(max(cap, 1), np.int32)
    f = ref_frame_lib().plviref_keyframe_lines_in_area
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    f(_p(kl), len(kl), _p(b), _p(q5), len(q5), _p(start), _p(out), cap)
    return _csr_lists(start, out)


def lines_in_area(keylines, q5):
    """The oracle's restatement of KeyFrame::GetLinesInArea (plvio_lines_in_area)."""
    kl = np.ascontiguousarray(keylines, KEYLINE_DTYPE)
    q5 = np.ascontiguousarray(q5, np.float32).reshape(-1, 5)
    f = lib().plvio_lines_in_area
    f.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
    tmp = np.zeros(max(len(kl), 1), np.int32)
    out = []
    for q in q5:
        k = f(_p(kl), len(kl), *[C.c_float(v) for v in q], _p(tmp))
        out.append(tmp[:k].copy())
    return out


def ref_keyframe_line_descriptor_mad(d0, d1):
    """The reference's KeyFrame::lineDescriptorMAD itself (src/KeyFrame.cc:411-435)."""
    d0, d1 = np.ascontiguousarray(d0, np.int32), np.ascontiguousarray(d1, np.int32)
    a, b = C.c_double(0), C.c_double(0)
    f = ref_frame_lib().plviref_keyframe_line_descriptor_mad
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    f(_p(d0), _p(d1), len(d0), C.byref(a), C.byref(b))
    return a.value, b.value


# ---- the keyframe searches on the reference's own KeyFrame / Frame classes (libplvi_ref_frame.so) ----------------------
def ref_fuse_real(keys, desc, bounds, scale_factors, inv_level_sigma2, uv, level, flags, qdesc, th=3.0, sim3=False):
    """ref_fuse with ORBmatcher.cc compiled against the reference's own KeyFrame (built by its constructor from a Frame):
    KeyFrame::GetFeaturesInArea, IsInImage, GetMapPoint, AddMapPoint are the reference's code."""
    keys, desc, _, b, sf, uv, lv, fl, qdesc = _ref_projection_args(keys, desc, np.zeros(1, [("min_x", "<f4"), ("min_y", "<f4"), ("inv_w", "<f4"), ("inv_h", "<f4")]),
                                                                    bounds, scale_factors, uv, level, flags, qdesc)
    inv = np.ascontiguousarray(inv_level_sigma2, np.float32)
    bi = np.full(max(len(uv), 1), -1, np.int32)
    f = ref_frame_lib().plviref_orb_fuse_realkeyframe
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    n = f(_p(keys), _p(desc), len(keys), _p(b), _p(sf), _p(inv), len(sf), _p(uv), _p(lv), _p(fl), _p(qdesc), len(uv), C.c_float(th),
          int(sim3), _p(bi))
    return n, bi[:len(uv)]


def ref_search_by_projection_kf_real(keys, desc, bounds, scale_factors, uv, level, flags, qdesc, th, ratio_hamming=1.0, matched_in=None):
    """ref_search_by_projection_kf on the reference's own KeyFrame class."""
    keys, desc, _, b, sf, uv, lv, fl, qdesc = _ref_projection_args(keys, desc, np.zeros(1, [("min_x", "<f4"), ("min_y", "<f4"), ("inv_w", "<f4"), ("inv_h", "<f4")]),
                                                                    bounds, scale_factors, uv, level, flags, qdesc)
    mi = None if matched_in is None else np.ascontiguousarray(matched_in, np.uint8)
    mt = np.full(max(len(keys), 1), -1, np.int32)
    f = ref_frame_lib().plviref_orb_search_by_projection_kf_realkeyframe
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
    n = f(_p(keys), _p(desc), len(keys), _p(mi), _p(b), _p(sf), len(sf), _p(uv), _p(lv), _p(fl), _p(qdesc), len(uv), int(th),
          C.c_float(ratio_hamming), _p(mt))
    return n, mt[:len(keys)]


def ref_search_bow_kf_f_real(keys1, desc1, mp1, fv1, keys2, desc2, fv2, bounds, nnratio=0.7, check_ori=True):
    """ref_search_bow_kf_f on the reference's own KeyFrame and Frame classes."""
    keys1, keys2 = np.ascontiguousarray(keys1, KEYPOINT_DTYPE), np.ascontiguousarray(keys2, KEYPOINT_DTYPE)
    desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    mp1 = np.ascontiguousarray(mp1, np.uint8)
    a, fa = _fv_args(fv1)
    b, fb = _fv_args(fv2)
    bd = np.array(bounds, np.float32)
    mt = np.full(max(len(keys2), 1), -1, np.int32)
    f = ref_frame_lib().plviref_orb_search_by_bow_kf_f_real
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
    n = f(_p(keys1), _p(desc1), _p(mp1), len(keys1), *fa, _p(keys2), _p(desc2), len(keys2), *fb, _p(bd), C.c_float(nnratio),
          int(check_ori), _p(mt))
    return n, mt[:len(keys2)]


def _ref_pinhole_lib():
    global _ref_ph
    if _ref_ph is None:
        if ref_build() is None or not _REF_PH_LIB.exists():
            raise RuntimeError("oracle/_ref/libplvi_ref_pinhole.so is not built and /root/reference is absent")
        lib()
        _ref_ph = C.CDLL(str(_REF_PH_LIB))
    return _ref_ph


def ref_epipolar_constrain(kp1, kp2, t12, unc, K=(1.0, 1.0, 0.0, 0.0)):
    """The reference's Pinhole::epipolarConstrain itself (src/CameraModels/Pinhole.cpp:135-157) for keypoint pairs, R12 = I,
    translation t12, per-pair unc; both cameras with intrinsics K (unit intrinsics: F12 = [t12]x exactly).  Returns bool[n]."""
    kp1, kp2 = np.ascontiguousarray(kp1, KEYPOINT_DTYPE), np.ascontiguousarray(kp2, KEYPOINT_DTYPE)
    Kf, t, u = np.array(K, np.float32), np.array(t12, np.float32), np.ascontiguousarray(unc, np.float32)
    ok = np.zeros(max(len(kp1), 1), np.uint8)
    f = _ref_pinhole_lib().plviref_pinhole_epipolar_constrain
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(_p(kp1), _p(kp2), len(kp1), _p(Kf), _p(t), _p(u), _p(ok))
    return ok[:len(kp1)].astype(bool)


def epipolar_constrain(kp1, kp2, F12, unc):
    """The oracle's restatement of the test at the end of Pinhole::epipolarConstrain for a given F12 (plvio_epipolar_constrain)."""
    F = np.ascontiguousarray(F12, np.float32).reshape(9)
    f = lib().plvio_epipolar_constrain
    f.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_float]
    return np.array([bool(f(float(a["x"]), float(a["y"]), float(b["x"]), float(b["y"]), _p(F), float(s))) for a, b, s in zip(kp1, kp2, unc)])


def ref_pinhole_project(K, xyz):
    """The reference's Pinhole::project(cv::Point3f) and toK: (uv [n,2], K 3x3)."""
    Kf = np.array(K, np.float32)
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    uv, Ko = np.zeros((len(xyz), 2), np.float32), np.zeros(9, np.float32)
    f = _ref_pinhole_lib().plviref_pinhole_project
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    f(_p(Kf), _p(xyz), len(xyz), _p(uv), _p(Ko))
    return uv, Ko.reshape(3, 3)


# ---- stereo line search (Frame::ComputeStereoMatches_Lines -> LineMatcher::matchGrid) -------------------------
_MG_ARGS = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]


def _match_grid_call(f, seg1, d1, seg2, d2, inv_width, inv_height, grid_rows, grid_cols, window):
    f.argtypes = _MG_ARGS
    f.restype = C.c_int
    seg1 = np.ascontiguousarray(seg1, np.float32).reshape(-1, 4)
    seg2 = np.ascontiguousarray(seg2, np.float32).reshape(-1, 4)
    d1, d2 = _ref_descs(d1, d2)
    assert len(seg1) == len(d1) and len(seg2) == len(d2)
    m = np.full(max(len(d1), 1), -1, np.int32)
    n = f(_p(seg1), _p(d1), len(d1), _p(seg2), _p(d2), len(d2), inv_width, inv_height, grid_rows, grid_cols,
          *[int(v) for v in window], _p(m))
    return n, m[:len(d1)]


def line_match_grid(seg1, d1, seg2, d2, inv_width, inv_height, grid_rows=48, grid_cols=64, window=(7, 0, 2, 2)):
    """Oracle restatement of the line search of Frame::ComputeStereoMatches_Lines (src/Frame.cc:1421-1448: grid fill
    along LineIterator, window (7, 0) x (2, 2)) + LineMatcher::matchGrid (src/LineMatcher.cpp:191-272).
    seg = (startPointX, startPointY, endPointX, endPointY) in pixels; window = (left, right, up, down) cells.
    Returns (nmatches, matches12)."""
    return _match_grid_call(lib().plvio_line_match_grid, seg1, d1, seg2, d2, inv_width, inv_height, grid_rows,
                            grid_cols, window)


def ref_line_match_grid(seg1, d1, seg2, d2, inv_width, inv_height, grid_rows=48, grid_cols=64, window=(7, 0, 2, 2)):
    """The same search through the reference's own GridStructure / LineIterator / LineMatcher::matchGrid
    (oracle/_ref/libplvi_ref.so, oracle/ref_glue_linematcher.cpp)."""
    return _match_grid_call(ref_lib().plviref_line_match_grid, seg1, d1, seg2, d2, inv_width, inv_height, grid_rows,
                            grid_cols, window)


# ---- the consumer's call pattern on the reference's own Frame class (ref_glue_frame.cpp: plviref_track_*) -------------
KEYLINE_BYTES, KEYPOINT_BYTES = 68, 28


def line_stereo_depth(seg1, seg2, matches12, mbf, seg1_un=None):
    """The disparity / overlap / depth filter after the search in Frame::ComputeStereoMatches_Lines (src/Frame.cc:1453-1500):
    (count, disparity [n1, 2] f32, depth [n1, 2] f32, mvle_l [n1, 3] f64)."""
    s1 = np.ascontiguousarray(seg1, np.float32).reshape(-1, 4)
    s2 = np.ascontiguousarray(seg2, np.float32).reshape(-1, 4)
    su = s1 if seg1_un is None else np.ascontiguousarray(seg1_un, np.float32).reshape(-1, 4)
    m = np.ascontiguousarray(matches12, np.int32)
    n1 = len(s1)
    disp = np.zeros((max(n1, 1), 2), np.float32); dep = np.zeros((max(n1, 1), 2), np.float32); le = np.zeros((max(n1, 1), 3), np.float64)
    f = lib().plvio_line_stereo_depth
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    k = f(_p(s1), n1, _p(s2), len(s2), _p(m), _p(su), C.c_float(mbf), _p(disp), _p(dep), _p(le))
    return k, disp[:n1], dep[:n1], le[:n1]


class RefTracker:
    """Frames built by the reference's own constructor Frame::Frame(imGray, ..., ORBextractor*, Lineextractor*, ...)
    (src/Frame.cc:537-642) and the two searches of Tracking::TrackWithMotionModelWithLines (src/Tracking.cc:3957,3990).
    Inside `with dropin():` the extractors / matchers underneath are the product's (libplvi_cuda.so)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, lsd_nfeatures=200, lsd_refine=0,
                 lsd_scale=0.8, line_levels=2, line_scale=2.0):
        self._l = ref_frame_lib()
        f = self._l.plviref_track_create
        f.restype = C.c_void_p
        f.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float]
        self._h = f(nfeatures, scale_factor, nlevels, ini_th, min_th, lsd_nfeatures, lsd_refine, lsd_scale, line_levels, line_scale)
        self._l.plviref_track_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        self._l.plviref_track_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        self._l.plviref_track_search.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int,
                                                 C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        self._l.plviref_track_destroy.argtypes = [C.c_void_p]

    def close(self):
        if self._h:
            self._l.plviref_track_destroy(self._h)
            self._h = None

    def frame(self, img, K=(458.654, 457.296, 367.215, 248.375), dist=(-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)):
        img = _u8(img)
        K = np.asarray(K, np.float32)
        d = np.asarray(dist, np.float32)
        return self._l.plviref_track_frame(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(K), _p(d), len(d))

    def get(self, field, which=0):
        buf = np.zeros(4 << 20, np.uint8)
        n = self._l.plviref_track_get(self._h, which, field, _p(buf), buf.nbytes)
        if n < 0:
            raise RuntimeError(f"plviref_track_get({field}) -> {n}")
        size = {0: KEYPOINT_BYTES, 1: KEYPOINT_BYTES, 2: 32, 3: KEYLINE_BYTES, 4: KEYLINE_BYTES, 5: 32, 6: 24}.get(field)
        return buf[: n * size].copy() if size else buf[:n].copy()

    def members(self, which=0):
        """Every extracted member of the frame as raw bytes, keyed by the reference's member name."""
        names = {0: "mvKeys", 1: "mvKeysUn", 2: "mDescriptors", 3: "mvKeys_Line", 4: "mvKeysUn_Line", 5: "mDescriptors_Line",
                 6: "mvKeyLineFunctions", 7: "mGrid", 8: "bounds+grid", 9: "scale tables", 10: "line scale tables", 11: "counts"}
        return {v: self.get(k, which) for k, v in names.items()}

    def search(self, th=15.0, mono=True, depth=None, obs0=None, t=(0.0, 0.0, 0.0), nnratio=0.9, check_ori=True, line_nnr=0.9):
        n_cur = int(np.frombuffer(self.get(11, 0), np.int32)[0])
        n_last_l = int(np.frombuffer(self.get(11, 1), np.int32)[1])
        pts = np.full(max(n_cur, 1), -9, np.int32)
        m12 = np.full(max(n_last_l, 1), -9, np.int32)
        nl = C.c_int(0)
        d = None if depth is None else np.ascontiguousarray(depth, np.float32)
        o = None if obs0 is None else np.ascontiguousarray(obs0, np.uint8)
        tt = np.asarray(t, np.float32)
        k = self._l.plviref_track_search(self._h, th, int(mono), _p(d), _p(o), _p(tt), nnratio, int(check_ori), line_nnr, _p(pts), _p(m12),
                                         C.byref(nl))
        return k, pts[:n_cur], nl.value, m12[:n_last_l]


def ref_search_frame_stereo(keys2, desc2, uright2, bounds, scale_factors, keys1, depth, flags, qdesc, K, mbf, t, th, check_ori=True,
                            blocked=None):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono=false) on a rectified-stereo frame
    (plviref_orb_search_by_projection_frame_stereo) -> (nmatches, match_train)."""
    keys2 = np.ascontiguousarray(keys2, KEYPOINT_DTYPE); keys1 = np.ascontiguousarray(keys1, KEYPOINT_DTYPE)
    desc2 = np.ascontiguousarray(desc2, np.uint8); qdesc = np.ascontiguousarray(qdesc, np.uint8)
    ur = np.ascontiguousarray(uright2, np.float32); b = np.asarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    dp = np.ascontiguousarray(depth, np.float32); fl = np.ascontiguousarray(flags, np.int32); Kf = np.asarray(K, np.float32)
    tt = np.asarray(t, np.float32)
    blk = None if blocked is None else np.ascontiguousarray(blocked, np.uint8)
    mt = np.full(max(len(keys2), 1), -9, np.int32)
    f = ref_frame_lib().plviref_orb_search_by_projection_frame_stereo
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
    k = f(_p(keys2), _p(desc2), _p(ur), len(keys2), _p(blk), _p(b), _p(sf), len(sf), _p(keys1), len(keys1), _p(dp), _p(fl), _p(qdesc), _p(Kf),
          float(mbf), _p(tt), float(th), int(check_ori), _p(mt))
    return k, mt[:len(keys2)]


def ref_fuse_stereo(keys, desc, uright, bounds, scale_factors, inv_level_sigma2, uv, depth, level, flags, qdesc, K, mbf, th=3.0):
    """ORBmatcher::Fuse(pKF, vpMapPoints, th) on a keyframe with mvuRight >= 0 observations -> (nFused, best_idx)."""
    keys = np.ascontiguousarray(keys, KEYPOINT_DTYPE); desc = np.ascontiguousarray(desc, np.uint8); qdesc = np.ascontiguousarray(qdesc, np.uint8)
    ur = np.ascontiguousarray(uright, np.float32); b = np.asarray(bounds, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    s2 = np.ascontiguousarray(inv_level_sigma2, np.float32); uvf = np.ascontiguousarray(uv, np.float32); dp = np.ascontiguousarray(depth, np.float32)
    lv = np.ascontiguousarray(level, np.int32); fl = np.ascontiguousarray(flags, np.int32); Kf = np.asarray(K, np.float32)
    nq = len(lv)
    best = np.full(max(nq, 1), -9, np.int32)
    f = ref_frame_lib().plviref_orb_fuse_stereo
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    k = f(_p(keys), _p(desc), _p(ur), len(keys), _p(b), _p(sf), _p(s2), len(sf), _p(uvf), _p(dp), _p(lv), _p(fl), _p(qdesc), nq, _p(Kf),
          float(mbf), float(th), _p(best))
    return k, best[:nq]


def ref_frame_stereo_lines(kl_left, desc_left, kl_right, desc_right, inv_width, inv_height, mbf):
    """Frame::ComputeStereoMatches_Lines (src/Frame.cc:1408-1529), unmodified -> (count, [n_left, 4] = disp_s, disp_e,
    depth_s, depth_e, mvle_l [n_left, 3])."""
    kl = np.ascontiguousarray(kl_left); kr = np.ascontiguousarray(kl_right)
    dl = np.ascontiguousarray(desc_left, np.uint8); dr = np.ascontiguousarray(desc_right, np.uint8)
    out = np.zeros((max(len(kl), 1), 4), np.float32)
    le = np.zeros((max(len(kl), 1), 3), np.float64)
    f = ref_frame_lib().plviref_frame_stereo_lines
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_float, C.c_void_p, C.c_void_p]
    k = f(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), float(inv_width), float(inv_height), float(mbf), _p(out), _p(le))
    return k, out[:len(kl)], le[:len(kl)]
