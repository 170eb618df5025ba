#!/usr/bin/env python3
"""Generates tests/golden/*.npz (run in the build container only).

Two kinds of vectors:
  * cv2_*  : outputs of OpenCV 4.13 (python cv2, IPP off) for the primitives the reference
             calls (resize, GaussianBlur, FAST per cell, fastAtan2, pyrDown, Sobel, f64
             resize/blur) on the reference's own frames (data2/color/*.png, data/color/*.png)
             and on synthetic frames.  They pin the oracle's restatement of OpenCV.
  * oracle_*: outputs of the oracle for the reference-owned logic (octree, IC_Angle, rBRIEF,
             LSD, LBD, matchers).  The reference ships no expected outputs and cannot be
             built here, so these pin the oracle against regressions only ("parity unpinned").
Large arrays are stored as CRC32; small ones in full.
"""
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import cv2  # noqa: E402

import oracle  # noqa: E402
from pl_vi_orbslam3_b200 import synth  # noqa: E402

cv2.ipp.setUseIPP(False)
cv2.setNumThreads(1)
OUT = ROOT / "tests" / "golden"
REF = Path("/root/reference")


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def frames():
    fr = {}
    for i in (1, 3):
        fr[f"data2_{i}"] = cv2.imread(str(REF / f"data2/color/{i}.png"), cv2.IMREAD_UNCHANGED)
    fr["data_1_gray"] = cv2.cvtColor(cv2.imread(str(REF / "data/color/1.png")), cv2.COLOR_BGR2GRAY)
    fr["synth_0"] = synth.frame_euroc(0)
    fr["synth_7_640"] = synth.frame_euroc(7, 640, 480)
    return fr


def cv2_grid_fast(level, ini_th=20, min_th=7):
    """ComputeKeyPointsOctTree's cell loop (src/ORBextractor.cc:763-855) with cv2.FAST per cell."""
    h, w = level.shape
    minB, maxBX, maxBY = 16, w - 16, h - 16
    width, height = float(maxBX - minB), float(maxBY - minB)
    nCols, nRows = int(width / 30), int(height / 30)
    wCell, hCell = int(np.ceil(width / nCols)), int(np.ceil(height / nRows))
    fd = cv2.FastFeatureDetector_create(threshold=ini_th, nonmaxSuppression=True)
    out = []
    for i in range(nRows):
        iniY = minB + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minB + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            roi = np.ascontiguousarray(level[iniY:maxY, iniX:maxX])
            fd.setThreshold(ini_th)
            k = fd.detect(roi)
            if not k:
                fd.setThreshold(min_th)
                k = fd.detect(roi)
            for p in k:
                out.append((p.pt[0] + j * wCell, p.pt[1] + i * hCell, p.response))
    return np.array(out, np.float32).reshape(-1, 3)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    S = float(np.float32(0.8))
    sigma = 0.6 / S
    g = {}
    for name, img in frames().items():
        if not name.startswith("synth"):   # synthetic frames are regenerated from their seed
            np.savez_compressed(OUT / f"frame_{name}.npz", img=img)
        h, w = img.shape
        plan = oracle.orb_plan(w, h)
        cur = img
        for l in range(1, 8):
            cur = cv2.resize(cur, (int(plan["w"][l]), int(plan["h"][l])), interpolation=cv2.INTER_LINEAR)
            g[f"cv2_pyr_{name}_{l}"] = crc(cur)
            if l in (1, 4, 7):
                g[f"cv2_blur7_{name}_{l}"] = crc(cv2.GaussianBlur(cur, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
                c = cv2_grid_fast(cur)
                g[f"cv2_gridfast_{name}_{l}"] = c if len(c) < 1200 else np.array([len(c), crc(c)], np.int64)
        g[f"cv2_blur7_{name}_0"] = crc(cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
        c0 = cv2_grid_fast(img)
        g[f"cv2_gridfast_{name}_0"] = np.array([len(c0), crc(c0)], np.int64)
        half = cv2.resize(img, (w // 2, h // 2), interpolation=cv2.INTER_LINEAR)
        g[f"cv2_half_{name}"] = crc(half)
        b5 = cv2.GaussianBlur(img, (5, 5), 1)
        g[f"cv2_blur5_{name}"] = crc(b5)
        pd = cv2.pyrDown(b5, dstsize=(w // 2, h // 2))
        g[f"cv2_pyrdown_{name}"] = crc(pd)
        g[f"cv2_sobelx_{name}"] = crc(cv2.Sobel(b5, cv2.CV_16S, 1, 0, ksize=3))
        g[f"cv2_sobely_{name}"] = crc(cv2.Sobel(pd, cv2.CV_16S, 0, 1, ksize=3))
        f64 = img.astype(np.float64)
        gb = cv2.GaussianBlur(f64, (7, 7), sigma)
        g[f"cv2_blurf64_sample_{name}"] = gb[::37, ::41].copy()          # compared with tolerance 1e-12
        sc = cv2.resize(gb, None, fx=S, fy=S, interpolation=cv2.INTER_LINEAR)
        g[f"cv2_resizef64_of_oracleblur_{name}"] = crc(
            cv2.resize(oracle.gaussian_blur_f64(f64, oracle.gaussian_kernel_f64(7, sigma)), None, fx=S, fy=S,
                       interpolation=cv2.INTER_LINEAR))
        g[f"cv2_scaled_shape_{name}"] = np.array(sc.shape)
        # oracle-defined outputs (regression pins)
        r = oracle.orb_extract(img)
        g[f"oracle_orb_{name}"] = np.array([len(r["keypoints"]), r["mono_index"], crc(r["keypoints"]), crc(r["descriptors"])], np.int64)
        g[f"oracle_orb_head_{name}"] = r["keypoints"][:16]
        lr = oracle.line_extract(img)
        g[f"oracle_line_{name}"] = np.array([len(lr["keylines"]), *lr["raw_counts"], crc(lr["keylines"]), crc(lr["descriptors"])], np.int64)
        g[f"oracle_line_head_{name}"] = lr["keylines"][:8]
    g["cv2_gauss_kernel7"] = cv2.getGaussianKernel(7, sigma, cv2.CV_64F).ravel()
    rng = np.random.RandomState(5)
    yx = rng.randint(-50000, 50000, (4000, 2)).astype(np.float32)
    yx[::9, 0] = 0
    yx[::13, 1] = 0
    g["cv2_atan2_in"] = yx
    g["cv2_atan2_out"] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in yx], np.float32)
    # BFMatcher knn-2 semantics (ties -> lowest train index)
    d1 = rng.randint(0, 256, (60, 32)).astype(np.uint8)
    d2 = rng.randint(0, 256, (50, 32)).astype(np.uint8)
    d2[10] = d2[4]
    d2[20] = d1[3]
    d2[21] = d1[3]
    m = cv2.BFMatcher(cv2.NORM_HAMMING, False).knnMatch(d1, d2, 2)
    g["cv2_knn_d1"], g["cv2_knn_d2"] = d1, d2
    g["cv2_knn"] = np.array([[a.trainIdx, a.distance, b.trainIdx, b.distance] for a, b in m], np.float32)
    # matcher regression pins on the C3 pair
    f1, f2, A = synth.warp_pair(3)
    r1, r2 = oracle.orb_extract(f1), oracle.orb_extract(f2)
    l1, l2 = oracle.line_extract(f1), oracle.line_extract(f2)
    n, m12 = oracle.line_match(l1["descriptors"], l2["descriptors"], 0.9)
    g["oracle_c3_line_match"] = np.array([n, crc(m12)], np.int64)
    np.savez_compressed(OUT / "golden.npz", **g)
    print("wrote", OUT / "golden.npz", len(g), "entries;", sum(f.stat().st_size for f in OUT.glob("*.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
