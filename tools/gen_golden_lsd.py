#!/usr/bin/env python3
"""Committed vectors for the round-2 additions of the line path (tests/golden/lsd_extra.npz), so that the checks of
tests/test_oracle_vs_cv2.py / test_oracle_vs_ref.py also run where cv2 and oracle/_ref are absent:
  * cv2.getGaussianKernel for the Gaussian sizes LSD uses at lsd_scale 0.5 / 0.6 / 0.9 (OpenCV's soft-float exp) and the
    CRC of cv2.resize(f64, fx = fy = lsd_scale) -- incl. the 2x2 area path OpenCV takes at exactly 0.5;
  * the reference's own lsd.cpp (oracle/_ref) raw segments for lsd_refine 1 / 2 and lsd_scale 1.0 / 0.6 on synthetic frames;
  * Frame::ComputeStereoMatches_Lines outputs (disparity, depth, mvle_l) of the reference's own Frame.cc.
Needs cv2 and /root/reference (build container only)."""
import math
import sys
import zlib
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
from pl_vi_orbslam3_b200 import synth  # noqa: E402

cv2.ipp.setUseIPP(False)
cv2.setNumThreads(1)
crc = lambda a: np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))
g = {}
img = synth.frame_euroc(9).astype(np.float64)
for scale in (0.5, 0.6, 0.9):
    S = float(np.float32(scale))
    sigma = 0.6 / S
    n = 1 + 2 * int(math.ceil(sigma * math.sqrt(2 * 3.0 * math.log(10.0))))
    g[f"cv2_kernel_{scale}"] = cv2.getGaussianKernel(n, sigma, cv2.CV_64F).ravel()
    mine = oracle.gaussian_blur_f64(img, oracle.gaussian_kernel_f64(n, sigma))
    ref = cv2.resize(mine, None, fx=S, fy=S, interpolation=cv2.INTER_LINEAR)
    g[f"cv2_resize_shape_{scale}"] = np.array(ref.shape)
    g[f"cv2_resize_crc_{scale}"] = crc(ref)
for seed, scale, refine in ((0, 0.8, 1), (0, 0.8, 2), (3, 0.6, 1), (3, 1.0, 2), (5, 1.0, 0), (5, 0.5, 0)):
    g[f"ref_lsd_{seed}_{scale}_{refine}"] = oracle.ref_lsd(synth.frame_euroc(seed), scale, refine)
import test_stereo_lines as SL  # noqa: E402
for seed in (1, 2, 9):
    s1, d1, s2, d2 = SL._depth_case(seed)
    k, out, le = oracle.ref_frame_stereo_lines(SL._keylines(s1), d1, SL._keylines(s2), d2, SL.INV_W, SL.INV_H, 47.9)
    g[f"ref_stereo_lines_{seed}_k"] = np.int32(k)
    g[f"ref_stereo_lines_{seed}_out"] = out
    g[f"ref_stereo_lines_{seed}_le"] = le
np.savez_compressed(ROOT / "tests" / "golden" / "lsd_extra.npz", **g)
print("wrote", len(g), "arrays")
