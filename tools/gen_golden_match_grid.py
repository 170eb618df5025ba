"""Writes tests/golden/ref_match_grid.npz: outputs of the REFERENCE's own GridStructure / LineIterator /
LineMatcher::matchGrid (oracle/_ref/libplvi_ref.so, needs /root/reference) on the seeded cases of
tests/test_stereo_lines.py.  Run from the repo root: python tools/gen_golden_match_grid.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
from test_stereo_lines import GOLD, INV_H, INV_W, SIZES, stereo_line_case  # noqa: E402

out = {}
for k, (n1, n2) in enumerate(SIZES):
    s1, d1, s2, d2 = stereo_line_case(100 + k, n1, n2)
    n, m = oracle.ref_line_match_grid(s1, d1, s2, d2, INV_W, INV_H)
    out[f"n_{k}"], out[f"m12_{k}"] = np.int32(n), m
    print(k, n1, n2, n)
np.savez_compressed(GOLD, **out)
