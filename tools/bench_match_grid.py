"""Times k_line_match_grid (stereo line search, plvi_line_match_grid) on P stereo pairs of 200 x 200 lines resident
in HBM: python tools/bench_match_grid.py [P].  Wall clock around 20 launches + stream sync (launch overhead is
negligible at P >= 1024); prints one JSON line."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from pl_vi_orbslam3_b200 import LineMatcher  # noqa: E402
from pl_vi_orbslam3_b200.capi import check, lib, ptr  # noqa: E402
from test_stereo_lines import INV_H, INV_W, stereo_line_case  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
base = [stereo_line_case(s, 200, 200) for s in range(16)]
dev = torch.device("cuda", 0)
s1 = torch.from_numpy(np.stack([base[i % 16][0] for i in range(P)])).to(dev)
d1 = torch.from_numpy(np.stack([base[i % 16][1] for i in range(P)])).to(dev)
s2 = torch.from_numpy(np.stack([base[i % 16][2] for i in range(P)])).to(dev)
d2 = torch.from_numpy(np.stack([base[i % 16][3] for i in range(P)])).to(dev)
n = torch.full((P,), 200, dtype=torch.int32, device=dev)
m12 = torch.empty((P, 200), dtype=torch.int32, device=dev)
nm = torch.empty(P, dtype=torch.int32, device=dev)
lm = LineMatcher(max_pairs=1, max_train=256, max_query=256)
torch.cuda.synchronize(dev)


def run():
    check(lib().plvi_line_match_grid(lm._h, P, ptr(s1), ptr(d1), ptr(n), 200, ptr(s2), ptr(d2), ptr(n), 200, INV_W, INV_H,
                                     48, 64, 7, 0, 2, 2, ptr(m12), ptr(nm)))


for _ in range(3):
    run()
lm.sync()
t0 = time.perf_counter()
for _ in range(20):
    run()
lm.sync()
dt = (time.perf_counter() - t0) / 20
print(json.dumps({"kernel": "k_line_match_grid", "pairs": P, "lines_per_side": 200, "ms_per_launch": dt * 1e3,
                  "pairs_per_s": P / dt, "mean_matches": float(nm.float().mean().item())}))
lm.close()
