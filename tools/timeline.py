"""Diagnostic: start / end of every extractor kernel of one un-serialised front-end step (points and lines on their
own streams, as in the timed bench run), from CUDA events recorded behind each launch (PLVI_TIMELINE=1).

    PLVI_TIMELINE=1 python tools/timeline.py [--batch 4096] [--distinct 64]
"""
import argparse
import os
import sys
from pathlib import Path

os.environ["PLVI_TIMELINE"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--distinct", type=int, default=128)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--host", action="store_true", help="steps through the host-buffer entry points (FrontEnd.step_host: pinned frames in, pinned results out)")
    args = ap.parse_args()
    from pl_vi_orbslam3_b200 import synth
    W, H = 752, 480
    affine = synth.warp_affine(W, H).astype(np.float32).reshape(6)
    d = min(args.distinct, args.batch)
    base = synth.pair_batch(d, W, H, base_seed=0, workers=os.cpu_count() or 1)
    frames = np.concatenate([base] * ((args.batch + d - 1) // d))[: args.batch]
    import torch
    from pl_vi_orbslam3_b200.capi import lib
    from pl_vi_orbslam3_b200.frontend import FrontEnd
    fe = FrontEnd(args.batch, w=W, h=H, pairs=True, affine=affine, out_sets=2)
    st = fe.stream
    h_frames = torch.from_numpy(frames).pin_memory() if args.host else None
    ios = [fe.alloc_host_io() for _ in range(2)] if args.host else None

    def one_step(i):
        if args.host:
            fe.step_host(h_frames, ios[i % 2])
        else:
            fe.step(d_frames)

    with torch.cuda.stream(st):
        d_frames = torch.from_numpy(frames).to(fe.device)
        for i in range(2):
            one_step(i)
    st.synchronize()
    if args.host:
        fe.sync_host()
    fe.set_profile(True)
    rows = []
    e = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(st):
        e[0].record(st)
        for i in range(args.steps):
            one_step(i)
            e[i + 1].record(st)
            if i == args.steps - 1:
                st.synchronize()
                fe.line_stream.synchronize()
                if args.host:
                    fe.sync_host()
                for tag, txt in (("orb", lib().plvi_orb_profile(fe.orb._h).decode()),
                                 ("line", lib().plvi_line_profile(fe.line._h).decode())):
                    for item in txt.split(";"):
                        if "=" in item:
                            k, v = item.split("=")
                            a, b = v.split(",")
                            rows.append((float(a), float(b), tag, k))
    torch.cuda.synchronize()
    print("steps ms:", [round(e[i].elapsed_time(e[i + 1]), 2) for i in range(args.steps)])
    t0 = min(r[0] for r in rows)
    for a, b, tag, k in sorted(rows):
        print(f"{tag:5s} {k:18s} {a - t0:8.2f} -> {b - t0:8.2f}  ({b - a:7.2f} ms)")
    if fe._mev:
        print("search ms", fe._mev[0].elapsed_time(fe._mev[1]), "line match ms", fe._mev[1].elapsed_time(fe._mev[2]) if fe.lm else None)


if __name__ == "__main__":
    main()
