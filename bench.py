#!/usr/bin/env python3
"""Benchmark of the front-end hot path (BASELINE.json): frames/s on synthetic 752x480
EuRoC-shaped frames through ORB (1000 feats, 8 levels, 1.2, FAST 20/7) + LSD/LBD lines
(200 lines, refine 0, lsd_scale 0.8, 2 levels) + frame-to-frame Hamming matching
(SearchByProjection semantics for points, LineMatcher::match for lines).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path, one JSON line
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port) on host cores

One "step" = one batch of B frames per GPU through the whole path.  `value` is measured
with the frames already resident in HBM; `e2e` includes the pinned-host -> device copy of
every frame and the device -> host read of every result.  Frames shard across ranks by
contiguous range with no collective (weak scaling: B frames per GPU per step).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H = 752, 480
WORKLOAD = "C2+C3: 752x480, ORB 1000f/8lvl/1.2 FAST20/7 + LSD/LBD 200 lines (refine0, 0.8, 2 lvl) + frame-to-frame Hamming match"

# Algorithmic (compulsory) bytes per frame of each kernel at 752x480 -- DESIGN.md section 4.
P_ORB = 1117367                       # pyramid pixels, 8 levels
P_LSD = 602 * 384 + 301 * 192         # LSD working pixels, 2 octaves
P_RAW = 752 * 480 + 376 * 240         # line pyramid pixels
ALGO_BYTES = {
    "k_resize": 1089227 + 756407 + 360960 + 90240,
    "k_fast": P_ORB + 8 * 13000,
    "k_octree": 13000 * 6 + 1021 * 4,
    "k_blur7": 2 * P_ORB,
    "k_layout": 1021 * 8,
    "k_orient_desc": 1021 * (709 + 512 + 60),
    "k_lsd_rowfilter": P_RAW * (1 + 8),
    "k_lsd_scale_grad": P_RAW * 8 + P_LSD * (4 + 16 + 8) + P_LSD // 8,
    "k_lsd_grow": P_LSD * 4 + (P_LSD // 2) * (8 + 4) + P_LSD // 8,
    # band speculation: private bitmap copies (rows below each band's first row: 8.5x / 2.5x the octave bitmaps)
    # + the zeroed phantom bitmap
    "k_lsd_spec_init": (P_LSD // 8) * 2 + 267000,
    # every defined pixel's cos/sin once (8 B) + its list entry (4 B), seeds: angle + f64-derived cos/sin (12 B) and a
    # 16 B record each (~16k regions per frame), the touched part of the private bitmaps read and written once
    "k_lsd_spec": (P_LSD // 2) * (8 + 4) + 16000 * (12 + 16) + 2 * (P_LSD // 8),
    # records + pixel lists of the speculative regions, the availability bitmap read and written once, the ~12 % of
    # the pixels that are re-grown serially (cos/sin + list entry) and the region table
    "k_lsd_commit": 16000 * 16 + (P_LSD // 2) * 4 + 2 * (P_LSD // 8) + (P_LSD // 16) * (8 + 4) + 2600 * 16,
    "k_lsd_rect": (P_LSD // 2) * (4 + 8) + 2600 * 32,
    "k_line_assemble": 2600 * 16 + 200 * 68,
    "k_gauss5": 2 * 752 * 480,
    "k_pyrdown": 752 * 480 + 376 * 240,
    "k_sobel": P_RAW * (1 + 4),
    "k_lbd_rows": 200 * 63 * 60 * 4 + 200 * (68 + 2048),
    "k_lbd_fold": 200 * (2048 + 32 + 24),
    "k_search(+queries)": 2 * 1021 * (28 + 32) + 1021 * 28 + 1021 * 8,
    "k_line_match": 2 * 200 * 32 + 200 * 4,
}


GROW_NOTE = ("LSD region growing is a serial dependency chain per band (k_lsd_spec) / per frame and octave (k_lsd_commit): "
             "bound by instruction latency along the chain, not by HBM (ncu: IPC 0.6-1.7, DRAM throughput < 10 %); "
             "see DESIGN.md section 4 and profiles/r01_k_lsd_spec.md")


SCALE_FACTORS = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2))]).astype(np.float32), dtype=np.float32)


def _cpu_pair_job(idx_frames):
    """CPU arm over a contiguous shard: every frame is extracted once (ORB + lines) and matched
    against the previous frame of the shard, like a sequential tracker."""
    import oracle
    from pl_vi_orbslam3_b200.capi import QUERY_DTYPE
    from pl_vi_orbslam3_b200.matchers import frame_grid
    frames = idx_frames
    prev_r = prev_l = None
    grid = frame_grid(0, W, 0, H)
    use_ref = oracle.ref_available()
    if use_ref:
        oracle.ref_set_monotone(False)   # plain malloc while timing
    for f in frames:
        # the reference's own ORBextractor / Lineextractor / ORBmatcher / LineMatcher code (oracle/_ref: sources compiled
        # unmodified, the matchers against stand-in Frame / MapPoint classes) when it was built, else the oracle port
        r = oracle.ref_orb_extract(f) if use_ref else oracle.orb_extract(f)
        l = oracle.ref_line_extract(f) if use_ref else oracle.line_extract(f)
        if prev_r is not None:
            k = prev_r["keypoints"]
            q = np.zeros(len(k), QUERY_DTYPE)
            q["u"], q["v"] = k["x"], k["y"]
            q["radius"] = np.float32(15.0) * (np.float32(1.2) ** k["octave"].astype(np.float32))
            q["min_level"], q["max_level"], q["angle"] = k["octave"] - 1, k["octave"] + 1, k["angle"]
            if use_ref:
                # SearchByProjection(CurrentFrame, LastFrame, 15, true) with the identity pose: every tracked point of the
                # last frame projects onto its own position
                oracle.ref_search_frame(r["keypoints"], r["descriptors"], grid, (0.0, float(W), 0.0, float(H)), SCALE_FACTORS,
                                        k, np.stack([k["x"], k["y"]], 1), np.zeros(len(k), np.int32), prev_r["descriptors"], 15.0, True)
                if len(prev_l["descriptors"]) >= 2 and len(l["descriptors"]) >= 2:
                    oracle.ref_line_match(prev_l["descriptors"], l["descriptors"], 0.9, "match")
            else:
                oracle.search_frame(r["keypoints"], r["descriptors"], grid, q, prev_r["descriptors"], 100, True)
                oracle.line_match(prev_l["descriptors"], l["descriptors"], 0.9)
        prev_r, prev_l = r, l
    return len(frames)


def cpu_kind():
    import oracle
    if oracle.ref_available():
        return "reference", ("oracle/_ref: the reference's own ORBextractor.cc / LSD/lsd.cpp / LSDDetector_custom.cpp / "
                             "binary_descriptor_custom.cpp / LineExtractor.cc compiled unmodified with g++ -O2; the OpenCV "
                             "primitives underneath (resize, GaussianBlur, FAST, pyrDown, Sobel) are the oracle's scalar "
                             "models, not OpenCV's SIMD code, so this understates a real OpenCV build; the two searches "
                             "are the reference's own ORBmatcher::SearchByProjection(Frame, Frame) and LineMatcher::match "
                             "(ORBmatcher.cc / LineMatcher.cpp compiled unmodified against stand-in Frame / MapPoint classes)")
    return "port", "oracle/ C++ port of the reference path (oracle/_ref not built)"


def cpu_arm(frames, cores):
    """Times the oracle port over `frames` with `cores` worker processes (contiguous shards).
    Returns frames/s.  Must run before CUDA is initialised in this process (fork)."""
    import multiprocessing as mp
    import oracle
    oracle.build()
    shards = [s for s in np.array_split(frames, cores) if len(s)]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(shards)) as pool:
        pool.map(_cpu_pair_job, [s[:1] for s in shards])  # warm-up: library load, page-in
        t0 = time.perf_counter()
        done = sum(pool.map(_cpu_pair_job, shards))
        dt = time.perf_counter() - t0
    return done / dt, dt


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("PLVI_BENCH_BATCH", 4096)))
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames in the CPU baseline sample (0 = 32 per core: 10-30 s of CPU work)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--orb-only", action="store_true", help="diagnostic: time ORB extraction alone (not the bench metric)")
    ap.add_argument("--profile-out", default="")
    ap.add_argument("--no-overlap", action="store_true", help="run the line pipeline on the same stream as ORB")
    ap.add_argument("--line-priority", type=int, default=0)
    ap.add_argument("--pipe-mode", choices=("slice", "alternate"), default=os.environ.get("PLVI_BENCH_PIPE_MODE", "alternate"),
                    help="alternate: whole batches go to the pipelines in turn (consecutive batches in flight at "
                         "different phases); slice: every batch is split across the pipelines")
    ap.add_argument("--out-sets", type=int, default=2, help="alternating output buffer sets per pipeline (e2e: the D2H of "
                    "step i overlaps the compute of step i+1)")
    ap.add_argument("--pipes", type=int, default=int(os.environ.get("PLVI_BENCH_PIPES", 1)),
                    help="independent pipelines the batch is split over (overlap across slices)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    cores = os.cpu_count() or 1

    from pl_vi_orbslam3_b200 import build as _build, synth
    if not _build.LIB.exists():        # the built library normally travels with the repo snapshot
        _build.build(force=True)

    # ------------------------------------------------------------------ CPU arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        per_step = args.cpu_sample or min(32 * cores, 1024)
        frames = synth.frame_batch(per_step, W, H, base_seed=0, distinct=16)
        for _ in range(max(args.warmup, 0)):
            cpu_arm(frames[: max(cores, 2)], cores)
        t = 0.0
        done = 0
        for _ in range(args.steps):
            fps, dt = cpu_arm(frames, cores)
            t += dt
            done += len(frames)
        v = done / t
        print(json.dumps({
            "impl": "reference", "metric": "frames/s", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": per_step},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": cpu_kind()[0],
                             "sample": f"{per_step} synthetic frames per step, contiguous shards over {cores} worker processes; "
                                       + cpu_kind()[1]},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ------------------------------------------------------------------ CUDA arm
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nsamp = args.cpu_sample or min(32 * cores, 1024)
        fps, dt = cpu_arm(synth.frame_batch(nsamp, W, H, base_seed=0, distinct=16), cores)
        cpu_base = {"value": fps, "unit": "frames/s", "cores": cores, "kind": cpu_kind()[0],
                    "sample": f"{nsamp} synthetic 752x480 frames (same generator), contiguous shards over {cores} worker "
                              f"processes, {dt:.1f} s wall; " + cpu_kind()[1]}

    import torch
    import torch.distributed as dist
    from pl_vi_orbslam3_b200.frontend import PipelinedFrontEnd

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    frames = synth.frame_batch(B, W, H, base_seed=1000 * rank, distinct=16)
    h_frames = torch.from_numpy(frames).pin_memory()
    fe = PipelinedFrontEnd(B, pipes=args.pipes, mode=args.pipe_mode, device=local_rank, w=W, h=H, with_lines=not args.orb_only,
                           with_match=not args.orb_only, overlap_lines=not args.no_overlap,
                           line_priority=args.line_priority, out_sets=args.out_sets)
    st = fe.stream
    with torch.cuda.stream(st):
        d_frames = h_frames.to(dev, non_blocking=True)
    st.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value")
    with torch.cuda.stream(st):
        for _ in range(args.warmup):
            fe.step(d_frames)
        fe.drain()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(args.steps):
            launches += fe.step(d_frames)
        fe.drain()
        e1.record(st)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)

    # ---- end to end: pinned host frames in, every result back on the host.  Copies run on their own
    # streams: the H2D of step i+1 overlaps the compute of step i (two input buffers); the D2H of step
    # i must finish before step i+1 overwrites the output buffers.
    alt = args.pipe_mode == "alternate"
    npipe = fe.alive_steps                     # steps whose outputs are alive at the same time
    outs = fe.outputs()
    h_out = [{k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in o.items()} for o in outs]
    nbuf = npipe + 1
    d_in = [torch.empty_like(d_frames) for _ in range(nbuf)]
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def e2e_run(nsteps):
        ev_h2d = [torch.cuda.Event() for _ in range(nsteps)]
        ev_cmp = [None] * nsteps
        ev_d2h = [torch.cuda.Event() for _ in range(nsteps)]
        start = torch.cuda.Event(enable_timing=True)
        stop = torch.cuda.Event(enable_timing=True)
        start.record(st)
        s_in.wait_event(start)
        s_out.wait_event(start)
        for i in range(nsteps):
            with torch.cuda.stream(s_in):
                if i >= nbuf:
                    s_in.wait_event(ev_cmp[i - nbuf])       # input buffer i % nbuf was read by step i - nbuf
                d_in[i % nbuf].copy_(h_frames, non_blocking=True)
                ev_h2d[i].record(s_in)
            with torch.cuda.stream(st):
                # step i overwrites the output buffers of step i - npipe: those must be on the host
                waits = [ev_h2d[i]] + ([ev_d2h[i - npipe]] if i >= npipe else [])
                fe.step(d_in[i % nbuf], wait=waits)
                ev = fe.done_event()
                if ev is None:
                    ev = torch.cuda.Event()
                    ev.record(st)
                else:                                       # the event object is reused by step i + npipe
                    own = torch.cuda.Event()
                    s_out.wait_event(ev)
                    own.record(s_out)
                    ev = own
                ev_cmp[i] = ev
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[i])
                for o, ho in zip(fe.outputs(), h_out):
                    for k, v in o.items():
                        ho[k].copy_(v, non_blocking=True)
                ev_d2h[i].record(s_out)
        st.wait_event(ev_d2h[nsteps - 1])
        fe.drain()
        stop.record(st)
        return start, stop

    e2e_run(2)
    barrier()
    f0, f1 = e2e_run(args.steps)
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    h2d = int(h_frames.numel())
    d2h = int(sum(v.numel() * v.element_size() for ho in h_out for v in ho.values()))

    # ---- per-kernel profile of one extra step (events after every launch; not part of the timed numbers)
    fe.set_profile(True)
    with torch.cuda.stream(st):
        fe.step(d_frames, serialize=True)   # points and lines back to back: per-kernel times without co-scheduling
    st.synchronize()
    prof = fe.profile()
    fe.set_profile(False)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    counts = float(np.mean([o["counts"].float().mean().item() for o in outs]))
    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        top = max(prof, key=prof.get) if prof else None
        roof = None
        traffic = None
        if top:
            tj = ROOT / "profiles" / f"r01_{top}_traffic.json"   # dram bytes from the committed ncu --set full capture
            if tj.exists():
                traffic = json.loads(tj.read_text())["dram_bytes_per_frame"] * B
            ach = ALGO_BYTES.get(top, 0) * B / (prof[top] * 1e-3) / 1e9
            roof = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "algorithmic_bytes_per_launch": ALGO_BYTES.get(top, 0) * B, "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
                    "kernel_ms_per_launch": prof[top], "frames_per_launch": B,
                    "note": GROW_NOTE if top in ("k_lsd_grow", "k_lsd_spec", "k_lsd_commit") else ""}
        total_prof = sum(prof.values()) or 1.0
        line = {
            "metric": "frames/s", "value": world * B * args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD if not args.orb_only else "DIAGNOSTIC orb-only", "frames_per_step_per_gpu": B,
                       "width": W, "height": H, "l2_policy": "inputs larger than L2 (%.0f MB of frames per step)" % (B * W * H / 1e6),
                       "parallelism": f"frame-sharded x{world}, no collective", "mean_keypoints": counts,
                       "pipelines_per_gpu": args.pipes, "pipeline_mode": args.pipe_mode},
            "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu_base,
            # every kernel against the HBM roofline: algorithmic GB/s = ALGO_BYTES x frames / its time in the serialised
            # profile pass, as a fraction of the measured copy peak (streaming kernels should be judged on this one;
            # k_lsd_spec / k_lsd_commit / k_octree / k_search are latency or issue bound)
            "kernel_hbm": {k: {"gbs": round(ALGO_BYTES[k] * B / (v * 1e-3) / 1e9, 1),
                               "frac": round(ALGO_BYTES[k] * B / (v * 1e-3) / 1e9 / peak, 4)}
                           for k, v in sorted(prof.items(), key=lambda kv: -kv[1]) if k in ALGO_BYTES and v > 0},
            "kernel_ms": {k: round(v, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])},
            "kernel_share": {k: round(v / total_prof, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])},
        }
        print(json.dumps(line))
        if args.profile_out:
            Path(args.profile_out).write_text(json.dumps(line, indent=1))
    fe.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
