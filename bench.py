#!/usr/bin/env python3
"""Benchmark of the front-end hot path (BASELINE.json): frames/s on synthetic EuRoC-shaped frames through ORB
(8 levels, 1.2, FAST 20/7) + LSD/LBD lines (200 lines, refine 0, lsd_scale 0.8, 2 levels) + the Hamming searches.

  python bench.py --gpus N --steps K --warmup W            # our CUDA path, one JSON line
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU code (oracle/_ref) on the host cores

One "step" = one batch of B frames per GPU through the whole path.  `value` is measured with the frames already
resident in HBM; `e2e` goes through the reference-facing C ABI with HOST buffers (plvi_line_extract_batch_async /
plvi_orb_extract_batch_async: pinned host frames in, every result in pinned host memory -- the calls the C++ shim
makes).  Frames shard across ranks by contiguous range with no collective.

  --config c23     (default) BASELINE configs 2+3 at 752x480 / 1000 features: B frames per GPU = B/2 C3 pairs
                   (frame_euroc(s), its warp), every frame from its own seed; both frames of a pair are
                   extracted (ORB + LSD/LBD) and the pair is matched in the two point modes of C3
                   (SearchByProjection(Frame, Frame) with the warp in the place of the pose; SearchForInitialization)
                   plus LineMatcher::match.  Weak scaling.
  --config c4_640 / c4_1280    config 4: 640x480 / 1280x720 frames, 2000 ORB features, 256 frames per GPU, consecutive
                   frames matched under the identity pose.  Weak scaling.
  --config c5      config 5: one 4096-frame sequence (seeds 0...4095) sharded over the ranks (4096/N frames per GPU,
                   frontend.shard_range), consecutive frames matched.  Strong scaling.
  --batch 1        latency mode: one frame per call through the host-buffer ABI (ORB + lines, no pair to match).
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

CONFIGS = {
    #         w     h    nfeat  default batch   pairs  scaling
    "c23":     (752, 480, 1000, 4096, True, "weak"),
    "c4_640":  (640, 480, 2000, 256, False, "weak"),
    "c4_1280": (1280, 720, 2000, 256, False, "weak"),
    "c5":      (752, 480, 1000, 4096, False, "strong"),
}
WORKLOADS = {
    "c23": "C2+C3: 752x480, ORB 1000f/8lvl/1.2 FAST20/7 + LSD/LBD 200 lines (refine0, 0.8, 2 lvl) + frame-to-frame Hamming match",
    "c4_640": "C4: 640x480, ORB 2000f/8lvl/1.2 FAST20/7 + LSD/LBD 200 lines + frame-to-frame Hamming match, 256 frames per GPU",
    "c4_1280": "C4: 1280x720, ORB 2000f/8lvl/1.2 FAST20/7 + LSD/LBD 200 lines + frame-to-frame Hamming match, 256 frames per GPU",
    "c5": "C5: 4096-frame 752x480 sequence sharded over the GPUs, ORB 1000f/8lvl + LSD/LBD 200 lines + frame-to-frame Hamming match",
}
MATCH_TH, INIT_WINDOW, NNRATIO = 15.0, 100.0, 0.9


def algo_bytes(w, h, cap, lw, lh, ow, oh, sw, sh, lines=200):
    """Algorithmic (compulsory) bytes per frame of each kernel -- DESIGN.md section 4 / SURVEY.md 8(d): every stage
    reads its input once and writes its output once.  lw/lh: ORB pyramid level sizes, ow/oh: line pyramid octaves,
    sw/sh: LSD working sizes (after lsd_scale)."""
    P_ORB = int(sum(int(a) * int(b) for a, b in zip(lw, lh)))
    P_LSD = int(sum(int(a) * int(b) for a, b in zip(sw, sh)))
    P_RAW = int(sum(int(a) * int(b) for a, b in zip(ow, oh)))
    lvl = [int(a) * int(b) for a, b in zip(lw, lh)]
    cand = 13000 * P_ORB // 1117367     # FAST candidates per frame (13 k on a 752x480 EuRoC frame)
    regs = 16000 * P_LSD // 288960      # regions per frame
    segs = 2600 * P_LSD // 288960
    return {
        "k_resize": sum(lvl[:-1]) + sum(lvl[1:]) + int(ow[0]) * int(oh[0]) + (int(ow[1]) * int(oh[1]) if len(ow) > 1 else 0),
        "k_fast": P_ORB + 8 * cand,
        "k_octree": cand * 6 + cap * 4,
        "k_blur7": 2 * P_ORB,
        "k_layout": cap * 8,
        "k_orient_desc": cap * (709 + 512 + 60),
        "k_lsd_rowfilter": P_RAW * (1 + 8),
        "k_lsd_scale_grad": P_RAW * 8 + P_LSD * (4 + 16 + 8) + P_LSD // 8,
        "k_lsd_pre": P_RAW * 1 + P_LSD * (4 + 16 + 8) + P_LSD // 8,
        "k_lsd_grow": P_LSD * 4 + (P_LSD // 2) * (8 + 4) + P_LSD // 8,
        # band speculation: private bitmap copies (rows below each band's first row) + the zeroed phantom bitmap
        "k_lsd_spec_init": (P_LSD // 8) * 2 + 267000 * P_LSD // 288960,
        # every defined pixel's cos/sin once (8 B) + its list entry (4 B), seeds: angle + f64-derived cos/sin (12 B) and a
        # 16 B record each, the touched part of the private bitmaps read and written once
        "k_lsd_spec": (P_LSD // 2) * (8 + 4) + regs * (12 + 16) + 2 * (P_LSD // 8),
        # records + pixel lists of the speculative regions, the availability bitmap read and written once, the ~12 % of
        # the pixels that are re-grown serially (cos/sin + list entry) and the region table
        "k_lsd_commit": regs * 16 + (P_LSD // 2) * 4 + 2 * (P_LSD // 8) + (P_LSD // 16) * (8 + 4) + segs * 16,
        "k_lsd_band_rounds": P_LSD * 4 + (P_LSD // 2) * (8 + 4) + P_LSD // 8,
        "k_lsd_band_finish": (P_LSD // 2) * 4 + segs * 16,
        "k_lsd_rect": (P_LSD // 2) * (4 + 8) + segs * 32,
        "k_line_assemble": segs * 16 + lines * 68,
        "k_gauss5": 2 * int(ow[0]) * int(oh[0]),
        "k_pyrdown": P_RAW,
        "k_sobel": P_RAW * (1 + 4),
        "k_lbd_pre": P_RAW * (1 + 4),
        "k_lbd_rows": lines * 63 * 60 * 4 + lines * (68 + 2048),
        "k_lbd_fold": lines * (2048 + 32 + 24),
        "k_search(+queries)": 2 * cap * (28 + 32) + cap * 28 + cap * 8,
        "k_line_match": 2 * lines * 32 + lines * 4,
    }


GROW_NOTE = ("LSD region growing is a serial dependency chain per band (k_lsd_spec) / per frame and octave (k_lsd_commit): "
             "k_lsd_spec is bound by the latency of dependent L2 / DRAM round trips with 17 warps per SM (ncu at 4096 frames: issue "
             "active 32 %, L1 / L2 / DRAM throughput 29 / 29 / 15 %, DRAM traffic 7.7x the algorithmic bytes because every 32-byte "
             "sector of the neighbour records is fetched several times), k_lsd_commit half by instruction issue (64 %); neither by "
             "HBM bandwidth; see DESIGN.md section 4 and profiles/r02b_notes.md, profiles/r02c_notes.md (the step is the sum of its kernels)")


SCALE_FACTORS = np.cumprod(np.concatenate([[np.float32(1.0)], np.full(7, np.float32(1.2))]).astype(np.float32), dtype=np.float32)


# --------------------------------------------------------------------------------------------------- CPU arm
def _cpu_extract(oracle, use_ref, f, nfeat):
    if use_ref:
        return oracle.ref_orb_extract(f, nfeat), oracle.ref_line_extract(f)
    return oracle.orb_extract(f, nfeat), oracle.line_extract(f)


def _cpu_match(oracle, use_ref, grid, w, h, r1, l1, r2, l2, affine, stages=None):
    """The searches of one pair (frame 1 = "last", frame 2 = "current") on the CPU: the reference's own ORBmatcher /
    LineMatcher code (oracle/_ref) when built, else the oracle port.  affine None: identity pose, projection search
    only (sequence mode)."""
    from pl_vi_orbslam3_b200.capi import QUERY_DTYPE
    k = r1["keypoints"]
    a = np.asarray(affine if affine is not None else [1, 0, 0, 0, 1, 0], np.float32)
    u = (a[0] * k["x"] + a[1] * k["y"] + a[2]).astype(np.float32)
    v = (a[3] * k["x"] + a[4] * k["y"] + a[5]).astype(np.float32)
    t0 = time.perf_counter()
    if use_ref:
        # the reference drops projections outside the image bounds itself (src/ORBmatcher.cc:2007-2010)
        oracle.ref_search_frame(r2["keypoints"], r2["descriptors"], grid, (0.0, float(w), 0.0, float(h)), SCALE_FACTORS, k,
                                np.stack([u, v], 1), np.zeros(len(k), np.int32), r1["descriptors"], MATCH_TH, True)
    else:
        q = np.zeros(len(k), QUERY_DTYPE)
        q["u"], q["v"] = u, v
        q["radius"] = np.float32(MATCH_TH) * (np.float32(1.2) ** k["octave"].astype(np.float32))
        q["min_level"], q["max_level"], q["angle"] = k["octave"] - 1, k["octave"] + 1, k["angle"]
        q["flags"] = ((u < 0) | (u > w) | (v < 0) | (v > h)).astype(np.int32)
        oracle.search_frame(r2["keypoints"], r2["descriptors"], grid, q, r1["descriptors"], 100, True)
    t1 = time.perf_counter()
    if affine is not None:
        if use_ref:
            oracle.ref_search_init(k, r1["descriptors"], r2["keypoints"], r2["descriptors"], grid, np.stack([k["x"], k["y"]], 1),
                                   int(INIT_WINDOW), NNRATIO, True)
        else:
            from pl_vi_orbslam3_b200.matchers import ORBmatcher
            oracle.search_init(r2["keypoints"], r2["descriptors"], grid,
                               ORBmatcher.init_queries(k, np.stack([k["x"], k["y"]], 1), INIT_WINDOW), r1["descriptors"], 50, NNRATIO, True)
    t2 = time.perf_counter()
    if len(l1["descriptors"]) >= 2 and len(l2["descriptors"]) >= 2:
        if use_ref:
            oracle.ref_line_match(l1["descriptors"], l2["descriptors"], NNRATIO, "match")
        else:
            oracle.line_match(l1["descriptors"], l2["descriptors"], NNRATIO)
    t3 = time.perf_counter()
    if stages is not None:
        stages["search_by_projection"] = stages.get("search_by_projection", 0.0) + t1 - t0
        stages["search_for_initialization"] = stages.get("search_for_initialization", 0.0) + t2 - t1
        stages["line_match"] = stages.get("line_match", 0.0) + t3 - t2


def _cpu_job(job):
    """CPU arm over a contiguous shard of frames.  pairs: frames (2p, 2p + 1) are C3 pairs; else every frame is
    matched against the previous frame of the shard, like a sequential tracker."""
    frames, w, h, nfeat, pairs, affine = job
    import oracle
    from pl_vi_orbslam3_b200.matchers import frame_grid
    grid = frame_grid(0, w, 0, h)
    use_ref = oracle.ref_available()
    if use_ref:
        oracle.ref_set_monotone(False)   # plain malloc while timing
    prev = None
    for i, f in enumerate(frames):
        cur = _cpu_extract(oracle, use_ref, f, nfeat)
        if prev is not None and (not pairs or i % 2 == 1):
            _cpu_match(oracle, use_ref, grid, w, h, prev[0], prev[1], cur[0], cur[1], affine if pairs else None)
        prev = cur
    return len(frames)


def _cpu_stage_job(job):
    """Single-core ms per stage of the reference path on a few pairs and the reference's threading
    shape: points || lines on two threads per frame (src/Frame.cc:558-561), then the searches."""
    frames, w, h, nfeat, pairs, affine = job
    import oracle
    from pl_vi_orbslam3_b200.matchers import frame_grid
    grid = frame_grid(0, w, 0, h)
    use_ref = oracle.ref_available()
    if use_ref:
        oracle.ref_set_monotone(False)
    st, ex = {}, []
    for f in frames:
        t0 = time.perf_counter()
        r = oracle.ref_orb_extract(f, nfeat) if use_ref else oracle.orb_extract(f, nfeat)
        t1 = time.perf_counter()
        l = oracle.ref_line_extract(f) if use_ref else oracle.line_extract(f)
        t2 = time.perf_counter()
        st["orb_extract"] = st.get("orb_extract", 0.0) + t1 - t0
        st["line_extract"] = st.get("line_extract", 0.0) + t2 - t1
        ex.append((r, l))
    npair = 0
    for p in range(0, len(ex) - 1, 2 if pairs else 1):
        _cpu_match(oracle, use_ref, grid, w, h, ex[p][0], ex[p][1], ex[p + 1][0], ex[p + 1][1], affine if pairs else None, st)
        npair += 1
    n = len(frames)
    out = {k: 1e3 * v / (n if k.endswith("extract") else max(npair, 1)) for k, v in st.items()}
    match_per_frame = sum(v for k, v in out.items() if not k.endswith("extract")) * max(npair, 1) / n
    out["frame_ms_two_threads(points||lines, then searches)"] = max(out["orb_extract"], out["line_extract"]) + match_per_frame
    return {k: round(v, 3) for k, v in out.items()}


def cpu_kind():
    import oracle
    if oracle.ref_available():
        return "reference", ("oracle/_ref: the reference's own ORBextractor.cc / LSD/lsd.cpp / LSDDetector_custom.cpp / "
                             "binary_descriptor_custom.cpp / LineExtractor.cc compiled unmodified with g++ -O2; the OpenCV "
                             "primitives underneath (resize, GaussianBlur, FAST, pyrDown, Sobel) are the oracle's scalar "
                             "models, not OpenCV's SIMD code, so this understates a real OpenCV build; the searches "
                             "are the reference's own ORBmatcher::SearchByProjection(Frame, Frame) / SearchForInitialization and "
                             "LineMatcher::match (ORBmatcher.cc / LineMatcher.cpp compiled unmodified against stand-in Frame / MapPoint classes)")
    return "port", "oracle/ C++ port of the reference path (oracle/_ref not built)"


def cpu_arm(frames, cores, w, h, nfeat, pairs, affine, stages=False):
    """Times the CPU path over `frames` with `cores` worker processes (contiguous shards, pairs kept together).
    Returns (frames/s, seconds, per-stage dict or None).  Must run before CUDA is initialised in this process (fork)."""
    import multiprocessing as mp
    import oracle
    oracle.build()
    if oracle.ref_available():     # also map the libraries in this process (the workers are forks of it)
        try:
            oracle.ref_lib(), oracle.ref_orbmatcher_lib()
        except Exception:
            pass
    unit = 2 if pairs else 1
    nunits = len(frames) // unit
    bounds = np.linspace(0, nunits, min(cores, max(nunits, 1)) + 1).astype(int) * unit
    shards = [frames[a:b] for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    ctx = mp.get_context("fork")
    mk = lambda s: (s, w, h, nfeat, pairs, affine)
    with ctx.Pool(len(shards)) as pool:
        pool.map(_cpu_job, [mk(s[:unit]) for s in shards])  # warm-up: library load, page-in
        t0 = time.perf_counter()
        done = sum(pool.map(_cpu_job, [mk(s) for s in shards]))
        dt = time.perf_counter() - t0
        st = pool.apply(_cpu_stage_job, (mk(frames[:8]),)) if stages else None
    return done / dt, dt, st


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


def latest_profile(pattern):
    """The newest committed ncu-derived json under profiles/ matching rNN_<pattern> (highest round wins)."""
    c = sorted((ROOT / "profiles").glob(f"r[0-9][0-9]*_{pattern}"))   # r02_x < r02b_x < r03_x
    return c[-1] if c else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("PLVI_BENCH_CONFIG", "c23"), choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=int(os.environ.get("PLVI_BENCH_BATCH", 0)),
                    help="frames per GPU per step (0 = the config's own: 4096 / 256 / 4096 in total for c5)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames in the CPU baseline sample (0 = 32 per core: 10-30 s of CPU work)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--orb-only", action="store_true", help="diagnostic: time ORB extraction alone (not the bench metric)")
    ap.add_argument("--profile-out", default="")
    ap.add_argument("--no-overlap", action="store_true", help="run the line pipeline on the same stream as ORB")
    ap.add_argument("--line-priority", type=int, default=0)
    ap.add_argument("--pipe-mode", choices=("slice", "alternate"), default=os.environ.get("PLVI_BENCH_PIPE_MODE", "alternate"),
                    help="alternate: whole batches go to the pipelines in turn (consecutive batches in flight at "
                         "different phases); slice: every batch is split across the pipelines")
    ap.add_argument("--out-sets", type=int, default=2, help="alternating output buffer sets per pipeline")
    ap.add_argument("--pipes", type=int, default=int(os.environ.get("PLVI_BENCH_PIPES", 0)),
                    help="independent pipelines = consecutive steps in flight (0 = by batch size: 1 at 4096 frames per step, "
                         "8 at <= 512: region growing has a latency floor per batch, mid-size batches overlap)")
    ap.add_argument("--synth-workers", type=int, default=0, help="processes generating the synthetic frames (0 = cores / ranks)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    cores = os.cpu_count() or 1
    W, H, NFEAT, B0, pairs, scaling = CONFIGS[args.config]
    workload = WORKLOADS[args.config]

    from pl_vi_orbslam3_b200 import build as _build, synth
    from pl_vi_orbslam3_b200.frontend import shard_range
    if not _build.LIB.exists():        # the built library normally travels with the repo snapshot
        _build.build(force=True)
    affine = synth.warp_affine(W, H).astype(np.float32).reshape(6) if pairs else None
    gen = synth.pair_batch if pairs else synth.seq_batch
    gen_workers = args.synth_workers or max(1, cores // world)

    # ------------------------------------------------------------------ CPU arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        per_step = args.cpu_sample or min(32 * cores, 1024)
        frames = gen(per_step, W, H, base_seed=0, workers=cores)
        for _ in range(max(args.warmup, 0)):
            cpu_arm(frames[: 2 * max(cores, 2)], cores, W, H, NFEAT, pairs, affine)
        t = 0.0
        done = 0
        st = None
        for i in range(args.steps):
            fps, dt, s1 = cpu_arm(frames, cores, W, H, NFEAT, pairs, affine, stages=(i == 0))
            st = st or s1
            t += dt
            done += len(frames)
        v = done / t
        print(json.dumps({
            "impl": "reference", "metric": "frames/s", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "frames_per_step": per_step, "width": W, "height": H, "orb_features": NFEAT,
                       "frames": "C3 pairs (seed s, its warp), seeds 0.." if pairs else "sequence, seeds 0.."},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": cpu_kind()[0],
                             "sample": f"{per_step} synthetic frames per step, contiguous shards over {cores} worker processes; "
                                       + cpu_kind()[1],
                             "single_core_ms_per_stage": st},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ------------------------------------------------------------------ CUDA arm
    if args.config == "c5":
        lo, hi = shard_range(args.batch or B0, rank, world)
        B, seed0 = hi - lo, lo
    else:
        B = args.batch or B0
        seed0 = rank * (B // 2 if pairs else B)
    if B < 2:
        pairs, affine = False, None
    if args.pipes <= 0:   # frames in flight ~ 4096 frames of 752x480 (device memory: ~25 MB per frame of capacity at that size)
        # (2048 frames of 752x480 and more: one pipeline -- two 2048-frame batches in flight measured 20.9 k against 21.6 k)
        args.pipes = 1 if (B < 32 or B * W * H >= 2048 * 752 * 480) else max(1, min(8, int(4096 * 752 * 480 / (B * W * H))))
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        nsamp = args.cpu_sample or min(32 * cores, 1024)
        fps, dt, st = cpu_arm(gen(nsamp, W, H, base_seed=0, workers=cores), cores, W, H, NFEAT, pairs, affine, stages=True)
        cpu_base = {"value": fps, "unit": "frames/s", "cores": cores, "kind": cpu_kind()[0],
                    "sample": f"{nsamp} synthetic {W}x{H} frames (same generator and pairing), contiguous shards over {cores} worker "
                              f"processes, {dt:.1f} s wall; " + cpu_kind()[1],
                    "single_core_ms_per_stage": st}
    # every frame from its own seed when this rank's share of the host cores makes them within ~75 s, else the first
    # `distinct` frames from their seeds and cyclic shifts of them (stated in config.frames)
    frames, distinct = gen(B, W, H, base_seed=seed0, workers=gen_workers, budget_s=75.0, with_info=True)     # forks: before CUDA is initialised

    import torch
    import torch.distributed as dist
    from pl_vi_orbslam3_b200.frontend import PipelinedFrontEnd

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h_frames = torch.from_numpy(frames).pin_memory()
    fe = PipelinedFrontEnd(B, pipes=args.pipes, mode=args.pipe_mode, device=local_rank, w=W, h=H, nfeatures=NFEAT,
                           with_lines=not args.orb_only, with_match=not args.orb_only, overlap_lines=not args.no_overlap,
                           line_priority=args.line_priority, out_sets=args.out_sets, pairs=pairs, affine=affine,
                           match_th=MATCH_TH, nnratio=NNRATIO, init_window=INIT_WINDOW)
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

    # ---- end to end through the host-buffer C ABI (FrontEnd.step_host): the handles upload the pinned host frames on
    # their own copy streams (two input buffers: the upload of step i+1 overlaps the kernels of step i) and write every
    # result to pinned host memory; two host result sets, a set is reused once its step has completed (a consumer would
    # be reading it meanwhile).  Timed on the device: start event before the first upload, stop event after the last
    # result copy, with the host kept at most two steps ahead.
    fes = fe.fes
    nset = 2
    ios = [[f.alloc_host_io() for _ in range(nset)] for f in fes]

    sliced = args.pipe_mode == "slice" and len(fes) > 1
    h_parts = [h_frames[o:o + z] for o, z in zip(fe.offsets, fe.sizes)] if sliced else None

    def e2e_run(nsteps):
        done = [None] * nsteps
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(st)
        for f in fes:
            f.wait_event_all(start)
        ahead = nset if sliced else nset * len(fes)      # steps the host may run ahead of the results
        for i in range(nsteps):
            if i >= ahead:
                for ev in done[i - ahead]:
                    ev.synchronize()               # the host buffers of that step are free again
            if sliced:       # every pipeline takes its slice of the batch
                k = i % nset
                evs = []
                for j, f in enumerate(fes):
                    f.step_host(h_parts[j], ios[j][k])
                    evs += f.host_done_events()
                done[i] = evs
            else:            # whole batches go to the pipelines in turn
                f = fes[i % len(fes)]
                k = (i // len(fes)) % nset
                f.step_host(h_frames, ios[i % len(fes)][k])
                done[i] = f.host_done_events()
        for evs in done[-ahead:]:
            for ev in evs or ():
                ev.wait_on(st)
        stop.record(st)
        stop.synchronize()
        return start, stop

    e2e_run(max(2, 2 * len(fes)))   # every pipeline has used both of its staging buffers / result sets (allocations, graph captures)
    barrier()
    f0, f1 = e2e_run(args.steps)
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    # plvi_orb_extract_batch_async_from_line: the ORB handle reads the line handle's upload (one upload per batch); without
    # it each extractor call takes the image
    h2d = int(h_frames.numel()) * (2 if (fes[0].line is not None and not fes[0].share_upload) else 1)
    d2h = int(sum(v.numel() * v.element_size() for j in range(len(fes) if sliced else 1) for v in ios[j][0].values()))

    # ---- per-kernel profile of one extra step (events after every launch; not part of the timed numbers)
    fe.set_profile(True)
    with torch.cuda.stream(st):
        fe.step(d_frames, serialize=True)   # points and lines back to back: per-kernel times without co-scheduling
    st.synchronize()
    prof = fe.profile()
    fe.set_profile(False)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    nfr = torch.tensor([B], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nfr, op=dist.ReduceOp.SUM)
    ms, ms_e2e, total = float(t[0]), float(t[1]), float(nfr[0])
    outs = fe.outputs()
    counts = float(np.mean([o["counts"].float().mean().item() for o in outs]))
    if rank == 0:
        f0_ = fes[0]
        lw, lh = f0_.orb.level_sizes(W, H)
        if f0_.line is not None:
            ow, oh, sw, sh = f0_.line.octave_sizes(W, H)
        else:
            ow, oh, sw, sh = [W], [H], [0], [0]
        AB = algo_bytes(W, H, f0_.orb.capacity, lw, lh, ow, oh, sw, sh)
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        top = max(prof, key=prof.get) if prof else None
        roof = None
        if top:
            traffic, tsrc = None, None
            tj = latest_profile(f"{top}_traffic.json")   # dram bytes from the committed ncu --set full capture
            if tj is not None and args.config == "c23":
                traffic = json.loads(tj.read_text())["dram_bytes_per_frame"] * B
                tsrc = str(tj.relative_to(ROOT))
            ach = AB.get(top, 0) * B / (prof[top] * 1e-3) / 1e9
            roof = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": tsrc, "algorithmic_bytes_per_launch": AB.get(top, 0) * B,
                    "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback",
                    "kernel_ms_per_launch": prof[top], "frames_per_launch": B,
                    "note": GROW_NOTE if top.startswith("k_lsd_") and top not in ("k_lsd_pre", "k_lsd_scale_grad", "k_lsd_rowfilter") else ""}
        # integer / issue roofline of the compare- and bit-op-bound kernels: executed thread instructions per frame (from
        # the committed ncu captures, profiles/rNN_int_ops.json) x frames / the kernel's live time, against the measured
        # IADD3 issue peak (profiles/r01_instr_peaks.json, tools/int_peak.cu)
        int_roof = None
        ij, pj = latest_profile("int_ops.json"), latest_profile("instr_peaks.json")
        if ij is not None and pj is not None and args.config == "c23":
            # every committed capture, the newest one of a kernel wins (a later capture may hold only the kernels that changed)
            ops = {"kernels": {}}
            srcs = sorted((ROOT / "profiles").glob("r[0-9][0-9]*_int_ops.json"))
            for sj in srcs:
                ops["kernels"].update(json.loads(sj.read_text()).get("kernels", {}))
            ipk = json.loads(pj.read_text())["ginstr_per_s"]
            int_roof = {"peak_ginstr_per_s": ipk["iadd3"], "peak_source": str(pj.relative_to(ROOT)),
                        "ops_source": ", ".join(str(sj.relative_to(ROOT)) for sj in srcs), "kernels": {}}
            for k, o in ops.get("kernels", {}).items():
                ms_k = prof.get(k) or prof.get(k + "(+queries)")
                if ms_k:
                    g = o["thread_inst_per_frame"] * B / (ms_k * 1e-3) / 1e9
                    int_roof["kernels"][k] = {"ginstr_per_s": round(g, 1), "int_frac": round(g / ipk["iadd3"], 4),
                                              "thread_inst_per_frame": o["thread_inst_per_frame"]}
        total_prof = sum(prof.values()) or 1.0
        lat = None
        if B == 1:
            lat = {"ms_per_frame_device_resident": ms / args.steps, "ms_per_frame_host_buffers": ms_e2e / args.steps}
        line = {
            "metric": "frames/s", "value": total * args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload if not args.orb_only else "DIAGNOSTIC orb-only", "name": args.config,
                       "frames_per_step_per_gpu": B, "frames_per_step": int(total), "orb_features": NFEAT,
                       "frames": ((f"C3 pairs: frame_euroc(s) and its warp, s = {seed0}..{seed0 + distinct // 2 - 1} on rank 0" if pairs
                                   else f"frame_euroc(s), s = {seed0}..{seed0 + distinct - 1} on rank 0, consecutive frames matched") +
                                  (" (all distinct)" if distinct >= B else f" ({distinct} distinct frames, the rest cyclic shifts of them: "
                                                                             f"{gen_workers} generator processes on this rank)")),
                       "searches": ("SearchByProjection(Frame,Frame) th=15 under the warp + SearchForInitialization window 100 + LineMatcher::match 0.9"
                                    if pairs else "SearchByProjection(Frame,Frame) th=15, identity pose + LineMatcher::match 0.9"),
                       "width": W, "height": H, "l2_policy": "inputs larger than L2 (%.0f MB of frames per step)" % (B * W * H / 1e6) if B * W * H > 130e6
                       else "inputs smaller than L2: every step re-uploads them (e2e) / reads what the previous step left (value)",
                       "parallelism": f"frame-sharded x{world}, no collective", "mean_keypoints": counts,
                       "pipelines_per_gpu": args.pipes, "pipeline_mode": args.pipe_mode},
            "e2e": {"value": total * args.steps / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "api": "plvi_line_extract_batch_async + plvi_orb_extract_batch_async[_from_line] (host image in, host results out) + "
                           "plvi_search_by_projection / plvi_line_match on the handles' device results, match tables to host"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "int_roofline": int_roof,
            "latency": lat,
            "cpu_baseline": cpu_base,
            # every kernel against the HBM roofline: algorithmic GB/s = algo_bytes x frames / its time in the serialised
            # profile pass, as a fraction of the measured copy peak (streaming kernels should be judged on this one;
            # region growing / k_octree / k_search are latency or issue bound)
            "kernel_hbm": {k: {"gbs": round(AB[k] * B / (v * 1e-3) / 1e9, 1),
                               "frac": round(AB[k] * B / (v * 1e-3) / 1e9 / peak, 4)}
                           for k, v in sorted(prof.items(), key=lambda kv: -kv[1]) if k in AB and v > 0},
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
