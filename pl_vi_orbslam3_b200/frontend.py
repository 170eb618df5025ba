"""Batched front-end step on one GPU: ORB + LSD/LBD extraction of a batch of frames and
frame-to-frame Hamming matching of consecutive frames, all on one CUDA stream with every
buffer resident in HBM.  This is the unit bench.py times and the multi-GPU runs shard:
frames are independent, so each rank owns a contiguous range of frames and there is no
collective.

torch is used for device memory and streams only; every kernel is in libplvi_cuda.so.
"""
import os

import numpy as np

from .capi import QUERY_DTYPE, check, lib, ptr
from .lineextractor import Lineextractor
from .matchers import LineMatcher, ORBmatcher, frame_grid
from .orbextractor import ORBextractor


def shard_range(n_items, rank, world):
    """Contiguous frame range [lo, hi) of `rank` (4096/G frames per GPU for config 5)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _Event:
    """A torch event or a raw cudaEvent_t of a handle behind one interface."""

    def __init__(self, torch_event=None, raw=None):
        self.ev, self.raw = torch_event, raw

    def synchronize(self):
        if self.ev is not None:
            self.ev.synchronize()
        elif self.raw:
            check(lib().plvi_event_synchronize(ptr(self.raw)))

    def query(self):
        if self.ev is not None:
            return self.ev.query()
        self.synchronize()
        return True

    def wait_on(self, stream):
        if self.ev is not None:
            stream.wait_event(self.ev)
        elif self.raw:
            check(lib().plvi_stream_wait_event(ptr(stream.cuda_stream), ptr(self.raw)))


class FrontEnd:
    def __init__(self, batch, w=752, h=480, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7,
                 lsd_nfeatures=200, lsd_scale=0.8, line_levels=2, line_scale=2.0, device=0, stream=None,
                 with_lines=True, with_match=True, match_th=15.0, nnratio=0.9, overlap_lines=True, line_priority=0,
                 out_sets=1, pairs=False, affine=None, init_window=100.0, band_run_max=None):
        import torch
        self.torch = torch
        self.B, self.w, self.h = batch, w, h
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        # PLVI_ORB_PRIORITY < 0: the stream of the ORB kernels and the searches outranks the line stream, so their blocks
        # take the block slots that the (long, latency-bound) region-growing kernels free
        self.stream = stream if stream is not None else torch.cuda.Stream(
            device=self.device, priority=int(os.environ.get("PLVI_ORB_PRIORITY", "0")))
        sp = self.stream.cuda_stream
        # points and lines are independent (the reference runs them on two threads, src/Frame.cc:558-561):
        # the line pipeline gets its own stream, forked from / joined into self.stream every step, so the
        # latency-bound LSD region growing overlaps the throughput-bound ORB kernels
        if line_priority == 0:
            line_priority = int(os.environ.get("PLVI_LINE_PRIORITY", "0"))
        self.line_stream = torch.cuda.Stream(device=self.device, priority=line_priority) if overlap_lines else self.stream
        self._ev_fork = torch.cuda.Event()
        self._ev_join = torch.cuda.Event()
        self.scale_factor = float(scale_factor)
        self.match_th = float(match_th)
        self.orb = ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th, max_width=w, max_height=h,
                                max_batch=batch, device=device, stream=sp)
        # out_sets > 1: the results of consecutive steps go to alternating output buffers, so the
        # device-to-host copy of step i can still run while step i+1 computes
        self.out_sets = max(1, int(out_sets))
        # ORB kernels start when the line pipeline reaches region growing ("1", default: the whole ORB sequence waits;
        # "2": the pyramid runs at once -- measured slower, its short chained kernels starve beside k_lsd_pre; "0": no coupling)
        self.skew = os.environ.get("PLVI_SKEW", "1")
        self.skew = int(self.skew) if self.skew in ("0", "1", "2", "3") else 1
        self.pyr_first = os.environ.get("PLVI_PYR_FIRST", "0") != "0"   # step(): ORB pyramid ahead of the line pipeline
        self.share_upload = os.environ.get("PLVI_SHARE_UPLOAD", "1") != "0"   # step_host: the frames are uploaded once for both extractors
        self._set = 0
        self.orb_outs = [self.orb.alloc_device_outputs(batch, self.device) for _ in range(self.out_sets)]
        self.orb_out = self.orb_outs[0]
        self.line = None
        self.om = self.lm = None
        if with_lines:
            self.line = Lineextractor(lsd_nfeatures, 0, lsd_scale, line_levels, line_scale, 0, max_width=w,
                                      max_height=h, max_batch=batch, device=device,
                                      stream=self.line_stream.cuda_stream,
                                      band_run_max=-1 if band_run_max is None else int(band_run_max))
            self.line_outs = [self.line.alloc_device_outputs(batch, self.device) for _ in range(self.out_sets)]
            self.line_out = self.line_outs[0]
        # pairs=True: the batch holds C3 pairs (frame 2p, its warp 2p + 1) and every pair is matched in the
        # two point modes of C3 -- SearchByProjection(Frame, Frame) semantics with the warp as the "pose" and
        # SearchForInitialization semantics -- plus LineMatcher::match; else consecutive frames under the identity pose
        self.pairs = bool(pairs)
        self.affine = np.asarray(affine if affine is not None else [1, 0, 0, 0, 1, 0], np.float32).reshape(6)
        self.bounds = np.asarray([0, w, 0, h], np.float32)
        self.init_window = float(init_window)
        if with_match and batch > 1 and self.pairs:
            cap = self.orb.capacity
            P, S = batch // 2, 2 * cap
            self.om = ORBmatcher(nnratio, True, max_pairs=P, max_train=S, max_query=S, device=device, stream=sp)
            self.grid = frame_grid(0, w, 0, h)
            self.q_proj = torch.zeros((P, S, 7), dtype=torch.float32, device=self.device)
            self.q_init = torch.zeros((P, S, 7), dtype=torch.float32, device=self.device)
            self.qcount = torch.zeros(P, dtype=torch.int32, device=self.device)
            self.tcount = torch.zeros(P, dtype=torch.int32, device=self.device)
            self.scratch_mq = torch.zeros((P, S), dtype=torch.int32, device=self.device)
            self.match_sets = [tuple(torch.zeros((P, S), dtype=torch.int32, device=self.device) for _ in range(2)) +
                               tuple(torch.zeros(P, dtype=torch.int32, device=self.device) for _ in range(2))
                               for _ in range(self.out_sets)]
            self.match_train, self.init_m12, self.nmatches, self.init_nm = self.match_sets[0]
            if with_lines:
                lc = self.line.capacity
                self.lm = LineMatcher(max_pairs=P, max_train=2 * lc, max_query=2 * lc, device=device, stream=sp)
                self.lq = torch.zeros(P, dtype=torch.int32, device=self.device)
                self.lt = torch.zeros(P, dtype=torch.int32, device=self.device)
                self.lmatch_sets = [(torch.zeros((P, 2 * lc), dtype=torch.int32, device=self.device),
                                     torch.zeros(P, dtype=torch.int32, device=self.device)) for _ in range(self.out_sets)]
                self.line_m12, self.line_nm = self.lmatch_sets[0]
        elif with_match and batch > 1:
            cap = self.orb.capacity
            self.om = ORBmatcher(nnratio, True, max_pairs=batch, max_train=cap, max_query=cap, device=device, stream=sp)
            self.grid = frame_grid(0, w, 0, h)
            P = batch - 1
            self.queries = torch.zeros((batch, cap, 7), dtype=torch.float32, device=self.device)
            self.match_sets = [(torch.zeros((P, cap), dtype=torch.int32, device=self.device),
                                torch.zeros((P, cap), dtype=torch.int32, device=self.device),
                                torch.zeros(P, dtype=torch.int32, device=self.device)) for _ in range(self.out_sets)]
            self.match_train, self.match_query, self.nmatches = self.match_sets[0]
            if with_lines:
                lc = self.line.capacity
                self.lm = LineMatcher(max_pairs=batch, max_train=lc, max_query=lc, device=device, stream=sp)
                self.lmatch_sets = [(torch.zeros((P, lc), dtype=torch.int32, device=self.device),
                                     torch.zeros(P, dtype=torch.int32, device=self.device)) for _ in range(self.out_sets)]
                self.line_m12, self.line_nm = self.lmatch_sets[0]
        self.launches = 0
        self._profiling = False
        self._mev = None

    def close(self):
        for o in (self.orb, self.line, self.om, self.lm):
            if o is not None:
                o.close()

    def step(self, d_frames, serialize=False):
        """d_frames: uint8 CUDA tensor [n<=B, h, w].  Enqueues everything on self.stream (the line
        pipeline on self.line_stream, joined before matching).  serialize=True makes the ORB kernels
        wait for the line pipeline (clean per-kernel timings for the profile pass)."""
        n = d_frames.shape[0]
        nl = 0
        self._next_set()
        forked = self.line is not None and self.line_stream is not self.stream
        if forked and self.pyr_first and not serialize:
            # the ORB pyramid (seven short chained kernels) before anything else is in flight: beside k_lsd_pre or the
            # region growing its blocks wait for slots and the chain takes 5x as long
            check(lib().plvi_orb_pyramid_device(self.orb._h, ptr(d_frames), n, d_frames.shape[2], d_frames.shape[1],
                                                d_frames.stride(1), d_frames.stride(0)))
        if forked:
            self._ev_fork.record(self.stream)
            self.line_stream.wait_event(self._ev_fork)
        if self.line is not None:
            kl, ldesc, leq, lcounts = self.line.extract_batch_device(d_frames, out=self.line_out)
            nl += self.line.last_launches
        if forked and serialize:
            self._ev_join.record(self.line_stream)
            self.stream.wait_event(self._ev_join)
        elif forked and self.skew:
            # hold the (issue-bound) ORB kernels back until the line pipeline has reached region growing
            self._orb_wait_stage()
        kps, desc, counts, mono = self.orb.extract_batch_device(d_frames, out=self.orb_out)
        nl += self.orb.last_launches
        # the point searches only need the ORB results: they run while the line pipeline (the longer one) is still busy;
        # the line matches follow the join
        join = None
        if forked and not serialize:
            self._ev_join.record(self.line_stream)
            join = self._ev_join
        nl += self._match(kps, desc, counts, ldesc if self.line is not None else None,
                          lcounts if self.line is not None else None, n, join)
        if join is not None and (self.lm is None or n < 2):
            self.stream.wait_event(join)
        self.launches = nl
        return nl


    def _orb_wait_stage(self):
        fn = lib().plvi_orb_wait_event_after_pyramid if self.skew == 2 else lib().plvi_orb_wait_event
        check(fn(self.orb._h, lib().plvi_line_stage_event(self.line._h)))
        if self.skew == 3:   # ... and until the blocks of the region-growing kernel are resident
            import ctypes as C
            target = C.c_int(0)
            ctr = lib().plvi_line_stage_counter(self.line._h, C.byref(target))
            check(lib().plvi_orb_wait_counter(self.orb._h, C.c_void_p(ctr), target.value))

    def _next_set(self):
        if self.out_sets > 1:   # next output buffer set
            self._set = (self._set + 1) % self.out_sets
            self.orb_out = self.orb_outs[self._set]
            if self.line is not None:
                self.line_out = self.line_outs[self._set]
            if self.om is not None and self.pairs:
                self.match_train, self.init_m12, self.nmatches, self.init_nm = self.match_sets[self._set]
            elif self.om is not None:
                self.match_train, self.match_query, self.nmatches = self.match_sets[self._set]
            if self.lm is not None:
                self.line_m12, self.line_nm = self.lmatch_sets[self._set]

    def _match(self, kps, desc, counts, ldesc, lcounts, n, join=None):
        """The searches of one step on device-resident extraction results (tensors or raw device addresses), enqueued on
        self.stream; `join`: event of the line pipeline the line matches wait for.  Returns the number of kernel launches."""
        nl = 0
        if self.om is None or n < 2:
            return 0
        if self._profiling:
            self._mev = [self.torch.cuda.Event(enable_timing=True) for _ in range(4)]
            self._mev[0].record(self.stream)
        cap = self.orb.capacity
        a = (lambda t: t.data_ptr()) if hasattr(kps, "data_ptr") else (lambda t: t)
        if self.pairs:
            P, S = n // 2, 2 * cap
            check(lib().plvi_pair_queries(self.om._h, ptr(kps), ptr(counts), P, cap, S, self.match_th, self.scale_factor,
                                          ptr(self.affine), ptr(self.bounds), self.init_window, ptr(self.q_proj), ptr(self.q_init),
                                          ptr(self.qcount), ptr(self.tcount)))
            # frame 2p + 1 is the searched ("current") frame, frame 2p the query ("last") frame of pair p
            k1, d1 = a(kps) + cap * 28, a(desc) + cap * 32
            check(lib().plvi_search_by_projection(
                self.om._h, 0, P, ptr(k1), ptr(d1), None, ptr(self.tcount), S, ptr(self.grid), ptr(self.q_proj), ptr(desc),
                ptr(self.qcount), S, ORBmatcher.TH_HIGH, self.om.mfNNratio, 1, ptr(self.match_train), ptr(self.scratch_mq),
                ptr(self.nmatches), 1))
            check(lib().plvi_search_by_projection(
                self.om._h, 2, P, ptr(k1), ptr(d1), None, ptr(self.tcount), S, ptr(self.grid), ptr(self.q_init), ptr(desc),
                ptr(self.qcount), S, ORBmatcher.TH_LOW, self.om.mfNNratio, 1, ptr(self.scratch_mq), ptr(self.init_m12),
                ptr(self.init_nm), 1))
            nl += 3
            if self._profiling:
                self._mev[1].record(self.stream)
            if self.lm is not None:
                if join is not None:
                    self.stream.wait_event(join)
                lc = self.line.capacity
                sp = self.stream.cuda_stream
                check(lib().plvi_gather_i32(ptr(sp), ptr(lcounts), P, 0, 2, ptr(self.lq)))
                check(lib().plvi_gather_i32(ptr(sp), ptr(lcounts), P, 1, 2, ptr(self.lt)))
                check(lib().plvi_line_match(self.lm._h, P, ptr(ldesc), ptr(self.lq), 2 * lc, ptr(a(ldesc) + lc * 32), ptr(self.lt),
                                            2 * lc, 0.9, 1, ptr(self.line_m12), ptr(self.line_nm), 1))
                nl += 3
                if self._profiling:
                    self._mev[2].record(self.stream)
            return nl
        P = n - 1
        # frame p's keypoints are the "last frame" points searched in frame p+1
        check(lib().plvi_queries_from_keypoints(self.om._h, ptr(kps), ptr(counts), P, cap, self.match_th,
                                                self.scale_factor, ptr(self.queries)))
        check(lib().plvi_search_by_projection(
            self.om._h, 0, P, ptr(a(kps) + cap * 28), ptr(a(desc) + cap * 32), None, ptr(a(counts) + 4), cap, ptr(self.grid),
            ptr(self.queries), ptr(desc), ptr(counts), cap, ORBmatcher.TH_HIGH, self.om.mfNNratio, 1,
            ptr(self.match_train), ptr(self.match_query), ptr(self.nmatches), 1))
        nl += 2
        if self._profiling:
            self._mev[1].record(self.stream)
        if self.lm is not None:
            if join is not None:
                self.stream.wait_event(join)
            lc = self.line.capacity
            check(lib().plvi_line_match(self.lm._h, P, ptr(ldesc), ptr(lcounts), lc, ptr(a(ldesc) + lc * 32), ptr(a(lcounts) + 4),
                                        lc, 0.9, 1, ptr(self.line_m12), ptr(self.line_nm), 1))
            nl += 1
            if self._profiling:
                self._mev[2].record(self.stream)
        return nl

    # ---- the same step through the reference-facing C ABI with HOST buffers (what the C++ shim calls) -------------------
    def alloc_host_io(self):
        """Pinned host result buffers of one step (the caller-owned vectors / cv::Mat of the reference's operator())."""
        t = self.torch
        cap, B = self.orb.capacity, self.B
        io = {"kps": t.zeros((B, cap, 7), dtype=t.float32).pin_memory(), "desc": t.zeros((B, cap, 32), dtype=t.uint8).pin_memory(),
              "counts": t.zeros(B, dtype=t.int32).pin_memory(), "mono": t.zeros(B, dtype=t.int32).pin_memory()}
        if self.line is not None:
            lc = self.line.capacity
            io.update(keylines=t.zeros((B, lc, 17), dtype=t.float32).pin_memory(), line_desc=t.zeros((B, lc, 32), dtype=t.uint8).pin_memory(),
                      line_eq=t.zeros((B, lc, 3), dtype=t.float64).pin_memory(), line_counts=t.zeros(B, dtype=t.int32).pin_memory())
        for k, v in self.match_outputs().items():
            io[k] = t.zeros(v.shape, dtype=v.dtype).pin_memory()
        return io

    def match_outputs(self):
        out = {}
        if self.om is not None and self.pairs:
            out.update(match_train=self.match_train, nmatches=self.nmatches, init_matches=self.init_m12, init_nmatches=self.init_nm)
        elif self.om is not None:
            out.update(match_train=self.match_train, nmatches=self.nmatches)
        if self.lm is not None:
            out.update(line_matches=self.line_m12, line_nmatches=self.line_nm)
        return out

    def step_host(self, h_frames, io):
        """h_frames: pinned uint8 host tensor [n, h, w]; io: alloc_host_io().  plvi_orb_extract_batch_async +
        plvi_line_extract_batch_async (host image in, host results out; uploads on the handles' copy streams overlap
        the previous call's kernels), then the searches on the handles' device copies of the results and the match
        tables back to the host.  Nothing blocks: sync_host() waits for everything issued."""
        n, h, w = h_frames.shape
        nl = 0
        self._next_set()
        if self.line is not None:
            if self.line_stream is not self.stream and getattr(self, "_ev_match", None) is not None:
                self.line_stream.wait_event(self._ev_match)   # the searches of the previous step read the line results
            check(lib().plvi_line_extract_batch_async(self.line._h, ptr(h_frames), n, w, h, h_frames.stride(1), h_frames.stride(0),
                                                      ptr(io["keylines"]), ptr(io["line_desc"]), ptr(io["line_eq"]),
                                                      ptr(io["line_counts"])))
            nl += self.line.last_launches
            if self.line_stream is not self.stream:
                self._ev_join.record(self.line_stream)
                if self.skew:   # as in step(): the ORB kernels start when the line pipeline has reached region growing
                    self._orb_wait_stage()
        if self.line is not None and self.share_upload:
            # one upload per batch: the ORB handle reads the device copy the line handle just made
            check(lib().plvi_orb_extract_batch_async_from_line(self.orb._h, self.line._h, 0, 0, ptr(io["kps"]), ptr(io["desc"]),
                                                               ptr(io["counts"]), ptr(io["mono"])))
        else:
            check(lib().plvi_orb_extract_batch_async(self.orb._h, ptr(h_frames), n, w, h, h_frames.stride(1), h_frames.stride(0), 0, 0,
                                                     ptr(io["kps"]), ptr(io["desc"]), ptr(io["counts"]), ptr(io["mono"])))
        nl += self.orb.last_launches
        if self.om is not None and n > 1:
            import ctypes as C
            dk, dd, dc, dm = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
            check(lib().plvi_orb_device_results(self.orb._h, C.byref(dk), C.byref(dd), C.byref(dc), C.byref(dm)))
            ld = lcn = None
            if self.line is not None:
                lk, ldp, le, lcp = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
                check(lib().plvi_line_device_results(self.line._h, C.byref(lk), C.byref(ldp), C.byref(le), C.byref(lcp)))
                ld, lcn = ldp.value, lcp.value
            join = self._ev_join if (self.line is not None and self.line_stream is not self.stream) else None
            nl += self._match(dk.value, dd.value, dc.value, ld, lcn, n, join)
            if join is not None and self.lm is None:
                self.stream.wait_event(join)
            with self.torch.cuda.stream(self.stream):
                for k, v in self.match_outputs().items():
                    io[k].copy_(v, non_blocking=True)
            if self.line is not None and self.line_stream is not self.stream:
                self._ev_match = self.torch.cuda.Event()
                self._ev_match.record(self.stream)
        self.launches = nl
        return nl

    def sync_host(self):
        if self.line is not None:
            self.line.sync()
        self.orb.sync()
        self.stream.synchronize()

    def wait_event_all(self, ev):
        """Every stream of this front end waits for `ev` before the work issued next."""
        self.stream.wait_event(ev)
        if self.line is not None and self.line_stream is not self.stream:
            self.line_stream.wait_event(ev)

    def host_done_events(self):
        """Events (objects with .synchronize(), .query() and .wait_on(stream)) that complete when everything the last
        step_host() issued -- kernels, the handles' result copies on their device-to-host streams and the match tables --
        has reached the host buffers."""
        evs = [_Event(torch_event=self.torch.cuda.Event())]
        evs[0].ev.record(self.stream)
        evs.append(_Event(raw=lib().plvi_orb_results_event(self.orb._h)))
        if self.line is not None:
            evs.append(_Event(raw=lib().plvi_line_results_event(self.line._h)))
        return evs

    def outputs(self):
        out = {"kps": self.orb_out[0], "desc": self.orb_out[1], "counts": self.orb_out[2], "mono": self.orb_out[3]}
        if self.line is not None:
            out.update(keylines=self.line_out[0], line_desc=self.line_out[1], line_eq=self.line_out[2],
                       line_counts=self.line_out[3])
        out.update(self.match_outputs())
        return out

    def set_profile(self, on=True):
        self._profiling = bool(on)
        check(lib().plvi_orb_set_profile(self.orb._h, int(on)))
        if self.line is not None:
            check(lib().plvi_line_set_profile(self.line._h, int(on)))

    def profile(self):
        """{kernel: ms} of the last step (extractor kernels; needs set_profile(True) before the step)."""
        txt = lib().plvi_orb_profile(self.orb._h).decode()
        if self.line is not None:
            txt += lib().plvi_line_profile(self.line._h).decode()
        prof = {}
        if self.om is not None and getattr(self, "_mev", None):
            self.stream.synchronize()
            prof["k_search(+queries)"] = self._mev[0].elapsed_time(self._mev[1])
            if self.lm is not None:
                prof["k_line_match"] = self._mev[1].elapsed_time(self._mev[2])
        for item in txt.split(";"):
            if "=" in item:
                k, v = item.split("=")
                prof[k] = prof.get(k, 0.0) + float(v)
        return prof


class PipelinedFrontEnd:
    """`pipes` independent FrontEnd pipelines (own handles, streams and scratch).

    mode="slice": every step is split into `pipes` contiguous slices of the batch that run side
    by side (frame-to-frame matching stays inside a slice).
    mode="alternate": whole batches go to the pipelines in turn (step i -> pipeline i % pipes)
    with no join in between, so consecutive batches are in flight at different phases: the
    latency-bound LSD region growing of one batch (two long serial chains per frame, the GPU's
    issue slots mostly idle) overlaps the throughput-bound kernels of the next.  The outputs of
    step i stay valid until step i + pipes; `done_event()` is the completion of the last step."""

    def __init__(self, batch, pipes=2, device=0, mode="slice", **kw):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.Stream(device=self.device)      # master stream: fork / join point
        self.mode = mode
        self.pipes = pipes
        if mode == "alternate":
            self.sizes = [batch] * pipes
            self.offsets = [0] * pipes
        else:
            self.sizes = [shard_range(batch, i, pipes)[1] - shard_range(batch, i, pipes)[0] for i in range(pipes)]
            self.offsets = [shard_range(batch, i, pipes)[0] for i in range(pipes)]
        if pipes > 1 and "band_run_max" not in kw:
            kw["band_run_max"] = 32      # several batches in flight: the speculation / commit schedule overlaps across them
        self.fes = [FrontEnd(sz, device=device, **kw) for sz in self.sizes]
        self._fork = torch.cuda.Event()
        self._joins = [torch.cuda.Event() for _ in self.fes]
        self._k = 0          # steps issued (alternate mode)
        self._last = None    # pipeline of the last step
        self.B = batch

    def close(self):
        for fe in self.fes:
            fe.close()

    def step(self, d_frames, serialize=False, wait=()):
        """Enqueue one step.  slice mode: forked from / joined into self.stream.  alternate mode:
        runs on the next pipeline's own stream after the events in `wait` and whatever self.stream
        held when the step was issued; nothing is joined back (see done_event / drain)."""
        nl = 0
        if self.mode != "alternate":
            for w in wait:
                self.stream.wait_event(w)
        self._fork.record(self.stream)
        if self.mode == "alternate":
            idx = self._k % self.pipes
            self._k += 1
            fe, ev = self.fes[idx], self._joins[idx]
            fe.stream.wait_event(self._fork)
            for w in wait:
                fe.stream.wait_event(w)
            with self.torch.cuda.stream(fe.stream):
                nl = fe.step(d_frames, serialize=serialize)
            ev.record(fe.stream)
            self._last = idx
            if serialize:
                self.stream.wait_event(ev)
            return nl
        for fe, off, sz, ev in zip(self.fes, self.offsets, self.sizes, self._joins):
            fe.stream.wait_event(self._fork)
            with self.torch.cuda.stream(fe.stream):
                nl += fe.step(d_frames[off:off + sz], serialize=serialize)
            ev.record(fe.stream)
            if serialize:
                self.stream.wait_event(ev)
                self._fork.record(self.stream)
        for ev in self._joins:
            self.stream.wait_event(ev)
        return nl

    @property
    def alive_steps(self):
        """Number of consecutive steps whose outputs exist at the same time."""
        return self.fes[0].out_sets * (self.pipes if self.mode == "alternate" else 1)

    def done_event(self):
        """Completion event of the last issued step (alternate mode); None in slice mode, where
        the step is already joined into self.stream."""
        return self._joins[self._last] if self.mode == "alternate" and self._last is not None else None

    def drain(self):
        """Make self.stream wait for every step issued so far."""
        if self.mode == "alternate":
            for i, ev in enumerate(self._joins):
                if self._k > i:
                    self.stream.wait_event(ev)

    def outputs(self):
        """slice mode: list of per-slice output dicts.  alternate mode: [outputs of the last issued step]."""
        if self.mode == "alternate":
            return [self.fes[self._last or 0].outputs()]
        return [fe.outputs() for fe in self.fes]

    def set_profile(self, on=True):
        for fe in self.fes:
            fe.set_profile(on)

    def profile(self):
        """{kernel: ms} of the last step."""
        prof = {}
        for fe in ([self.fes[self._last or 0]] if self.mode == "alternate" else self.fes):
            for k, v in fe.profile().items():
                prof[k] = prof.get(k, 0.0) + v
        return prof
