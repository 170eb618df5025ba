"""Host-side mirror of the reference's ORBextractor (include/ORBextractor.h:45-110).

Same constructor arguments, same call semantics (returns monoIndex; keypoints laid out
as cv::KeyPoint PODs; descriptors N x 32 u8), plus the batch / device-resident entry
points used for configs 4-5 of BASELINE.json.  All work happens in libplvi_cuda.so.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import KEYPOINT_DTYPE, check, lib, ptr


class ORBextractor:
    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST,
                 max_width=752, max_height=480, max_batch=1, device=0, stream=None):
        self._h = C.c_void_p()
        check(lib().plvi_orb_create(C.byref(self._h), int(nfeatures), float(scaleFactor), int(nlevels),
                                    int(iniThFAST), int(minThFAST), int(max_width), int(max_height),
                                    int(max_batch), int(device), ptr(stream) if stream else None))
        self.nfeatures, self.nlevels = int(nfeatures), int(nlevels)
        self.max_batch = int(max_batch)
        self.capacity = check(lib().plvi_orb_capacity(self._h))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().plvi_orb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (include/ORBextractor.h:62-82)
    def GetLevels(self):
        return lib().plvi_orb_levels(self._h)

    def GetScaleFactor(self):
        return lib().plvi_orb_scale_factor(self._h)

    def _tables(self):
        t = [np.empty(self.nlevels, np.float32) for _ in range(4)]
        check(lib().plvi_orb_scale_factors(self._h, *[ptr(a) for a in t]))
        return t

    def GetScaleFactors(self):
        return self._tables()[0]

    def GetInverseScaleFactors(self):
        return self._tables()[1]

    def GetScaleSigmaSquares(self):
        return self._tables()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._tables()[3]

    def features_per_level(self):
        q = np.empty(self.nlevels, np.int32)
        check(lib().plvi_orb_features_per_level(self._h, ptr(q)))
        return q

    def level_sizes(self, w, h):
        lw, lh = np.empty(self.nlevels, np.int32), np.empty(self.nlevels, np.int32)
        check(lib().plvi_orb_level_sizes(self._h, w, h, ptr(lw), ptr(lh)))
        return lw, lh

    @property
    def stream(self):
        return lib().plvi_orb_stream(self._h)

    @property
    def last_launches(self):
        return lib().plvi_orb_last_launches(self._h)

    def stereo_matches(self, right, left_out, right_out, mb, mbf):
        """Frame::ComputeStereoMatches (src/Frame.cc:1228-1406) for the last device batch of this (left) extractor and
        `right`: left_out / right_out = the (kps, desc, counts, mono) CUDA tensors extract_batch_device returned.
        Returns CUDA tensors (mvuRight [n, cap] f32, mvDepth [n, cap] f32, nstereo [n] i32); -1 = no stereo match."""
        import torch
        kl, dl, cl, _ = left_out
        kr, dr, cr, _ = right_out
        n, cap = cl.shape[0], kl.shape[1]
        ur = torch.empty((n, cap), dtype=torch.float32, device=kl.device)
        dp = torch.empty((n, cap), dtype=torch.float32, device=kl.device)
        ns = torch.empty(n, dtype=torch.int32, device=kl.device)
        check(lib().plvi_orb_stereo_matches(self._h, right._h, n, ptr(kl), ptr(dl), ptr(cl), ptr(kr), ptr(dr), ptr(cr), cap,
                                            float(mb), float(mbf), ptr(ur), ptr(dp), ptr(ns)))
        return ur, dp, ns

    def graph_stats(self):
        """(captured CUDA graphs, graph replays) of this handle's per-batch launch sequence."""
        import ctypes
        c = ctypes.c_int(0)
        r = lib().plvi_orb_graph_stats(self._h, ctypes.byref(c))
        return c.value, r

    # ---- operator() (include/ORBextractor.h:58-60)
    def __call__(self, image, mask=None, vLappingArea=(0, 0)):
        """Returns (monoIndex, keypoints[KEYPOINT_DTYPE], descriptors[N,32] u8).
        An empty image returns (-1, [], []) like the reference."""
        image = np.asarray(image)
        if image.size == 0:
            return -1, np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise ValueError("ORBextractor expects a CV_8UC1 image")
        kps, desc, counts, mono = self.extract_batch(image[None], vLappingArea)
        n = int(counts[0])
        return int(mono[0]), kps[0, :n].copy(), desc[0, :n].copy()

    # ---- batch of equally sized host frames [n,h,w] u8
    def extract_batch(self, frames, vLappingArea=(0, 0), out=None, sync=True):
        frames = np.asarray(frames)
        assert frames.dtype == np.uint8 and frames.ndim == 3
        if not frames.flags.c_contiguous:
            frames = np.ascontiguousarray(frames)
        n, h, w = frames.shape
        if out is None:
            out = self.alloc_host_outputs(n)
        kps, desc, counts, mono = out
        fn = lib().plvi_orb_extract_batch if sync else lib().plvi_orb_extract_batch_async
        check(fn(self._h, ptr(frames), n, w, h, frames.strides[1], frames.strides[0],
                 int(vLappingArea[0]), int(vLappingArea[1]), ptr(kps), ptr(desc), ptr(counts), ptr(mono)))
        return kps, desc, counts, mono

    def alloc_host_outputs(self, n):
        return (np.zeros((n, self.capacity), KEYPOINT_DTYPE), np.zeros((n, self.capacity, 32), np.uint8),
                np.zeros(n, np.int32), np.zeros(n, np.int32))

    def sync(self):
        check(lib().plvi_orb_sync(self._h))

    # ---- device-resident batch: torch uint8 CUDA tensor [n,h,w]; outputs are torch tensors
    def extract_batch_device(self, frames, vLappingArea=(0, 0), out=None):
        import torch
        assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 3
        n, h, w = frames.shape
        assert frames.stride(2) == 1
        if out is None:
            out = self.alloc_device_outputs(n, frames.device)
        kps, desc, counts, mono = out
        check(lib().plvi_orb_extract_batch_device(
            self._h, ptr(frames), n, w, h, frames.stride(1), frames.stride(0),
            int(vLappingArea[0]), int(vLappingArea[1]), ptr(kps), ptr(desc), ptr(counts), ptr(mono)))
        return kps, desc, counts, mono

    def alloc_device_outputs(self, n, device):
        import torch
        return (torch.zeros((n, self.capacity, 7), dtype=torch.float32, device=device),
                torch.zeros((n, self.capacity, 32), dtype=torch.uint8, device=device),
                torch.zeros(n, dtype=torch.int32, device=device),
                torch.zeros(n, dtype=torch.int32, device=device))

    # ---- debug read-back of the last batch (parity tests)
    def read_level(self, frame, level, w, h, blurred=False):
        lw, lh = self.level_sizes(w, h)
        out = np.empty((int(lh[level]), int(lw[level])), np.uint8)
        check(lib().plvi_orb_read_level(self._h, frame, level, int(blurred), ptr(out)))
        return out

    def read_candidates(self, frame, level):
        cap = 1 << 18
        buf = np.empty(cap, np.uint32)
        cnt = C.c_int(0)
        check(lib().plvi_orb_read_candidates(self._h, frame, level, ptr(buf), cap, C.byref(cnt)))
        v = buf[:cnt.value]
        return np.stack([v & 0xFFF, (v >> 12) & 0xFFF, v >> 24], axis=1).astype(np.int32)
