"""Host-side mirror of the reference's Lineextractor (include/LineExtractor.h:52-92).

Same constructor arguments and call semantics (keylines as KeyLine PODs, N x 32 LBD
descriptors, normalised line equations appended per call), plus batch / device-resident
entry points.  All work happens in libplvi_cuda.so.
"""
import ctypes as C

import numpy as np

from .capi import KEYLINE_DTYPE, check, lib, ptr


class Lineextractor:
    def __init__(self, lsd_nfeatures, lsd_refine, lsd_scale, nlevels, scale, extractor=0,
                 max_width=752, max_height=480, max_batch=1, device=0, stream=None, band_run_max=-1):
        self._h = C.c_void_p()
        check(lib().plvi_line_create_ex(C.byref(self._h), int(lsd_nfeatures), int(lsd_refine), float(lsd_scale),
                                        int(nlevels), float(scale), int(extractor), int(max_width), int(max_height),
                                        int(max_batch), int(device), ptr(stream) if stream else None, int(band_run_max)))
        self.nlevels_l = int(nlevels)
        self.capacity = check(lib().plvi_line_capacity(self._h))
        t = [np.empty(self.nlevels_l, np.float32) for _ in range(4)]
        check(lib().plvi_line_scale_factors(self._h, *[ptr(a) for a in t]))
        # public members read by Frame (src/Frame.cc:569-574); unlike the reference they do
        # not grow on every call (src/LineExtractor.cc:90-101 push_back without clear)
        self.mvScaleFactor_l, self.mvInvScaleFactor_l, self.mvLevelSigma2_l, self.mvInvLevelSigma2_l = t

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().plvi_line_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return lib().plvi_line_stream(self._h)

    @property
    def last_launches(self):
        return lib().plvi_line_last_launches(self._h)

    def graph_stats(self):
        """(captured CUDA graphs, graph replays) of this handle's per-batch launch sequence."""
        import ctypes
        c = ctypes.c_int(0)
        r = lib().plvi_line_graph_stats(self._h, ctypes.byref(c))
        return c.value, r

    def octave_sizes(self, w, h):
        a = [np.empty(self.nlevels_l, np.int32) for _ in range(4)]
        check(lib().plvi_line_octave_sizes(self._h, w, h, *[ptr(x) for x in a]))
        return a

    def __call__(self, image, mask=None):
        """Returns (keylines[KEYLINE_DTYPE], descriptors[N,32] u8, keylineFunction[N,3] f64).
        Raises like the reference on a non-8-bit image or a mask of the wrong shape/type."""
        image = np.asarray(image)
        if mask is not None and getattr(mask, "size", 0) and (mask.shape != image.shape or mask.dtype != np.uint8):
            raise RuntimeError("Mask error while detecting lines: please check its dimensions and that data type is CV_8UC1")
        if image.dtype != np.uint8:
            raise RuntimeError("Error, depth image!= 0")
        if image.ndim != 2:
            raise ValueError("Lineextractor expects a single-channel image")
        kl, desc, eq, counts = self.extract_batch(image[None])
        n = int(counts[0])
        if n < 0:
            check(n)
        return kl[0, :n].copy(), desc[0, :n].copy(), eq[0, :n].copy()

    def alloc_host_outputs(self, n):
        return (np.zeros((n, self.capacity), KEYLINE_DTYPE), np.zeros((n, self.capacity, 32), np.uint8),
                np.zeros((n, self.capacity, 3), np.float64), np.zeros(n, np.int32))

    def extract_batch(self, frames, out=None, sync=True):
        frames = np.asarray(frames)
        assert frames.dtype == np.uint8 and frames.ndim == 3
        if not frames.flags.c_contiguous:
            frames = np.ascontiguousarray(frames)
        n, h, w = frames.shape
        if out is None:
            out = self.alloc_host_outputs(n)
        kl, desc, eq, counts = out
        fn = lib().plvi_line_extract_batch if sync else lib().plvi_line_extract_batch_async
        check(fn(self._h, ptr(frames), n, w, h, frames.strides[1], frames.strides[0], ptr(kl), ptr(desc), ptr(eq),
                 ptr(counts)))
        return kl, desc, eq, counts

    def sync(self):
        check(lib().plvi_line_sync(self._h))

    def alloc_device_outputs(self, n, device):
        import torch
        return (torch.zeros((n, self.capacity, 17), dtype=torch.float32, device=device),
                torch.zeros((n, self.capacity, 32), dtype=torch.uint8, device=device),
                torch.zeros((n, self.capacity, 3), dtype=torch.float64, device=device),
                torch.zeros(n, dtype=torch.int32, device=device))

    def extract_batch_device(self, frames, out=None):
        import torch
        assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 3 and frames.stride(2) == 1
        n, h, w = frames.shape
        if out is None:
            out = self.alloc_device_outputs(n, frames.device)
        kl, desc, eq, counts = out
        check(lib().plvi_line_extract_batch_device(self._h, ptr(frames), n, w, h, frames.stride(1), frames.stride(0),
                                                   ptr(kl), ptr(desc), ptr(eq), ptr(counts)))
        return kl, desc, eq, counts

    # ---- debug read-back (parity tests)
    def set_debug(self, on=True):
        check(lib().plvi_line_set_debug(self._h, int(on)))

    def read_lsd(self, frame, octave, what, w, h):
        ow, oh, sw, sh = self.octave_sizes(w, h)
        W, H = int(sw[octave]), int(sh[octave])
        if what == "scaled":
            out = np.empty((H, W), np.float64)
            check(lib().plvi_line_read_lsd(self._h, frame, octave, 0, ptr(out), 0, None))
        elif what == "angle_deg":
            out = np.empty((H, W), np.float32)
            check(lib().plvi_line_read_lsd(self._h, frame, octave, 1, ptr(out), 0, None))
        elif what == "modgrad":
            out = np.empty((H, W), np.float64)
            check(lib().plvi_line_read_lsd(self._h, frame, octave, 2, ptr(out), 0, None))
        elif what == "segments":
            cap = 16384
            buf = np.empty((cap, 4), np.float32)
            cnt = C.c_int(0)
            check(lib().plvi_line_read_lsd(self._h, frame, octave, 3, ptr(buf), cap, C.byref(cnt)))
            out = buf[:max(cnt.value, 0)].copy()
        elif what == "octave":
            out = np.empty((int(oh[octave]), int(ow[octave])), np.uint8)
            check(lib().plvi_line_read_lsd(self._h, frame, octave, 4, ptr(out), 0, None))
        elif what in ("lbd_image", "lbd_grad"):
            lw, lh = w >> octave, h >> octave
            if what == "lbd_image":
                out = np.empty((lh, lw), np.uint8)
                check(lib().plvi_line_read_lsd(self._h, frame, octave, 5, ptr(out), 0, None))
            else:
                out = np.empty((lh, lw, 2), np.int16)
                check(lib().plvi_line_read_lsd(self._h, frame, octave, 6, ptr(out), 0, None))
        else:
            raise ValueError(what)
        return out
