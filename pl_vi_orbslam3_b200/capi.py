"""ctypes bindings of include/plvi.h.  Fails loudly when libplvi_cuda.so is missing."""
import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libplvi_cuda.so"

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")]
)
KEYLINE_DTYPE = np.dtype(
    [("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"), ("pt_y", "<f4"),
     ("response", "<f4"), ("size", "<f4"), ("startPointX", "<f4"), ("startPointY", "<f4"),
     ("endPointX", "<f4"), ("endPointY", "<f4"), ("sPointInOctaveX", "<f4"),
     ("sPointInOctaveY", "<f4"), ("ePointInOctaveX", "<f4"), ("ePointInOctaveY", "<f4"),
     ("lineLength", "<f4"), ("numOfPixels", "<i4")]
)
assert KEYPOINT_DTYPE.itemsize == 28 and KEYLINE_DTYPE.itemsize == 68


class PlviError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"plvi error {code}: {msg}")
        self.code = code


_lib = None
vp, ci, cf, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

_SIGS = {
    "plvi_last_error": (C.c_char_p, []),
    "plvi_device_count": (ci, []),
    "plvi_orb_create": (ci, [C.POINTER(vp), ci, cf, ci, ci, ci, ci, ci, ci, ci, vp]),
    "plvi_orb_destroy": (None, [vp]),
    "plvi_orb_capacity": (ci, [vp]),
    "plvi_orb_levels": (ci, [vp]),
    "plvi_orb_scale_factor": (cf, [vp]),
    "plvi_orb_scale_factors": (ci, [vp, vp, vp, vp, vp]),
    "plvi_orb_features_per_level": (ci, [vp, vp]),
    "plvi_orb_level_sizes": (ci, [vp, ci, ci, vp, vp]),
    "plvi_orb_stream": (vp, [vp]),
    "plvi_orb_extract_batch": (ci, [vp, vp, ci, ci, ci, ci, sz, ci, ci, vp, vp, vp, vp]),
    "plvi_orb_extract_batch_async": (ci, [vp, vp, ci, ci, ci, ci, sz, ci, ci, vp, vp, vp, vp]),
    "plvi_orb_sync": (ci, [vp]),
    "plvi_orb_extract_batch_device": (ci, [vp, vp, ci, ci, ci, ci, sz, ci, ci, vp, vp, vp, vp]),
    "plvi_orb_read_level": (ci, [vp, ci, ci, ci, vp]),
    "plvi_orb_read_candidates": (ci, [vp, ci, ci, vp, ci, vp]),
    "plvi_orb_last_launches": (ci, [vp]),
    "plvi_orb_graph_stats": (ci, [vp, vp]),
    "plvi_orb_wait_event": (ci, [vp, vp]),
    "plvi_orb_wait_counter": (ci, [vp, vp, ci]),
    "plvi_orb_pyramid_device": (ci, [vp, vp, ci, ci, ci, ci, sz]),
    "plvi_orb_wait_event_after_pyramid": (ci, [vp, vp]),
    "plvi_orb_stereo_matches_host": (ci, [vp, vp, vp, vp, ci, vp, vp, ci, cf, cf, vp, vp, vp]),
    "plvi_orb_stereo_matches": (ci, [vp, vp, ci, vp, vp, vp, vp, vp, vp, ci, cf, cf, vp, vp, vp]),
    "plvi_orb_set_profile": (ci, [vp, ci]),
    "plvi_orb_profile": (C.c_char_p, [vp]),
    "plvi_line_set_profile": (ci, [vp, ci]),
    "plvi_line_set_band_run_max": (ci, [vp, ci]),
    "plvi_line_profile": (C.c_char_p, [vp]),
    "plvi_line_create": (ci, [C.POINTER(vp), ci, ci, cf, ci, cf, ci, ci, ci, ci, ci, vp]),
    "plvi_line_create_ex": (ci, [C.POINTER(vp), ci, ci, cf, ci, cf, ci, ci, ci, ci, ci, vp, ci]),
    "plvi_line_destroy": (None, [vp]),
    "plvi_line_capacity": (ci, [vp]),
    "plvi_line_levels": (ci, [vp]),
    "plvi_line_stream": (vp, [vp]),
    "plvi_line_last_launches": (ci, [vp]),
    "plvi_line_graph_stats": (ci, [vp, vp]),
    "plvi_line_stage_event": (vp, [vp]),
    "plvi_line_stage_counter": (vp, [vp, vp]),
    "plvi_line_scale_factors": (ci, [vp, vp, vp, vp, vp]),
    "plvi_line_octave_sizes": (ci, [vp, ci, ci, vp, vp, vp, vp]),
    "plvi_line_extract_batch": (ci, [vp, vp, ci, ci, ci, ci, sz, vp, vp, vp, vp]),
    "plvi_line_extract_batch_async": (ci, [vp, vp, ci, ci, ci, ci, sz, vp, vp, vp, vp]),
    "plvi_line_sync": (ci, [vp]),
    "plvi_line_extract_batch_device": (ci, [vp, vp, ci, ci, ci, ci, sz, vp, vp, vp, vp]),
    "plvi_line_set_debug": (ci, [vp, ci]),
    "plvi_line_read_lsd": (ci, [vp, ci, ci, ci, vp, ci, vp]),
    "plvi_matcher_create": (ci, [C.POINTER(vp), ci, ci, ci, ci, vp]),
    "plvi_matcher_destroy": (None, [vp]),
    "plvi_matcher_stream": (vp, [vp]),
    "plvi_matcher_last_launches": (ci, [vp]),
    "plvi_hamming256": (ci, [vp, vp, vp, ci, ci, vp, ci]),
    "plvi_search_by_projection": (ci, [vp, ci, ci, vp, vp, vp, vp, ci, vp, vp, vp, vp, ci, ci, cf, ci,
                                       vp, vp, vp, ci]),
    "plvi_search_by_bow": (ci, [vp, ci, vp, vp, vp, ci, vp, ci, vp, vp, vp, ci, ci, cf, ci, vp, vp, vp, ci]),
    "plvi_search_by_bow_kf": (ci, [vp, ci, vp, vp, vp, vp, ci, vp, ci, vp, vp, vp, ci, ci, cf, ci, vp, vp, vp, ci]),
    "plvi_search_for_triangulation": (ci, [vp, ci, vp, vp, vp, vp, ci, vp, ci, vp, vp, vp, ci, vp, ci, ci, vp, vp]),
    "plvi_search_in_radius": (ci, [vp, ci, vp, vp, vp, ci, vp, vp, vp, vp, ci, vp, C.c_double, ci, vp, vp, vp]),
    "plvi_line_fuse_search": (ci, [vp, ci, vp, vp, vp, ci, vp, vp, vp, vp, ci, ci, vp, vp, vp]),
    "plvi_queries_from_keypoints": (ci, [vp, vp, vp, ci, ci, cf, cf, vp]),
    "plvi_line_match": (ci, [vp, ci, vp, vp, ci, vp, vp, ci, cf, ci, vp, vp, ci]),
    "plvi_undistort_keypoints": (ci, [vp, vp, vp, ci, ci, vp, vp]),
    "plvi_undistort_keylines": (ci, [vp, vp, vp, ci, ci, vp, vp]),
    "plvi_assign_features_to_grid": (ci, [vp, vp, vp, ci, ci, vp, vp, vp]),
    "plvi_line_match_grid": (ci, [vp, ci, vp, vp, vp, ci, vp, vp, vp, ci, C.c_double, C.c_double, ci, ci, ci, ci, ci, ci, vp, vp]),
    "plvi_line_match_grid_host": (ci, [vp, vp, vp, ci, vp, vp, ci, C.c_double, C.c_double, ci, ci, ci, ci, ci, ci, vp, vp]),
    "plvi_line_match_grid_occ_host": (ci, [vp, vp, vp, ci, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, vp]),
    "plvi_line_stereo_depth": (ci, [vp, ci, vp, vp, ci, vp, vp, ci, vp, vp, cf, vp, vp, vp, vp]),
    "plvi_line_stereo_depth_host": (ci, [vp, vp, ci, vp, ci, vp, vp, cf, vp, vp, vp, vp]),
    "plvi_matcher_set_stereo": (ci, [vp, vp, vp, ci, ci, ci, ci]),
    "plvi_pair_queries": (ci, [vp, vp, vp, ci, ci, ci, cf, cf, vp, vp, cf, vp, vp, vp, vp]),
    "plvi_orb_extract_batch_async_from_line": (ci, [vp, vp, ci, ci, vp, vp, vp, vp]),
    "plvi_line_share_input": (ci, [vp, vp, vp, vp, vp, vp, vp, vp]),
    "plvi_line_share_done": (ci, [vp, vp]),
    "plvi_orb_results_event": (vp, [vp]),
    "plvi_line_results_event": (vp, [vp]),
    "plvi_event_synchronize": (ci, [vp]),
    "plvi_stream_wait_event": (ci, [vp, vp]),
    "plvi_orb_device_results": (ci, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "plvi_line_device_results": (ci, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "plvi_gather_i32": (ci, [vp, vp, ci, ci, ci, vp]),
    "plvi_search_in_radius_host": (ci, [vp, vp, vp, ci, vp, vp, vp, ci, vp, C.c_double, ci, vp, vp, vp, vp, vp]),
    "plvi_search_for_triangulation_host": (ci, [vp, vp, vp, vp, ci, vp, ci, vp, vp, ci, vp, ci, ci, vp, vp]),
    "plvi_line_fuse_search_host": (ci, [vp, vp, vp, ci, vp, vp, vp, ci, ci, vp, vp, vp]),
    "plvi_line_match_mad_host": (ci, [vp, vp, ci, vp, ci, vp, vp, C.c_double, vp, vp, vp]),
    "plvi_distinctive_descriptors_host": (ci, [vp, vp, vp, ci, ci, vp, vp]),
    "plvi_line_match_mad": (ci, [vp, ci, vp, vp, ci, vp, vp, ci, vp, vp, C.c_double, vp, vp, vp]),
    "plvi_distinctive_descriptors": (ci, [vp, vp, vp, ci, ci, vp, vp]),
    "plvi_vocab_create": (ci, [C.POINTER(vp), ci, ci, ci, ci, ci, vp, vp, vp, vp, ci]),
    "plvi_vocab_destroy": (None, [vp]),
    "plvi_vocab_words": (ci, [vp]),
    "plvi_bow_transform": (ci, [vp, vp, vp, vp, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
}

EPIPOLAR_DTYPE = np.dtype([("F12", "<f4", (9,)), ("ep_x", "<f4"), ("ep_y", "<f4"), ("scale_factors", "<f4", (16,)),
                           ("level_sigma2", "<f4", (16,)), ("coarse", "<i4"), ("check_epipole", "<i4")])
assert EPIPOLAR_DTYPE.itemsize == 180
CAMERA_DTYPE = np.dtype([("fx", "<f8"), ("fy", "<f8"), ("cx", "<f8"), ("cy", "<f8"), ("dist", "<f8", (14,)),
                         ("new_fx", "<f8"), ("new_fy", "<f8"), ("new_cx", "<f8"), ("new_cy", "<f8"),
                         ("iters", "<i4"), ("_pad", "<i4")])
assert CAMERA_DTYPE.itemsize == 184
QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("min_level", "<i4"),
                        ("max_level", "<i4"), ("angle", "<f4"), ("flags", "<i4")])
GRID_DTYPE = np.dtype([("min_x", "<f4"), ("min_y", "<f4"), ("inv_w", "<f4"), ("inv_h", "<f4")])
assert QUERY_DTYPE.itemsize == 28


def declared_symbols():
    """Every function name declared in include/plvi.h (parsed from the header)."""
    import re
    hdr = (_PKG.parent / "include" / "plvi.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    # plvi_inline_*: static inline helpers defined in the header itself (no exported symbol)
    return sorted(n for n in set(re.findall(r"\b(plvi_[a-z0-9_]+)\s*\(", hdr)) if not n.startswith("plvi_inline_"))


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m pl_vi_orbslam3_b200.build` "
                "(there is no CPU fallback)")
        _lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGS.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(rc):
    if rc < 0:
        raise PlviError(rc, lib().plvi_last_error().decode(errors="replace"))
    return rc


def ptr(a):
    """void* of a numpy array, torch tensor (host or device), int address or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))
