"""B200-native front-end hot path of PL-VI-ORBSLAM3 (ORB + LSD/LBD + Hamming search).

All compute lives in libplvi_cuda.so (hand-written CUDA for sm_100a, C ABI declared in
include/plvi.h).  This package is the thin Python host layer used by the tests and the
benchmark: ctypes bindings (`capi`) and mirrors of the reference's operator classes
(`ORBextractor`, ...).  There is no CPU fallback: importing the bindings without the
built library raises.
"""
from .capi import lib, PlviError, KEYPOINT_DTYPE, KEYLINE_DTYPE  # noqa: F401
from .orbextractor import ORBextractor  # noqa: F401
from .matchers import ORBmatcher, LineMatcher, FrameView, frame_grid  # noqa: F401
from .lineextractor import Lineextractor  # noqa: F401
