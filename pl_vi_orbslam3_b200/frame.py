"""Frame steps that follow the extractors, on device-resident extractor outputs:
Frame::UndistortKeyPoints / UndistortKeyLines (src/Frame.cc:1124-1197) and
Frame::AssignFeaturesToGrid (src/Frame.cc:644-675).  torch tensors carry the device
memory; every kernel is in libplvi_cuda.so."""
import numpy as np

from .capi import CAMERA_DTYPE, GRID_DTYPE, check, lib, ptr

FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48


def make_camera(fx, fy, cx, cy, dist=(), new_k=None, iters=0):
    """plvi_camera.  K and distCoeffs pass through float32 like the reference's CV_32F cv::Mat
    (Tracking::ParseCamParamFile); P defaults to K (cv::undistortPoints(..., cv::Mat(), mK))."""
    c = np.zeros(1, CAMERA_DTYPE)
    K = np.array([fx, fy, cx, cy], np.float32).astype(np.float64)
    c["fx"], c["fy"], c["cx"], c["cy"] = K
    d = np.asarray(dist, np.float32).astype(np.float64)
    c["dist"][0, :len(d)] = d
    P = K if new_k is None else np.asarray(new_k, np.float64)
    c["new_fx"], c["new_fy"], c["new_cx"], c["new_cy"] = P
    c["iters"] = iters
    return c


def _stream_ptr(stream):
    return int(stream.cuda_stream) if stream is not None else 0


def undistort_keypoints(d_kps, d_counts, cam, out=None, stream=None):
    """d_kps: CUDA tensor [n, cap, 7] float32 view of plvi_keypoint records (as the extractor's
    device API returns them), d_counts int32 [n].  Returns mvKeysUn in the same layout."""
    import torch
    if out is None:
        out = torch.empty_like(d_kps)
    n, cap = d_kps.shape[0], d_kps.shape[1]
    check(lib().plvi_undistort_keypoints(_stream_ptr(stream), ptr(d_kps), ptr(d_counts), n, cap, ptr(cam), ptr(out)))
    return out


def undistort_keylines(d_kl, d_counts, cam, out=None, stream=None):
    """d_kl: CUDA tensor [n, cap, 17] float32 view of plvi_keyline records -> mvKeysUn_Line."""
    import torch
    if out is None:
        out = torch.empty_like(d_kl)
    n, cap = d_kl.shape[0], d_kl.shape[1]
    check(lib().plvi_undistort_keylines(_stream_ptr(stream), ptr(d_kl), ptr(d_counts), n, cap, ptr(cam), ptr(out)))
    return out


def assign_features_to_grid(d_kps_un, d_counts, grid, stream=None):
    """mGrid of every frame as a CSR: (cell_start int32 [n, 64*48+1], items int32 [n, cap]); cell (i, j) of
    frame f = items[f, cell_start[f, i*48+j] : cell_start[f, i*48+j+1]], in keypoint order."""
    import torch
    n, cap = d_kps_un.shape[0], d_kps_un.shape[1]
    start = torch.empty((n, FRAME_GRID_COLS * FRAME_GRID_ROWS + 1), dtype=torch.int32, device=d_kps_un.device)
    items = torch.zeros((n, cap), dtype=torch.int32, device=d_kps_un.device)
    g = np.ascontiguousarray(grid, GRID_DTYPE)
    check(lib().plvi_assign_features_to_grid(_stream_ptr(stream), ptr(d_kps_un), ptr(d_counts), n, cap, ptr(g), ptr(start), ptr(items)))
    return start, items
