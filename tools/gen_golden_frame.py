"""Golden vectors for the Frame steps after extraction (tests/golden/undistort_euroc.npz).

cv2 (4.13, the oracle's OpenCV) is the reference implementation of cv::undistortPoints; the
points are a seeded grid + random sample over the EuRoC image, camera = EuRoC.yaml cam0.
    python tools/gen_golden_frame.py
"""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402


def main():
    cv2.setNumThreads(1)
    cam = oracle.EUROC_CAMERA
    K = np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], np.float32)
    D = np.array(cam["dist"], np.float32)
    rng = np.random.RandomState(42)
    gx, gy = np.meshgrid(np.linspace(0, 751, 24), np.linspace(0, 479, 16))
    pts = np.concatenate([np.stack([gx.ravel(), gy.ravel()], 1), rng.rand(640, 2) * [752, 480]]).astype(np.float32)
    und = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    np.savez_compressed(ROOT / "tests" / "golden" / "undistort_euroc.npz", points=pts, undistorted=und,
                        K=K, D=D, cv2_version=np.array(cv2.__version__))
    print("wrote", len(pts), "points; max shift", np.abs(und - pts).max())


if __name__ == "__main__":
    main()
