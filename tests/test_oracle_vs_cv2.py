"""CPU, build container only: the oracle against a live OpenCV (python cv2, IPP off) and
the reference's own frames.  Skipped where cv2 or /root/reference is absent (the GPU box);
tests/test_oracle_golden.py carries the same checks as committed vectors."""
from pathlib import Path

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted")

import oracle  # noqa: E402
from pl_vi_orbslam3_b200 import synth  # noqa: E402

cv2.ipp.setUseIPP(False)
cv2.setNumThreads(1)


def _imgs():
    out = [cv2.imread(str(REF / f"data2/color/{i}.png"), cv2.IMREAD_UNCHANGED) for i in (2, 5)]
    out.append(cv2.cvtColor(cv2.imread(str(REF / "data/color/4.png")), cv2.COLOR_BGR2GRAY))
    out.append(synth.frame_euroc(9))
    out.append(synth.frame_euroc(9, 1280, 720))
    return out


@pytest.mark.parametrize("idx", range(5))
def test_u8_primitives_bit_exact(idx):
    img = _imgs()[idx]
    h, w = img.shape
    plan = oracle.orb_plan(w, h)
    cur = img
    for l in range(1, 8):
        ref = cv2.resize(cur, (int(plan["w"][l]), int(plan["h"][l])), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(oracle.resize_linear(cur, ref.shape[1], ref.shape[0]), ref)
        cur = ref
    assert np.array_equal(oracle.gaussian_blur7(img), cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
    b5 = cv2.GaussianBlur(img, (5, 5), 1)
    assert np.array_equal(oracle.gaussian_blur5(img), b5)
    assert np.array_equal(oracle.pyr_down(b5), cv2.pyrDown(b5, dstsize=(w // 2, h // 2)))
    dx, dy = oracle.sobel3(b5)
    assert np.array_equal(dx, cv2.Sobel(b5, cv2.CV_16S, 1, 0, ksize=3))
    assert np.array_equal(dy, cv2.Sobel(b5, cv2.CV_16S, 0, 1, ksize=3))
    fd = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True)
    for th in (20, 7):
        fd.setThreshold(th)
        for roi in (img[50:88, 60:96], img[16:54, w - 52:w - 16], img):
            roi = np.ascontiguousarray(roi)
            k = fd.detect(roi)
            ref = np.array([[p.pt[0], p.pt[1], p.response] for p in k], np.float32).reshape(-1, 3)
            assert np.array_equal(oracle.fast_roi(roi, th), ref)


@pytest.mark.parametrize("idx", range(3))
def test_f64_primitives(idx):
    img = _imgs()[idx].astype(np.float64)
    S = float(np.float32(0.8))
    sigma = 0.6 / S
    k = oracle.gaussian_kernel_f64(7, sigma)
    assert np.array_equal(k, cv2.getGaussianKernel(7, sigma, cv2.CV_64F).ravel())
    mine = oracle.gaussian_blur_f64(img, k)
    assert np.abs(mine - cv2.GaussianBlur(img, (7, 7), sigma)).max() <= 1e-12
    ref = cv2.resize(mine, None, fx=S, fy=S, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(oracle.resize_linear_f64(mine, ref.shape[1], ref.shape[0], S, S), ref)


@pytest.mark.parametrize("scale", [0.3, 0.35, 0.4, 0.45, 0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95])
def test_lsd_gaussian_kernels_of_common_scales(scale):
    """The Gaussian of flsd for the tabulated lsd_scale settings (oracle/lsd_gauss_table.h): kernel bit-exact against
    cv2.getGaussianKernel (soft-float exp), blur within 1e-12, f64 resize bit-exact."""
    import math
    S = float(np.float32(scale))
    sigma = 0.6 / S
    n = 1 + 2 * int(math.ceil(sigma * math.sqrt(2 * 3.0 * math.log(10.0))))
    k = oracle.gaussian_kernel_f64(n, sigma)
    assert np.array_equal(k, cv2.getGaussianKernel(n, sigma, cv2.CV_64F).ravel())
    img = _imgs()[3].astype(np.float64)
    mine = oracle.gaussian_blur_f64(img, k)
    assert np.abs(mine - cv2.GaussianBlur(img, (n, n), sigma)).max() <= 1e-12
    ref = cv2.resize(mine, None, fx=S, fy=S, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(oracle.resize_linear_f64(mine, ref.shape[1], ref.shape[0], S, S), ref)


def test_lsd_segment_count_close_to_opencv_lsd():
    """cv2's LSD (4.13) seeds regions by gradient-sorted order, the vendored one in raster
    order, so only the population is comparable."""
    img = _imgs()[0]
    mine = oracle.lsd(img, 0.8)
    ref = cv2.createLineSegmentDetector(0, float(np.float32(0.8))).detect(img)[0]
    assert abs(len(mine) - len(ref)) <= 0.1 * len(ref)
