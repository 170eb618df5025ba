import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    """The CUDA library travels with the repo snapshot; build it only if it is missing."""
    from pl_vi_orbslam3_b200 import build
    if not build.LIB.exists():
        build.build(force=True)


_ensure_built()


def _have_gpu():
    try:
        from pl_vi_orbslam3_b200 import capi
        return capi.lib().plvi_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """GPU tests must not silently skip on the GPU box: missing library or device is a failure."""
    from pl_vi_orbslam3_b200 import capi
    n = capi.lib().plvi_device_count()
    assert n > 0, "no CUDA device: " + capi.lib().plvi_last_error().decode()
    return n
