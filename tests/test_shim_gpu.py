"""GPU: the reference-shaped C++ shim (ORBextractor / Lineextractor / matchers classes)
runs one frame end to end through libplvi_cuda.so."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def test_cpp_shim_runs(gpu, tmp_path):
    shim = ROOT / "pl_vi_orbslam3_b200" / "shim"
    exe = tmp_path / "shim_smoke"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-o", str(exe), str(shim / "shim_smoke.cpp"), "-I" + str(shim / "include"),
                        "-L" + str(ROOT / "pl_vi_orbslam3_b200"), "-lplvi_cuda",
                        "-Wl,-rpath," + str(ROOT / "pl_vi_orbslam3_b200")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "shim ok" in out.stdout, out.stdout + out.stderr
