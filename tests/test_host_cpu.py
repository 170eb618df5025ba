"""CPU: C-ABI library loads and exports every symbol include/plvi.h declares, the host
logic (sharding, grid parameters, synthetic inputs) and the N>1 path on gloo."""
import ctypes
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_all_declared_symbols():
    from pl_vi_orbslam3_b200 import capi
    lib = capi.lib()
    names = capi.declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    unbound = [n for n in names if n not in capi._SIGS]
    assert not unbound, f"ctypes signatures missing for {unbound}"


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    from pl_vi_orbslam3_b200 import capi
    lib = capi.lib()
    if lib.plvi_device_count() > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.plvi_orb_create(ctypes.byref(h), 1000, 1.2, 8, 20, 7, 752, 480, 1, 0, None)
    assert rc == -2 and not h.value
    assert b"cuda" in lib.plvi_last_error().lower()
    from pl_vi_orbslam3_b200 import ORBextractor, PlviError
    with pytest.raises(PlviError):
        ORBextractor(1000, 1.2, 8, 20, 7)


def test_product_does_not_import_oracle():
    for f in (ROOT / "pl_vi_orbslam3_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp") and f.is_file():
            txt = f.read_text(errors="replace")
            assert "import oracle" not in txt and "libplvi_oracle" not in txt and "oracle/" not in txt.replace("oracle/oracle_", ""), f


def test_pod_layouts():
    from pl_vi_orbslam3_b200.capi import GRID_DTYPE, KEYLINE_DTYPE, KEYPOINT_DTYPE, QUERY_DTYPE
    assert (KEYPOINT_DTYPE.itemsize, KEYLINE_DTYPE.itemsize, QUERY_DTYPE.itemsize, GRID_DTYPE.itemsize) == (28, 68, 28, 16)
    assert KEYLINE_DTYPE.names[:3] == ("angle", "class_id", "octave") and KEYLINE_DTYPE.names[-1] == "numOfPixels"


def test_shard_range_partitions_exactly():
    from pl_vi_orbslam3_b200.frontend import shard_range
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(4096, 3, 8) == (1536, 2048)


def test_frame_grid_matches_reference_formula():
    from pl_vi_orbslam3_b200 import frame_grid
    g = frame_grid(0, 752, 0, 480)
    assert g["inv_w"][0] == np.float32(64) / np.float32(752) and g["inv_h"][0] == np.float32(48) / np.float32(480)


def test_synthetic_frames_are_deterministic_and_textured():
    from pl_vi_orbslam3_b200 import synth
    a, b = synth.frame_euroc(3), synth.frame_euroc(3)
    assert a.shape == (480, 752) and a.dtype == np.uint8 and np.array_equal(a, b)
    assert not np.array_equal(a, synth.frame_euroc(4))
    import oracle
    assert len(oracle.grid_fast(a)) >= 1500 and len(oracle.lsd(a, 0.8)) >= 300   # SURVEY 8(d) C1 acceptance
    batch = synth.frame_batch(20, distinct=4)
    assert batch.shape == (20, 480, 752) and len({bytes(f[:8].tobytes()) for f in batch}) == 20
    f1, f2, A = synth.warp_pair(1)
    assert f1.shape == f2.shape and A.shape == (2, 3)


_GLOO_WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["PLVI_ROOT"])
import numpy as np, torch, torch.distributed as dist
import oracle
from pl_vi_orbslam3_b200 import synth
from pl_vi_orbslam3_b200.frontend import shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N = 6
lo, hi = shard_range(N, rank, world)
# each rank extracts its own contiguous frame range; there is no data-path collective
counts = [len(oracle.orb_extract(synth.frame_euroc(100 + i))["keypoints"]) for i in range(lo, hi)]
t = torch.tensor([float(10 + rank)], dtype=torch.float64)      # stand-in for the per-rank device time
dist.barrier()
dist.all_reduce(t, op=dist.ReduceOp.MAX)                        # bench.py: max over ranks
gathered = [None] * world
dist.all_gather_object(gathered, (lo, hi, counts))
if rank == 0:
    print(json.dumps({"max_ms": float(t[0]), "shards": gathered}))
dist.destroy_process_group()
'''


def test_two_rank_sharding_over_gloo(tmp_path):
    """world_size-2 run of the sharded path: contiguous frame ranges, no collective on the
    data path, max-over-ranks timing, rank 0 reports."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, PLVI_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["max_ms"] == 11.0
    shards = sorted(res["shards"])
    assert [s[:2] for s in shards] == [[0, 3], [3, 6]]
    import oracle
    from pl_vi_orbslam3_b200 import synth
    assert shards[1][2][0] == len(oracle.orb_extract(synth.frame_euroc(103))["keypoints"])


def test_bench_reference_arm_contract():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == ("reference" if __import__("oracle").ref_available() else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0


def test_cpp_shim_compiles_and_links(tmp_path):
    """The reference-shaped C++ classes (shim/*.h) build against include/plvi.h + libplvi_cuda.so."""
    shim = ROOT / "pl_vi_orbslam3_b200" / "shim"
    exe = tmp_path / "shim_smoke"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-o", str(exe), str(shim / "shim_smoke.cpp"), "-I" + str(shim / "include"),
                        "-L" + str(ROOT / "pl_vi_orbslam3_b200"), "-lplvi_cuda",
                        "-Wl,-rpath," + str(ROOT / "pl_vi_orbslam3_b200")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([str(exe), "tolerate-no-device"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_kernel_integer_identities():
    """Integer identities the kernels lean on (csrc/orb_kernels.cu), checked exhaustively on the host.
    * k_fast pass A gathers the four 0/1 flag bytes of a SWAR comparison into a nibble with one 32-bit multiplication:
      byte k (bit 8k) times 2^(21 - 7k) lands on bit 21 + k and no two of the sixteen partial products share a bit.
    * k_fast / k_blur7 map 256 threads onto 7 rows x 36 columns with tid * 1821 >> 16 = tid // 36."""
    for m in range(16):
        cnt = sum(((m >> k) & 1) << (8 * k) for k in range(4))
        assert (((cnt * 0x00204081) & 0xFFFFFFFF) >> 21) & 0xF == m
    bits = [8 * j + 21 - 7 * i for i in range(4) for j in range(4)]
    assert len(set(bits)) == 16                      # no two partial products collide (so no carries either)
    for tid in range(256):
        assert (tid * 1821) >> 16 == tid // 36
