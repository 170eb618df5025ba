"""Builds libplvi_cuda.so (hand-written CUDA for sm_100a) in-tree with nvcc.

`python -m pl_vi_orbslam3_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles
without a GPU; the resulting .so travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libplvi_cuda.so"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--fmad=true", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared",
]  # cudart is linked statically (nvcc default): the .so has no CUDA runtime dependency


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + [PKG.parent / "include" / "plvi.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """One object per .cu (compiled in parallel, rebuilt only when the source or a header is newer), then one link."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    extra = os.environ.get("PLVI_NVCC_EXTRA", "").split()
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    hdr_t = max(d.stat().st_mtime for d in list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "plvi.h"])
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = objdir / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_t):
            return obj, None
        r = subprocess.run([nvcc_path()] + flags + ["-c", "-o", str(obj), str(src)], capture_output=True, text=True)
        return obj, r

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    failed = False
    for obj, r in results:
        if r is not None and (verbose or r.returncode):
            sys.stderr.write(r.stdout + r.stderr)
        failed |= r is not None and r.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libplvi_cuda.so")
    r = subprocess.run([nvcc_path()] + NVCC_FLAGS + ["-o", str(LIB)] + [str(o) for o, _ in results], capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed linking libplvi_cuda.so")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
