"""In-tree build of libvfi_b200.so with plain nvcc for sm_100a (no torch extension machinery, no JIT cache).

The shared object lands next to this file so it travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = CSRC / "_obj"
LIB = PKG / "libvfi_b200.so"
SOURCES = ["abi.cu", "warp.cu", "dcn_simt.cu", "dcn_tc.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libvfi_b200.so cannot be built")
    return exe


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link libvfi_b200.so.  Cross-compiles without a GPU."""
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "vfi_b200.h"]
    jobs = []
    for src in SOURCES:
        s, o = CSRC / src, OBJ / (src[:-3] + ".o")
        if force or _stale(o, [s, *headers]):
            jobs.append([nvcc(), *NVCC_FLAGS, "-c", str(s), "-o", str(o)])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            print(" ".join(cmd), r.stdout, r.stderr, sep="\n")
        if r.returncode:
            raise RuntimeError(f"nvcc failed for {cmd[-3]}:\n{r.stderr}")

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(run, jobs))
    objs = [str(OBJ / (s[:-3] + ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        # cuda driver API (cuTensorMapEncodeTiled) is resolved at run time through cudaGetDriverEntryPoint,
        # so the library loads on machines without libcuda (the CPU build container).
        run([nvcc(), "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in os.sys.argv, verbose=True))
