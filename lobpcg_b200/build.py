"""Builds liblobpcg_b200.so in-tree (lobpcg_b200/_lib/) with nvcc for sm_100a only.

    python -m lobpcg_b200.build [--force]

The shared object is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "_lib"
LIB = OUT / "liblobpcg_b200.so"
SOURCES = ["dense.cu", "gram_wl.cu", "gram_tc5.cu", "gram_i8.cu", "nn_tc5.cu", "hostcopy.cu", "elementwise.cu", "spmm.cu", "smalldense.cu", "solver.cu", "capi.cu", "comm.cu", "multigpu.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--expt-relaxed-constexpr",
]
CUDA_LIB = "/usr/local/cuda/lib64"


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found — lobpcg_b200 has no non-CUDA build")


def _stamp() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [HERE.parent / "include" / "lobpcg_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    stamp_file = OUT / "build.stamp"
    stamp = _stamp()
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    nvcc = _nvcc()
    objdir = OUT / "obj"
    objdir.mkdir(exist_ok=True)

    def compile_one(src: str):
        obj = objdir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs),
            f"-L{CUDA_LIB}", "-lcublas", "-lcusolver", "-ldl",
            "-Xlinker", f"-rpath={CUDA_LIB}"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    shutil.rmtree(objdir, ignore_errors=True)
    stamp_file.write_text(stamp)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
