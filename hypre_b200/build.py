"""Builds libhb200.so (the C-ABI library, sm_100a CUDA + NCCL) in-tree with nvcc.

    python -m hypre_b200.build [--force] [--verbose]

The .so is written next to this file (hypre_b200/libhb200.so) so that it travels to the GPU
box with the repository snapshot.  nvcc cross-compiles for sm_100a without a GPU present.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libhb200.so")

SOURCES = [
    "runtime.cu",
    "kernels_spmv.cu",
    "kernels_sell.cu", "kernels_pat.cu", "kernels_offd.cu",
    "kernels_blas1.cu",
    "parcsr.cu",
    "parcsr_peer.cu",
    "relax.cu",
    "gs.cu",
    "amg.cu",
    "krylov.cu",
    "krylov_ext.cu",
    "ij.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-DHB200_WITH_NCCL",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libhb200 cannot be built (there is no CPU fallback)")
    return nvcc


def _newer(a: str, b: str) -> bool:
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "hb200.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(s, o) or hdr_time > os.path.getmtime(o):
            cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    print(out)
    if jobs or not os.path.exists(LIB):
        link = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        run(link)
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
