#!/usr/bin/env python
"""Build libb200lev.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python pydrobert-pytorch_b200/build.py [--force] [--verbose]

The library links only against the CUDA runtime (no torch, no Python): the C ABI in
include/b200lev.h is the whole boundary.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "b200lev")
LIB = os.path.join(OUT_DIR, "libb200lev.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["lev_abi.cu", "lev_pack.cu", "lev_dp.cu", "lev_group.cu", "lev_bitvec.cu", "lev_bvfused.cu", "lev_bvshort.cu", "lev_cta.cu", "lev_completion.cu", "lev_loss.cu", "lev_seqlp.cu", "lev_ragged.cu", "lev_decode.cu", "lev_mask16.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "b200lev.h"))
    return hdrs


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = os.environ.get("B200LEV_NVCC_EXTRA", "").split()  # tuning experiments only
    hdr_m = max(os.path.getmtime(h) for h in _deps())
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for " + cmd[-3])

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(run, jobs))
    if jobs or force or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
