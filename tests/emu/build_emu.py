"""Build the kernel sources against the SIMT emulator (tests/emu/emu_cuda.h) with g++.

TEST INFRASTRUCTURE ONLY: produces tests/emu/_build/libb200lev_emu.so, which only the
`-m "not gpu"` kernel-logic tests load.  The product package never looks for it.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pydrobert-pytorch_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libb200lev_emu.so")
SOURCES = ["lev_abi.cu", "lev_pack.cu", "lev_dp.cu", "lev_group.cu", "lev_bitvec.cu", "lev_bvfused.cu", "lev_bvshort.cu", "lev_cta.cu", "lev_completion.cu", "lev_loss.cu", "lev_seqlp.cu", "lev_ragged.cu", "lev_decode.cu", "lev_mask16.cu"]
FLAGS = ["-O1", "-g", "-std=c++17", "-fPIC", "-fno-fast-math", "-ffp-contract=off", "-w",
         "-include", os.path.join(HERE, "emu_cuda.h"), "-I", HERE]


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps += [os.path.join(HERE, "emu_cuda.h"), os.path.join(ROOT, "include", "b200lev.h")]
    dep_m = max(os.path.getmtime(d) for d in deps)
    jobs, objs = [], []
    for src in SOURCES + ["emu_runtime.cpp"]:
        s = os.path.join(CSRC if src.endswith(".cu") else HERE, src)
        o = os.path.join(OUT, src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), dep_m):
            jobs.append(["g++", "-x", "c++"] + FLAGS + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("g++ failed for " + cmd[-3])

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        r = subprocess.run(["g++", "-shared", "-o", LIB] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("emu link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
