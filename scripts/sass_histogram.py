#!/usr/bin/env python
"""SASS opcode histogram per kernel family of libb200lev.so -> profiles/r2_sass_opcodes.txt.

    python scripts/sass_histogram.py > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pydrobert-pytorch_b200", "b200lev", "libb200lev.so")
NOTE = ("UBLKCP", "SYNCS", "VIADDMNMX", "VIMNMX", "VIMNMX3", "LDGSTS", "ATOMS", "ATOMG", "REDG", "PRMT", "SHFL",
        "POPC", "CCTL", "MUFU", "REDUX")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fam = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"^void ", "", dem).split("<")[0].split("(")[0]
            cur = fam.setdefault(name, {"n": 0, "ops": collections.Counter(), "wide": collections.Counter()})
            cur["n"] += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur["ops"][m.group(1)] += 1
            if m.group(1) in ("LDG", "STG", "LDS", "STS") and ".128" in m.group(2):
                cur["wide"][m.group(1) + ".128"] += 1
    print("# SASS opcode histogram per kernel family of libb200lev.so (cuobjdump -sass, all template\n"
          "# instantiations of a family added up; scripts/sass_histogram.py); what to look for: UBLKCP / SYNCS =\n"
          "# TMA bulk copies on an mbarrier (lev_cta_kernel), VIADDMNMX / VIMNMX(3) = DPX min/add (wavefront and\n"
          "# short-reference kernels), LDGSTS = cp.async, LDG.128 / STG.128 = 128-bit global accesses, ATOMS =\n"
          "# shared-memory hash-table claims, CCTL = L2 prefetches.\n")
    for name, f in fam.items():
        total = sum(f["ops"].values())
        print(f"{name}  ({f['n']} instantiations, {total} instructions)")
        print("  " + ", ".join(f"{k} {v}" for k, v in f["ops"].most_common(14)))
        note = [f"{k} {f['ops'][k]}" for k in NOTE if f["ops"].get(k)] + [f"{k} {v}" for k, v in f["wide"].items()]
        if note:
            print("  of note: " + ", ".join(note))
        print()


if __name__ == "__main__":
    sys.exit(main())
