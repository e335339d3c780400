#!/usr/bin/env python
"""error_rate with non-uniform costs (the (cost, count) cells of the wavefront kernels) on the
config 1 / config 2 shapes: ms per call.  `B200LEV_LIB=<path>` times another build of the library."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import torch

import b200lev._abi as _abi

if os.environ.get("B200LEV_LIB"):
    _abi.LIB_PATH = os.environ["B200LEV_LIB"]
import b200lev.functional as F
import bench

dev = torch.device("cuda", 0)
for cfg, pairs in ((1, 65536), (2, 131072)):
    wl = bench.Workload(cfg)
    r, h, cells = wl.make(pairs, 1)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
    f = lambda: F.error_rate(tr, th, eos=0, ins_cost=3.0, del_cost=3.0, sub_cost=4.0, warn=False)
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"cfg{cfg} shape, {pairs} pairs, error_rate costs 3/3/4: {ms:.4f} ms = {cells / ms / 1e6:.0f} GCUPS")
