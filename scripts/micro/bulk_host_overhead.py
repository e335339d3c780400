#!/usr/bin/env python
"""Host cost of one bulk-scoring step (dist.bulk_error_rate) against its device time, for the
shard sizes of a strongly scaled config 4."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import torch

import bench
from b200lev import dist as D

dev = torch.device("cuda", 0)
wl = bench.Workload(4)
for pairs in (1000000, 250000, 125000):
    r, h, cells = wl.make(pairs, 1)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
    f = lambda: D.bulk_error_rate(tr, th, eos=-1)
    for _ in range(10):
        f()
    torch.cuda.synchronize()
    n = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"pairs {pairs}: enqueue {t_enq / n * 1e6:.1f} us/call, device {e0.elapsed_time(e1) / n * 1e3:.1f} us/call")
