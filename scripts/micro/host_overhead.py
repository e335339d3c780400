#!/usr/bin/env python
"""Is the n-best call CPU-bound?  Enqueue time per call (no sync) vs device time per call,
and a cProfile of the host layer."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import b200lev.functional as F
import bench

dev = torch.device("cuda", 0)
cases = [(2, 131072, "prefix_error_rates"), (2, 512, "prefix_error_rates"), (1, 32, "error_rate")]
if len(sys.argv) > 1:
    cases = [c for c in cases if str(c[1]) in sys.argv[1:]]
for cfg, pairs, fn in cases:
    wl = bench.Workload(cfg)
    r, h, cells = wl.make(pairs, 1)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
    f = (lambda: F.prefix_error_rates(tr, th, eos=0, warn=False)) if fn == "prefix_error_rates" else (
        lambda: F.error_rate(tr, th, eos=0, warn=False))
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
    t_all = time.perf_counter() - t0
    print(f"pairs {pairs}: enqueue {t_enq / n * 1e6:.1f} us/call, wall incl. sync {t_all / n * 1e6:.1f} us/call, "
          f"device {e0.elapsed_time(e1) / n * 1e3:.1f} us/call")
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    f()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
