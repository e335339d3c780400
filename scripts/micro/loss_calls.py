#!/usr/bin/env python
"""Host cost of the two losses at a training-step batch (64 utterances): forward without grad,
forward with grad, forward + backward."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import torch

import b200lev.functional as F
import bench

dev = torch.device("cuda", 0)


def timeit(f, n=100):
    for _ in range(10):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


wl = bench.Workload(2)
utts = 64
r, h, cells = wl.make(utts * 8, 1)
ref = torch.from_numpy(r[:, ::8].copy()).to(dev)
hyp = torch.from_numpy(h).to(dev).view(h.shape[0], utts, 8)
lp = torch.randn(utts, 8, device=dev, requires_grad=True)
with torch.no_grad():
    a = timeit(lambda: F.minimum_error_rate_loss(lp, ref, hyp, eos=0, warn=False))
b = timeit(lambda: F.minimum_error_rate_loss(lp, ref, hyp, eos=0, warn=False))
c = timeit(lambda: F.minimum_error_rate_loss(lp, ref, hyp, eos=0, warn=False).backward())
print(f"minimum_error_rate_loss 64 x 8, T=100: no_grad {a:.0f} us, forward {b:.0f} us, forward+backward {c:.0f} us")

wl3 = bench.Workload(3)
r, h, cells = wl3.make(64, 1)
ref, hyp = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
lg = torch.randn(h.shape[0], 64, 32, device=dev, requires_grad=True)
with torch.no_grad():
    a = timeit(lambda: F.hard_optimal_completion_distillation_loss(lg, ref, hyp, eos=0, warn=False))
b = timeit(lambda: F.hard_optimal_completion_distillation_loss(lg, ref, hyp, eos=0, warn=False))
c = timeit(lambda: F.hard_optimal_completion_distillation_loss(lg, ref, hyp, eos=0, warn=False).backward())
d = timeit(lambda: F.optimal_completion(ref, hyp, eos=0, warn=False))
print(f"hard OCD loss 64 pairs, T=200, V=32: no_grad {a:.0f} us, forward {b:.0f} us, forward+backward {c:.0f} us; "
      f"optimal_completion alone {d:.0f} us")
