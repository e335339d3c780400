import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import torch
import b200lev.functional as F
import bench
dev = torch.device("cuda", 0)
wl = bench.Workload(2)
r, h, cells = wl.make(131072, 1)
tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
f = lambda: F.prefix_error_rates(tr, th, eos=0, warn=False)
def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        out = f()
    e1.record()
    tq = time.perf_counter() - t0
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4), round(tq / n * 1e3, 4)
for _ in range(5):
    out = f()
torch.cuda.synchronize()
print("after warmup 5:", [timed(20) for _ in range(3)])
t = time.perf_counter()
while time.perf_counter() - t < 0.5:
    for _ in range(10):
        out = f()
    torch.cuda.synchronize()
print("after soak:", [timed(20) for _ in range(4)])
print("200:", timed(200))
print("mem", torch.cuda.memory_allocated() >> 20, torch.cuda.memory_reserved() >> 20)
