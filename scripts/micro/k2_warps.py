"""cfg5 (T=2000, costs 3/3/4) literal and saturated with B200LEV_CTA_WARPS = 4, 8, 16"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import torch
import b200lev.functional as F
import bench
dev = torch.device("cuda", 0)
wl = bench.Workload(5)
for pairs in (256, 1184):
    r, h, cells = wl.make(pairs, 1)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
    for nw in sys.argv[1:] or ["4", "8", "16"]:
        os.environ["B200LEV_CTA_WARPS"] = nw
        f = lambda: F.prefix_edit_distances(tr, th, eos=0, ins_cost=3., del_cost=3., sub_cost=4., warn=False)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"pairs {pairs} warps {nw}: {ms:.4f} ms  {cells / ms / 1e6:.1f} GCUPS", flush=True)
