import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import torch
import b200lev.functional as F
dev = torch.device("cuda", 0)
T, N, V = 101, 512, 10000
g = torch.Generator().manual_seed(1)
hyp = torch.randint(1, V, (T, N), generator=g)
lens = torch.randint(50, T + 1, (N,), generator=g)
hyp[torch.arange(T)[:, None] >= (lens - 1)[None, :]] = 0
hyp = hyp.to(dev)
lg = (torch.randn(T, N, V, device=dev) * 2).to(torch.bfloat16).requires_grad_(True)
for _ in range(3):
    o = F.sequence_log_probs(lg, hyp, 0, eos=0); o.sum().backward(); lg.grad = None
torch.cuda.synchronize()
