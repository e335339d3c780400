import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200"), os.path.join(ROOT, "scripts")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lev.functional as F
from bench_configs import seqs
rng = np.random.default_rng(0)
dev = torch.device("cuda", 0)
r, rl = seqs(rng, 201, 128, 32, 100, 200, 0, 0)
h, hl = seqs(rng, 201, 128, 32, 100, 200, 0, 0)
tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
logits = torch.randn(201, 128, 32, device=dev, requires_grad=True)
for _ in range(3):
    F.optimal_completion(tr, th, eos=0, warn=False)
    loss = F.hard_optimal_completion_distillation_loss(logits, tr, th, eos=0, warn=False); loss.backward()
torch.cuda.synchronize()
