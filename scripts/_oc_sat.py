import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pydrobert-pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import b200lev.functional as F
from bench_configs import seqs
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
r, rl = seqs(rng, 201, 8192, 32, 100, 200, 0, 0)
h, hl = seqs(rng, 201, 8192, 32, 100, 200, 0, 0)
tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
for _ in range(2):
    F.optimal_completion(tr, th, eos=0, warn=False)
torch.cuda.synchronize()
