import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pydrobert-pytorch_b200"))
import b200lev.functional as F
dev = torch.device("cuda", 0)
lg = (torch.randn(100, 512, 10000, device=dev) * 2).to(torch.bfloat16)
for _ in range(3):
    out = F.ctc_greedy_search(lg, None, 0)
torch.cuda.synchronize()
