import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import b200lev.functional as F
import parity_cases as PC
dev = torch.device("cuda", 0)
wide = len(sys.argv) > 1 and sys.argv[1] == "wide"
shared = not (len(sys.argv) > 2 and sys.argv[2] == "unshared")
PC.check_nbest_batch(F, dev, seed=101, R=101, H=101, n_utts=600, nbest=8, shared=shared, wide=wide)
torch.cuda.synchronize()
print("ok", wide, shared)
