import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lev.functional as F
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_configs import seqs, timed
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
r, rl = seqs(rng, 2001, 256, 64, 200, 2000, 0, 0)
h, hl = seqs(rng, 2001, 256, 64, 200, 2000, 0, 0)
tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
cells = int((rl.astype(np.int64) * hl).sum())
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for name, costs in (("prefix_edit_distances", (3, 3, 4)), ("prefix_error_rates", (3, 3, 4)), ("prefix_edit_distances", (1, 1, 1))):
    fn = getattr(F, name)
    ms = timed(lambda: fn(tr, th, eos=0, ins_cost=costs[0], del_cost=costs[1], sub_cost=costs[2], warn=False), reps)
    print(json.dumps(dict(call=name, costs=costs, ms=ms, gcups=cells / ms / 1e6)))
