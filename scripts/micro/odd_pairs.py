"""Per-pair errors of the bulk-scoring call for pair counts that are not multiples of 32."""
import os, sys
sys.path[:0] = [os.path.join(os.path.dirname(__file__), "..", ".."),
                os.path.join(os.path.dirname(__file__), "..", "..", "pydrobert-pytorch_b200")]
import numpy as np, torch
import bench
from oracle import oracle as O
from b200lev import dist as D
import b200lev.functional as F

wl = bench.Workload(4)
dev = torch.device("cuda", 0)
for P in (4100, 4096, 100, 37, 33):
    ref, hyp, cells = wl.make(P, seed=3)
    want = np.asarray(O.error_rate(ref, hyp, eos=-1, norm=False))
    tr, th = torch.from_numpy(ref).to(dev), torch.from_numpy(hyp).to(dev)
    er, acc = D.bulk_error_rate(tr, th, eos=-1)
    got = er.cpu().numpy()
    bad = np.nonzero(got != want)[0]
    er2 = F.error_rate(tr, th, eos=-1, include_eos=False, norm=False).cpu().numpy()
    bad2 = np.nonzero(er2 != want)[0]
    print(P, "sums call mismatches", len(bad), bad[:8], got[bad[:4]], want[bad[:4]], "| error_rate mismatches", len(bad2), bad2[:8],
          "acc", acc.tolist(), float(want.sum()))
