import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, ctypes
import b200lev.functional as F
from b200lev import _abi
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_configs import timed
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
L = _abi.lib()
buf = (ctypes.c_float * 6)()
for T in (500, 1000, 2000):
    for N in (1, 148, 296, 592):
        ref = torch.from_numpy(rng.integers(1, 64, size=(T, N), dtype=np.int64)).to(dev)
        hyp = torch.from_numpy(rng.integers(1, 64, size=(T, N), dtype=np.int64)).to(dev)
        fn = lambda: F.prefix_edit_distances(ref, hyp, ins_cost=3, del_cost=3, sub_cost=4, warn=False)
        for _ in range(3): fn()
        _abi.check(L.b200lev_profile(1)); fn(); _abi.check(L.b200lev_profile_read(buf, 6)); _abi.check(L.b200lev_profile(0))
        dp_ms = buf[3]
        cells = T * T * N
        print(json.dumps(dict(T=T, N=N, dp_ms=round(dp_ms, 4), cyc_per_step=round(dp_ms * 1e-3 * 1.965e9 / (T + 31 + 3 * 38), 1), gcups=round(cells / dp_ms / 1e6, 1))))
