import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, ctypes
import b200lev.functional as F
from b200lev import _abi
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
L = _abi.lib()
buf = (ctypes.c_float * 6)()
def run(tag, fn):
    for _ in range(3): fn()
    _abi.check(L.b200lev_profile(1)); fn(); _abi.check(L.b200lev_profile_read(buf, 6)); _abi.check(L.b200lev_profile(0))
    return buf[3]
for T in (256, 1000, 2000):
    ref = torch.from_numpy(rng.integers(1, 64, size=(T, 1), dtype=np.int64)).to(dev)
    hyp = torch.from_numpy(rng.integers(1, 64, size=(T, 1), dtype=np.int64)).to(dev)
    res = {}
    for cta in ("1", "0"):
        os.environ["B200LEV_CTA_KERNEL"] = cta
        res[f"prefix_cta{cta}"] = run("p", lambda: F.prefix_edit_distances(ref, hyp, ins_cost=3, del_cost=3, sub_cost=4, warn=False))
        res[f"final_cta{cta}"] = run("f", lambda: F.edit_distance(ref, hyp, ins_cost=3, del_cost=3, sub_cost=4, warn=False))
    print(json.dumps(dict(T=T, **{k: round(v * 1e-3 * 1.965e9 / T, 1) for k, v in res.items()})), "cycles/row")
