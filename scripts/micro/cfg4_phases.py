"""Per-kernel times (b200lev_profile) of the cfg4 bulk call: 1 M unrelated pairs, T = 31."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import b200lev.functional as F
from b200lev import _abi
rng = np.random.default_rng(0)
T, N = 31, 1_000_000
def seqs():
    tok = rng.integers(1, 5000, size=(T, N), dtype=np.int64)
    lens = rng.integers(5, T + 1, size=N)
    pos = np.arange(T)[:, None]
    tok[pos >= (lens - 1)[None, :]] = 0
    return torch.from_numpy(tok).cuda()
ref, hyp = seqs(), seqs()
L = _abi.lib()
names = ["pack_ref", "pack_hyp", "sort", "dp", "finalize", "standby", "bv_uid", "bv_dp"]
for _ in range(3):
    F.error_rate(ref, hyp, eos=0, warn=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    F.error_rate(ref, hyp, eos=0, warn=False)
b.record(); torch.cuda.synchronize()
print("whole call ms", a.elapsed_time(b) / 10)
_abi.check(L.b200lev_profile(1))
buf = (ctypes.c_float * 8)()
acc = np.zeros(8)
for _ in range(5):
    F.error_rate(ref, hyp, eos=0, warn=False)
    _abi.check(L.b200lev_profile_read(buf, 8))
    acc += np.array([max(x, 0) for x in buf])
print({n: round(v / 5 * 1000, 1) for n, v in zip(names, acc)}, "us")
