import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, ctypes
from b200lev import _abi, _ops
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
L = _abi.lib()
T = 2000
ref = torch.from_numpy(rng.integers(1, 64, size=(T, 1), dtype=np.int64)).to(dev)
hyp = torch.from_numpy(rng.integers(1, 64, size=(T, 1), dtype=np.int64)).to(dev)
rt, ht = _ops._tok_struct(ref, False), _ops._tok_struct(hyp, False)
o = _ops._opts(None, True, 3.0, 3.0, 4.0, False, False, -100, False, 1)
nbytes = L.b200lev_workspace_bytes(ctypes.byref(rt), ctypes.byref(ht), 2, 0)
ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
out = torch.empty((T + 1, 1), dtype=torch.float32, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
for _ in range(3):
    _abi.check(L.b200lev_prefix(ctypes.byref(rt), ctypes.byref(ht), ctypes.byref(o), out.data_ptr(), 1, 1, ws.data_ptr(), nbytes, None, st))
torch.cuda.synchronize()
# locate gmeta inside the workspace: scan for the nsteps marker
w = ws.view(torch.int32).cpu().numpy()
idx = np.where(w == T + 31)[0]
for i in idx[:8]:
    print("strip cycles", int(w[i - 1]), "steps", int(w[i]), "cycles/step", round(w[i - 1] / w[i], 1))
