"""Where the host-tensor (e2e) call spends its time: raw PCIe copies (1-D, 2-D column blocks,
both directions at once) next to the public call with and without the block pipeline."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "pydrobert-pytorch_b200"))
sys.path.insert(0, ROOT)
import b200lev.functional as F  # noqa: E402
from b200lev import _abi  # noqa: E402
from bench import make_batch, NBEST  # noqa: E402

dev = torch.device("cuda", 0)
L = _abi.lib()


def wall(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


T, N = 101, 131072
h = torch.empty((T, N), dtype=torch.int64).pin_memory()
d = torch.empty((T, N), dtype=torch.int64, device=dev)
mb = T * N * 8 / 1e6
ms = wall(lambda: d.copy_(h, non_blocking=True))
print(f"H2D 1-D {mb:.0f} MB: {ms:.3f} ms {mb / ms:.1f} GB/s")
for blocks in (4, 8, 16):
    nb = N // blocks
    dk = [torch.empty((T, nb), dtype=torch.int64, device=dev) for _ in range(blocks)]
    st = torch.cuda.current_stream().cuda_stream

    def f():
        for k in range(blocks):
            _abi.check(L.b200lev_copy2d_async(dk[k].data_ptr(), nb * 8, h.data_ptr() + k * nb * 8, N * 8,
                                              nb * 8, T, 1, st))
    ms = wall(f)
    print(f"H2D 2-D {blocks} column blocks: {ms:.3f} ms {mb / ms:.1f} GB/s")
o = torch.empty((T, N), dtype=torch.float32, device=dev)
oh = torch.empty((T, N), dtype=torch.float32).pin_memory()
ms = wall(lambda: oh.copy_(o, non_blocking=True))
print(f"D2H 1-D {mb / 2:.0f} MB: {ms:.3f} ms {mb / 2 / ms:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def duplex():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        oh.copy_(o, non_blocking=True)


ms = wall(duplex)
print(f"duplex 2x{mb:.0f} MB H2D + {mb / 2:.0f} MB D2H: {ms:.3f} ms")
ms = wall(lambda: (d.copy_(h, non_blocking=True), d.copy_(h, non_blocking=True)))
print(f"2x H2D alone: {ms:.3f} ms")

ref_np, hyp_np, cells = make_batch(N // NBEST, seed=1)
ref_h = torch.from_numpy(np.repeat(ref_np, NBEST, axis=1)).pin_memory()
hyp_h = torch.from_numpy(hyp_np).pin_memory()
for blocks in (0, 2, 4, 8, 16, 32):
    if blocks == 0:
        F._PIPE_MIN_BYTES = 1 << 60
    else:
        F._PIPE_MIN_BYTES = 1 << 20
        F._PIPE_BLOCKS = blocks
    ms = wall(lambda: F.prefix_error_rates(ref_h, hyp_h, eos=0, warn=False))
    print(f"public call, host tensors, blocks={blocks}: {ms:.3f} ms  {cells / ms / 1e6:.1f} GCUPS")
rd, hd = ref_h.to(dev), hyp_h.to(dev)
ms = wall(lambda: F.prefix_error_rates(rd, hd, eos=0, warn=False), 20)
print(f"public call, device tensors: {ms:.3f} ms")
# page-locked allocation cost of the result
ms = wall(lambda: torch.empty((T, N), dtype=torch.float32, pin_memory=True))
print(f"pinned result allocation (cached): {ms:.3f} ms")
