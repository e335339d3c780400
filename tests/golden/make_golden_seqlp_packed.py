#!/usr/bin/env python
"""Golden fixtures for the PackedSequence branch of sequence_log_probs (_decoding.py:1551-1586)
from the UNMODIFIED reference.  Build container only:

    python tests/golden/make_golden_seqlp_packed.py   ->  tests/golden/seqlp_packed.npz
"""
import json
import os
import sys

import numpy as np
import torch

REF_SRC = os.environ.get("B200LEV_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF_SRC)
import pydrobert.torch.functional as F  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = torch.Generator().manual_seed(77)
    store, params = {}, {}
    k = 0
    for (N, Tmax, V, dim, sorted_) in [(5, 9, 7, 0, False), (5, 9, 7, 1, False), (3, 4, 11, 0, True),
                                       (8, 15, 5, 1, False), (1, 6, 4, 0, True), (6, 12, 30, 0, False)]:
        lens = torch.randint(1, Tmax + 1, (N,), generator=g)
        if sorted_:
            lens, _ = lens.sort(descending=True)
        seqs = [torch.randn(int(n), V, generator=g, requires_grad=False) for n in lens]
        T = int(lens.max())
        hyp = torch.randint(-1, V + 1, (T, N), generator=g)  # includes out-of-range tokens (padding)
        if dim == 1:
            hyp = hyp.t().contiguous()
        data = torch.cat(seqs).clone().requires_grad_(True)
        # rebuild the sequences from `data` so that the gradient lands on one leaf
        parts = list(data.split([int(n) for n in lens]))
        packed = torch.nn.utils.rnn.pack_sequence(parts, enforce_sorted=sorted_)
        out = F.sequence_log_probs(packed, hyp, dim)
        gout = torch.randn(out.shape, generator=g)
        (grad,) = torch.autograd.grad(out, data, gout)
        name = f"ps{k}"
        k += 1
        params[name] = dict(dim=dim, enforce_sorted=sorted_, lens=[int(n) for n in lens])
        store[f"{name}.data"] = data.detach().numpy()
        store[f"{name}.hyp"] = hyp.numpy()
        store[f"{name}.out"] = out.detach().numpy()
        store[f"{name}.grad_out"] = gout.numpy()
        store[f"{name}.grad"] = grad.numpy()
    store["params"] = np.array(json.dumps(params))
    np.savez_compressed(os.path.join(HERE, "seqlp_packed.npz"), **store)
    print(f"seqlp_packed.npz: {len(params)} cases")


if __name__ == "__main__":
    main()
