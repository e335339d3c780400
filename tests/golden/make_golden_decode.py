#!/usr/bin/env python
"""Golden fixtures of the N-best producers' step functions from the UNMODIFIED reference
(_decoding.py:41-155 beam_search_advance, :1207-1283 random_walk_advance).

    python tests/golden/make_golden_decode.py        (build container only: imports /root/reference/src)

Writes tests/golden/decode.npz.  Inputs are tie-free random floats (torch.topk leaves the order
of equal scores unspecified); random_walk cases record the tokens torch.multinomial drew, so the
path bookkeeping can be replayed without torch's random stream."""
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
    g = torch.Generator().manual_seed(20260)
    store, params = {}, {}
    k = 0
    for dtype in (torch.float32, torch.float64):
        for (N, Kp, V, W, S, lens_mode) in [
            (4, 1, 20, 1, 0, "none"), (4, 1, 20, 5, 0, "zeros"), (3, 4, 7, 5, 6, "none"),
            (3, 4, 7, 5, 6, "ragged"), (3, 4, 7, 5, 6, "full"), (2, 3, 5, 40, 4, "ragged"),
            (5, 8, 300, 8, 12, "ragged"), (1, 16, 256, 40, 9, "short"), (2, 2, 2, 3, 1, "full"),
            (6, 5, 33, 7, 3, "short"),
        ]:
            lpt = torch.randn(N, Kp, V, generator=g, dtype=dtype)
            lpp = torch.randn(N, Kp, generator=g, dtype=dtype)
            yp = torch.randint(0, V, (S, N, Kp), generator=g)
            if lens_mode == "none":
                lens = None
            elif lens_mode == "zeros":
                lens = torch.zeros(N, Kp, dtype=torch.long)
            elif lens_mode == "full":
                lens = torch.full((N, Kp), S, dtype=torch.long)
            elif lens_mode == "short":  # nobody reaches the last row: y does not grow
                lens = torch.randint(0, max(S, 1), (N, Kp), generator=g)
            else:
                lens = torch.randint(0, S + 1, (N, Kp), generator=g)
                lens[0, 0] = S
            y_next, y_lens, lp_next, src = F.beam_search_advance(lpt, W, lpp, yp, lens)
            name = f"beam{k}"
            k += 1
            params[name] = dict(width=W, has_lens=lens is not None, K=min(W, Kp * V))
            store[f"{name}.log_probs_t"] = lpt.numpy()
            store[f"{name}.log_probs_prev"] = lpp.numpy()
            store[f"{name}.y_prev"] = yp.numpy()
            if lens is not None:
                store[f"{name}.y_prev_lens"] = lens.numpy()
            store[f"{name}.y_next"] = y_next.numpy()
            store[f"{name}.y_next_lens"] = y_lens.numpy()
            store[f"{name}.log_probs_next"] = lp_next.numpy()
            store[f"{name}.next_src"] = src.numpy()
    k = 0
    for (N, V, S, lens_mode) in [(5, 11, 0, "none"), (5, 11, 4, "none"), (5, 11, 4, "ragged"),
                                 (7, 3, 6, "short"), (1, 40, 2, "full"), (9, 64, 8, "ragged")]:
        lpt = torch.randn(N, V, generator=g).log_softmax(1)
        lpp = torch.randn(N, generator=g)
        yp = torch.randint(0, V, (S, N), generator=g)
        if lens_mode == "none":
            lens = None
        elif lens_mode == "full":
            lens = torch.full((N,), S, dtype=torch.long)
        elif lens_mode == "short":
            lens = torch.randint(0, max(S, 1), (N,), generator=g)
        else:
            lens = torch.randint(0, S + 1, (N,), generator=g)
            lens[0] = S
        torch.manual_seed(100 + k)
        y_next, lp_next = F.random_walk_advance(lpt, lpp, yp, lens)
        # the token drawn for path n sits at its new position
        pos = lens if lens is not None else torch.full((N,), S, dtype=torch.long)
        y_t = y_next[pos, torch.arange(N)]
        name = f"walk{k}"
        params[name] = dict(has_lens=lens is not None, seed=100 + k)
        k += 1
        store[f"{name}.log_probs_t"] = lpt.numpy()
        store[f"{name}.log_probs_prev"] = lpp.numpy()
        store[f"{name}.y_prev"] = yp.numpy()
        if lens is not None:
            store[f"{name}.y_prev_lens"] = lens.numpy()
        store[f"{name}.y_t"] = y_t.numpy()
        store[f"{name}.y_next"] = y_next.numpy()
        store[f"{name}.log_probs_next"] = lp_next.numpy()
    store["params"] = np.array(json.dumps(params))
    np.savez_compressed(os.path.join(HERE, "decode.npz"), **store)
    print(f"decode.npz: {len(params)} cases")


if __name__ == "__main__":
    main()
