#!/usr/bin/env python
"""Secondary measurements: the five BASELINE.json configs at their literal shapes (public
API, device-resident inputs, CUDA events, >= 3 warm-ups).  One JSON line per config; the
headline contract lives in bench.py."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import b200lev.functional as F
from b200lev import dist as D


def seqs(rng, T, n, V, lo, hi, eos, pad):
    tok = rng.integers(1, V, size=(T, n), dtype=np.int64)
    lens = rng.integers(lo, hi + 1, size=n)
    pos = np.arange(T)[:, None]
    tok[pos == (lens - 1)[None, :]] = eos
    tok[pos > (lens - 1)[None, :]] = pad
    return tok, lens


def timed(fn, reps, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()  # > L2: evict the (small) inputs between iterations
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


def main():
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []

    # cfg1: error_rate, batch 32, T~50, V=30
    r, rl = seqs(rng, 51, 32, 30, 25, 50, 0, 0)
    h, hl = seqs(rng, 51, 32, 30, 25, 50, 0, 0)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
    ms = timed(lambda: F.error_rate(tr, th, eos=0, warn=False), 50, flush)
    cells = int(((rl - 1) * (hl - 1)).sum())
    out.append(dict(cfg=1, call="error_rate", pairs=32, ms=ms, gcups=cells / ms / 1e6))

    # cfg2: MWER fwd+bwd and prefix_error_rates, 64 x 8-best, T=100, V=10k, bf16 log-probs
    r, rl = seqs(rng, 101, 64, 10000, 50, 100, 0, 0)
    h, hl = seqs(rng, 101, 512, 10000, 50, 100, 0, 0)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h.reshape(101, 64, 8)).to(dev)
    lp = torch.randn(64, 8, device=dev, dtype=torch.bfloat16, requires_grad=True)

    def mwer():
        loss = F.minimum_error_rate_loss(lp, tr, th, eos=0, warn=False)
        loss.backward()

    ms = timed(mwer, 50, flush)
    cells = int((np.repeat(rl, 8) * hl).sum())
    out.append(dict(cfg=2, call="minimum_error_rate_loss fwd+bwd", pairs=512, ms=ms,
                    gcups=cells / ms / 1e6, hyps_per_s=512 / ms * 1e3))
    # "next #1": sequence_log_probs on the same batch, bf16 logits with vocab 10 k (1.02 GB)
    lg = (torch.randn(101, 512, 10000, device=dev) * 2).to(torch.bfloat16).requires_grad_(True)
    h512 = torch.from_numpy(h).to(dev)
    ms = timed(lambda: F.sequence_log_probs(lg.detach(), h512, 0, eos=0), 20)
    rows = int(np.minimum(hl, 101).sum())  # rows that count (the others are never read)
    out.append(dict(cfg=2, call="sequence_log_probs fwd (bf16, V=10k)", pairs=512, ms=ms,
                    gbs_all_rows=lg.numel() * 2 / ms / 1e6, gbs_rows_read=rows * 10000 * 2 / ms / 1e6))

    def slp_torch():  # the same sum with stock torch ops (what the reference runs on a GPU)
        lp = torch.log_softmax(lg.detach(), -1)
        is_eos = h512 == 0
        bad = (is_eos.cumsum(0) - is_eos.long()) > 0
        return lp.gather(-1, h512.unsqueeze(-1)).squeeze(-1).masked_fill(bad, 0.0).sum(0)

    ms_t = timed(slp_torch, 10)
    out[-1]["stock_torch_ops_ms"] = ms_t

    def slp():
        o = F.sequence_log_probs(lg, h512, 0, eos=0)
        o.sum().backward()
        lg.grad = None

    ms = timed(slp, 10)
    out.append(dict(cfg=2, call="sequence_log_probs fwd+bwd (bf16, V=10k)", pairs=512, ms=ms,
                    gbs_3_passes=3 * lg.numel() * 2 / ms / 1e6))
    # "next #4": ctc_greedy_search on the same logits (one read: 1.01 GB)
    ms = timed(lambda: F.ctc_greedy_search(lg.detach(), None, 0), 20)
    out.append(dict(cfg=2, call="ctc_greedy_search fwd (bf16, V=10k)", pairs=512, ms=ms,
                    gbs=lg.numel() * 2 / ms / 1e6))

    def ctc_torch():  # the reference's op sequence (_decoding.py:524-555) on the device
        RF_lp = lg.detach().log_softmax(2).transpose(0, 1)
        mx, am = RF_lp.max(2)
        keep = am != 0
        keep = torch.cat([keep[:, :1], keep[:, 1:] & (am[:, 1:] != am[:, :-1])], 1)
        ol = keep.long().sum(1)
        data = am.masked_select(keep)
        olm = torch.arange(am.size(1), device=dev).unsqueeze(0) < ol.unsqueeze(1)
        return mx.sum(1), am.masked_scatter_(olm, data).t(), ol

    out[-1]["stock_torch_ops_ms"] = timed(ctc_torch, 10)
    del lg
    trx = torch.from_numpy(np.repeat(r, 8, axis=1)).to(dev)
    th2 = torch.from_numpy(h).to(dev)
    ms = timed(lambda: F.prefix_error_rates(trx, th2, eos=0, warn=False), 50, flush)
    out.append(dict(cfg=2, call="prefix_error_rates", pairs=512, ms=ms, gcups=cells / ms / 1e6))

    # cfg3: optimal_completion + OCD loss fwd+bwd, 128 x T=200, V=32
    r, rl = seqs(rng, 201, 128, 32, 100, 200, 0, 0)
    h, hl = seqs(rng, 201, 128, 32, 100, 200, 0, 0)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
    logits = torch.randn(201, 128, 32, device=dev, requires_grad=True)
    ms = timed(lambda: F.optimal_completion(tr, th, eos=0, warn=False), 20, flush)
    cells = int((rl * hl).sum())
    out.append(dict(cfg=3, call="optimal_completion", pairs=128, ms=ms, gcups=cells / ms / 1e6))

    def ocd():
        loss = F.hard_optimal_completion_distillation_loss(logits, tr, th, eos=0, warn=False)
        loss.backward()

    ms = timed(ocd, 20, flush)
    out.append(dict(cfg=3, call="hard_optimal_completion_distillation_loss fwd+bwd", pairs=128,
                    ms=ms, gcups=cells / ms / 1e6))

    # cfg4: bulk WER, 1M pairs, T~30 (inputs 496 MB > L2)
    P = 1_000_000
    r, rl = seqs(rng, 31, P, 10000, 10, 30, -1, -2)
    h, hl = seqs(rng, 31, P, 10000, 10, 30, -1, -2)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
    res = {}

    def bulk():
        res["er"], res["acc"] = D.bulk_error_rate(tr, th, eos=-1)

    ms = timed(bulk, 10)
    cells = int(((rl - 1).astype(np.int64) * (hl - 1)).sum())
    acc = res["acc"].tolist()
    out.append(dict(cfg=4, call="bulk_error_rate (error_rate + device sums)", pairs=P, ms=ms,
                    gcups=cells / ms / 1e6, pairs_per_s=P / ms * 1e3, wer=acc[0] / acc[1]))
    del tr, th

    # cfg5: prefix_edit_distances / prefix_error_rates, 256 x T=2000, NIST costs 3/3/4
    r, rl = seqs(rng, 2001, 256, 64, 200, 2000, 0, 0)
    h, hl = seqs(rng, 2001, 256, 64, 200, 2000, 0, 0)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
    cells = int((rl.astype(np.int64) * hl).sum())
    for name in ("prefix_edit_distances", "prefix_error_rates"):
        fn = getattr(F, name)
        ms = timed(lambda: fn(tr, th, eos=0, ins_cost=3, del_cost=3, sub_cost=4, warn=False), 5, flush)
        out.append(dict(cfg=5, call=name + " (3,3,4)", pairs=256, ms=ms, gcups=cells / ms / 1e6))
    ms = timed(lambda: F.prefix_edit_distances(tr, th, eos=0, ins_cost=0.7, del_cost=1.1, sub_cost=1.3), 3, flush)
    out.append(dict(cfg=5, call="prefix_edit_distances (0.7,1.1,1.3) fp32", pairs=256, ms=ms,
                    gcups=cells / ms / 1e6))
    del tr, th
    # saturated variants (SURVEY 8d iii): the same per-pair shapes with the batch replicated
    # until the chip is full (cfg2's is bench.py itself)
    r, rl = seqs(rng, 51, 65536, 30, 25, 50, 0, 0)
    h, hl = seqs(rng, 51, 65536, 30, 25, 50, 0, 0)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
    ms = timed(lambda: F.error_rate(tr, th, eos=0, warn=False), 20, flush)
    cells = int(((rl - 1).astype(np.int64) * (hl - 1)).sum())
    out.append(dict(cfg=1, call="error_rate, saturated", pairs=65536, ms=ms, gcups=cells / ms / 1e6,
                    hyps_per_s=65536 / ms * 1e3))
    r, rl = seqs(rng, 201, 8192, 32, 100, 200, 0, 0)
    h, hl = seqs(rng, 201, 8192, 32, 100, 200, 0, 0)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
    ms = timed(lambda: F.optimal_completion(tr, th, eos=0, warn=False), 5, flush)
    cells = int((rl.astype(np.int64) * hl).sum())
    out.append(dict(cfg=3, call="optimal_completion, saturated", pairs=8192, ms=ms, gcups=cells / ms / 1e6,
                    hyps_per_s=8192 / ms * 1e3))
    r, rl = seqs(rng, 2001, 1184, 64, 200, 2000, 0, 0)
    h, hl = seqs(rng, 2001, 1184, 64, 200, 2000, 0, 0)
    tr, th = torch.from_numpy(r).to(dev), torch.from_numpy(h).to(dev)
    cells = int((rl.astype(np.int64) * hl).sum())
    ms = timed(lambda: F.prefix_edit_distances(tr, th, eos=0, ins_cost=3, del_cost=3, sub_cost=4, warn=False), 3,
               flush)
    out.append(dict(cfg=5, call="prefix_edit_distances (3,3,4), saturated", pairs=1184, ms=ms,
                    gcups=cells / ms / 1e6))
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
