#!/usr/bin/env python
"""Bulk-scoring front end on cfg4's shape (1 M utterances, T <= 31): where the time goes.

    python scripts/bench_scoring.py [--utts 1000000] [--files 20000]

Prints one JSON line per stage: the whole `score_corpora` call from flat arrays (recode +
padded matrices + pipelined error_rate), its parts, and `load_token_data_dir` on a
directory of `--files` utterance files (the reference's on-disk format; torch.load per
file is the floor there).
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pydrobert-pytorch_b200"))

import b200lev.scoring as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=1_000_000)
    ap.add_argument("--files", type=int, default=20_000)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    N, T, V = a.utts, 31, 10_000
    rl = rng.integers(1, T + 1, N)
    hl = np.clip(rl + rng.integers(-3, 4, N), 0, T)
    roff = np.concatenate([[0], np.cumsum(rl)])
    hoff = np.concatenate([[0], np.cumsum(hl)])
    rt = rng.integers(0, V, int(roff[-1]))
    ht = rng.integers(0, V, int(hoff[-1]))
    ids = [f"u{i:07d}" for i in range(N)]
    ref, hyp = S.TokenCorpus(ids, rt, roff, "ref"), S.TokenCorpus(ids, ht, hoff, "hyp")

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            t = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t)
        return best

    whole = timed(lambda: S.score_corpora(ref, hyp, quiet=True))
    print(json.dumps({"stage": "score_corpora (flat arrays -> per-utterance errors)", "utts": N,
                      "ms": whole * 1e3, "utts_per_s": N / whole}))
    t_recode = timed(lambda: S._recode(ref, hyp, None, None, None))
    code, keep, ncodes = S._recode(ref, hyp, None, None, None)
    t_h2d = timed(lambda: (S._to_device(code[0]), S._to_device(code[1]), S._to_device(ref.offsets),
                           S._to_device(hyp.offsets)))
    rd, hd, ro, ho = (S._to_device(code[0]), S._to_device(code[1]), S._to_device(ref.offsets),
                      S._to_device(hyp.offsets))

    def dev_ms(fn, reps=20):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    Tr, Th = int(rl.max()) + 1, int(hl.max()) + 1
    ms_pad = dev_ms(lambda: (S._ops.ragged_to_padded(rd, ro, None, 0, N, Tr, -1, -2),
                             S._ops.ragged_to_padded(hd, ho, None, 0, N, Th, -1, -2)))
    rm, hm = S._ops.ragged_to_padded(rd, ro, None, 0, N, Tr, -1, -2), S._ops.ragged_to_padded(hd, ho, None, 0, N, Th, -1, -2)
    ms_call = dev_ms(lambda: S.F.error_rate(rm, hm, eos=-1, norm=False, batch_first=True, warn=False))
    pad_bytes = (rm.numel() + hm.numel()) * 2 + (rd.numel() + hd.numel()) * 2 + 2 * 8 * N
    print(json.dumps({"stage": "parts", "recode_ms": t_recode * 1e3, "h2d_flat_ms": t_h2d * 1e3,
                      "ragged_to_padded_x2_ms": ms_pad, "ragged_GBps": pad_bytes / ms_pad / 1e6,
                      "error_rate_device_ms": ms_call, "h2d_bytes": int(rd.numel() + hd.numel()) * 2 + 16 * N}))
    if a.files > 0:
        with tempfile.TemporaryDirectory() as tmp:
            for i in range(a.files):
                torch.save(torch.from_numpy(rt[roff[i]:roff[i + 1]]), os.path.join(tmp, ids[i] + ".pt"))
            for workers in (0, 8):
                t = time.perf_counter()
                c = S.load_token_data_dir(tmp, num_workers=workers)
                dt = time.perf_counter() - t
                assert len(c) == a.files
                print(json.dumps({"stage": "load_token_data_dir", "files": a.files, "num_workers": workers,
                                  "ms": dt * 1e3, "files_per_s": a.files / dt}))


if __name__ == "__main__":
    main()
