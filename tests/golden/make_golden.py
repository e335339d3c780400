#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (it imports /root/reference/src, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

The outputs (``*.npz``, a few hundred KB) are committed; the tests never import the
reference.  Every expected value in the fixtures was produced by
``pydrobert.torch.functional`` / ``pydrobert.torch._string`` of the reference checkout.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

REF_SRC = os.environ.get("B200LEV_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF_SRC)

import pydrobert.torch.functional as F  # noqa: E402
from pydrobert.torch import _string as SM  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TESTS = os.path.join(os.path.dirname(REF_SRC), "tests")

warnings.simplefilter("ignore")

COSTS = [
    (1.0, 1.0, 1.0),
    (2.0, 2.0, 2.0),
    (1.0, 2.0, 3.0),
    (3.0, 3.0, 4.0),
    (0.5, 1.0, 1.0),
    (2.0, 0.5, 1.0),
    (0.25, 0.25, 0.25),
    (1.0, 1.0, 2.0),
    (0.0, 1.0, 1.0),
]
FLOAT_COSTS = [(0.7, 1.1, 1.3), (0.3, 0.3, 0.3), (1.1, 0.9, 1.7)]


def make_tokens(g, T, N, V, eos, pad, min_len=0, no_eos_frac=0.0):
    """(T, N) int64: random tokens in [1, V), one eos at a random position, tail = pad."""
    tok = torch.randint(1, max(V, 2), (T, N), generator=g)
    for n in range(N):
        if T == 0:
            continue
        if float(torch.rand((), generator=g)) < no_eos_frac:
            continue
        pos = int(torch.randint(min(min_len, T - 1), T, (), generator=g))
        tok[pos, n] = eos
        tok[pos + 1 :, n] = pad
    return tok


def run_case(store, name, ref, hyp, p):
    """Evaluate every DP-backed public function of the reference on (ref, hyp)."""
    kw = dict(eos=p["eos"], include_eos=p["include_eos"], batch_first=p["batch_first"],
              ins_cost=p["ins"], del_cost=p["del"], sub_cost=p["sub"], warn=False)
    r, h = (ref.t().contiguous(), hyp.t().contiguous()) if p["batch_first"] else (ref, hyp)
    store[f"{name}.ref"] = r.numpy()
    store[f"{name}.hyp"] = h.numpy()
    def attempt(key, fn):
        # a few degenerate shapes (zero-length dims) make the reference itself raise
        # (e.g. SM:285 with H'=0, SM:275 with R=0); those outputs are simply not pinned
        try:
            store[f"{name}.{key}"] = fn().numpy()
        except (IndexError, RuntimeError) as e:
            print(f"  {name}.{key}: reference raised {type(e).__name__}; not pinned")

    attempt("error_rate", lambda: F.error_rate(r, h, norm=p["norm"], **kw))
    attempt("edit_distance", lambda: F.edit_distance(r, h, norm=p["norm"], **kw))
    pk = dict(norm=p["norm"], padding=p["padding"], exclude_last=p["exclude_last"], **kw)
    attempt("prefix_error_rates", lambda: F.prefix_error_rates(r, h, **pk))
    attempt("prefix_edit_distances", lambda: F.prefix_edit_distances(r, h, **pk))
    if p.get("mask", True):
        attempt("mask", lambda: SM._string_matching(
            r, h, p["eos"], p["include_eos"], p["batch_first"], p["ins"], p["del"],
            p["sub"], False, return_mask=True, exclude_last=p["exclude_last"]))
        attempt("optimal_completion", lambda: F.optimal_completion(
            r, h, padding=p["padding"], exclude_last=p["exclude_last"], **kw))


def gen_string_cases():
    g = torch.Generator().manual_seed(20261017)
    store, params = {}, {}
    idx = 0
    # small random grid over every flag
    for costs in COSTS + FLOAT_COSTS:
        for trial in range(6):
            R = int(torch.randint(0, 13, (), generator=g))
            H = int(torch.randint(0, 13, (), generator=g))
            N = int(torch.randint(1, 7, (), generator=g))
            V = int(torch.randint(2, 6, (), generator=g))
            use_eos = trial != 0
            eos = int(torch.randint(-1, 2, (), generator=g)) if use_eos else None
            if use_eos:  # the reference's _lens_from_eos rejects zero-length dims (SM:142)
                R, H = max(R, 1), max(H, 1)
            p = dict(
                eos=eos,
                include_eos=bool(torch.randint(0, 2, (), generator=g)),
                batch_first=bool(torch.randint(0, 2, (), generator=g)),
                norm=bool(torch.randint(0, 2, (), generator=g)),
                exclude_last=bool(torch.randint(0, 2, (), generator=g)),
                padding=int(torch.randint(-5, 3, (), generator=g)),
                ins=costs[0], **{"del": costs[1]}, sub=costs[2],
                # the mask uses exact float equality (SM:334): pin it only where the
                # arithmetic is exact (integer / dyadic costs)
                mask=costs not in FLOAT_COSTS,
            )
            e = 0 if eos is None else eos
            ref = make_tokens(g, R, N, V, e, -7 if e != -7 else -8, no_eos_frac=0.2)
            hyp = make_tokens(g, H, N, V, e, -9, no_eos_frac=0.2)
            if eos is None:  # no eos processing: plain random tokens
                ref = torch.randint(0, V, (R, N), generator=g)
                hyp = torch.randint(0, V, (H, N), generator=g)
            name = f"s{idx:03d}"
            run_case(store, name, ref, hyp, p)
            params[name] = p
            idx += 1
    # medium cases shaped like reduced BASELINE configs
    med = [
        ("cfg1", 51, 51, 32, 30, 0, dict(include_eos=False, norm=True), (1.0, 1.0, 1.0)),
        ("cfg2r", 41, 41, 24, 10000, 0, dict(include_eos=True, norm=True), (1.0, 1.0, 1.0)),
        ("cfg3r", 70, 75, 10, 32, 0, dict(include_eos=True, norm=False), (1.0, 1.0, 1.0)),
        ("cfg4r", 31, 31, 64, 10000, -1, dict(include_eos=False, norm=False), (1.0, 1.0, 1.0)),
        ("cfg5r", 150, 140, 6, 64, 0, dict(include_eos=True, norm=False), (3.0, 3.0, 4.0)),
        ("cfg5f", 90, 100, 4, 64, 0, dict(include_eos=True, norm=False), (0.7, 1.1, 1.3)),
        ("wide", 130, 37, 5, 8, 0, dict(include_eos=True, norm=True), (1.0, 2.0, 3.0)),
        ("tall", 33, 160, 5, 8, 0, dict(include_eos=False, norm=True), (2.0, 1.0, 1.0)),
    ]
    for name, R, H, N, V, eos, flags, costs in med:
        p = dict(eos=eos, batch_first=False, exclude_last=False, padding=-100,
                 ins=costs[0], **{"del": costs[1]}, sub=costs[2],
                 mask=costs != (0.7, 1.1, 1.3), **flags)
        ref = make_tokens(g, R, N, V, eos, -2, min_len=R // 3)
        hyp = make_tokens(g, H, N, V, eos, -2, min_len=H // 3)
        run_case(store, name, ref, hyp, p)
        params[name] = p
    store["params"] = np.array(json.dumps(params))
    np.savez_compressed(os.path.join(HERE, "string_matching.npz"), **store)
    print("string_matching.npz:", len(params), "cases")


def gen_loss_cases():
    g = torch.Generator().manual_seed(777)
    store, params = {}, {}
    # OCD loss
    idx = 0
    for reduction in ("mean", "sum", "none"):
        for batch_first in (False, True):
            for include_eos in (True, False):
                for use_w in (False, True):
                    H, R, N, V = 9, 11, 5, 7
                    eos = 0
                    ref = make_tokens(g, R, N, V, eos, V - 1, min_len=1)
                    hyp = make_tokens(g, H, N, V, eos, V - 1, min_len=0)
                    logits = torch.randn(H, N, V, generator=g)
                    if batch_first:
                        ref, hyp = ref.t().contiguous(), hyp.t().contiguous()
                        logits = logits.transpose(0, 1).contiguous()
                    w = (torch.rand(V, generator=g) + 0.5) if use_w else None
                    logits.requires_grad_(True)
                    loss = F.hard_optimal_completion_distillation_loss(
                        logits, ref, hyp, eos=eos, include_eos=include_eos,
                        batch_first=batch_first, weight=w, reduction=reduction,
                        ignore_index=-2, warn=False)
                    go = torch.randn(loss.shape, generator=g) if reduction == "none" else None
                    (grad,) = torch.autograd.grad(
                        [loss if go is None else (loss * go).sum()], [logits])
                    name = f"ocd{idx:02d}"
                    store[f"{name}.ref"] = ref.numpy()
                    store[f"{name}.hyp"] = hyp.numpy()
                    store[f"{name}.logits"] = logits.detach().numpy()
                    if w is not None:
                        store[f"{name}.weight"] = w.numpy()
                    if go is not None:
                        store[f"{name}.grad_output"] = go.numpy()
                    store[f"{name}.loss"] = loss.detach().numpy()
                    store[f"{name}.grad"] = grad.numpy()
                    params[name] = dict(eos=eos, include_eos=include_eos,
                                        batch_first=batch_first, reduction=reduction,
                                        ignore_index=-2, weight=use_w)
                    idx += 1
    # MWER loss
    idx = 0
    for reduction in ("mean", "sum", "none"):
        for batch_first in (False, True):
            for sub_avg in (True, False):
                for ref3d in (False, True):
                    for costs in ((1.0, 1.0, 1.0), (3.0, 3.0, 4.0)):
                        H, R, N, M, V = 8, 10, 4, 3, 6
                        eos = 0
                        hyp = make_tokens(g, H, N * M, V, eos, -1).view(H, N, M)
                        if ref3d:
                            ref = make_tokens(g, R, N * M, V, eos, -1).view(R, N, M)
                        else:
                            ref = make_tokens(g, R, N, V, eos, -1)
                        lp = torch.randn(N, M, generator=g)
                        if batch_first:
                            hyp = hyp.permute(1, 2, 0).contiguous()
                            ref = (ref.permute(1, 2, 0) if ref3d else ref.t()).contiguous()
                        lp.requires_grad_(True)
                        loss = F.minimum_error_rate_loss(
                            lp, ref, hyp, eos=eos, include_eos=True, sub_avg=sub_avg,
                            batch_first=batch_first, norm=True, ins_cost=costs[0],
                            del_cost=costs[1], sub_cost=costs[2], reduction=reduction,
                            warn=False)
                        go = torch.randn(loss.shape, generator=g) if reduction == "none" else None
                        (grad,) = torch.autograd.grad(
                            [loss if go is None else (loss * go).sum()], [lp])
                        name = f"mwer{idx:02d}"
                        store[f"{name}.ref"] = ref.numpy()
                        store[f"{name}.hyp"] = hyp.numpy()
                        store[f"{name}.log_probs"] = lp.detach().numpy()
                        if go is not None:
                            store[f"{name}.grad_output"] = go.numpy()
                        store[f"{name}.loss"] = loss.detach().numpy()
                        store[f"{name}.grad"] = grad.numpy()
                        params[name] = dict(eos=eos, include_eos=True, sub_avg=sub_avg,
                                            batch_first=batch_first, norm=True,
                                            ins=costs[0], **{"del": costs[1]}, sub=costs[2],
                                            reduction=reduction)
                        idx += 1
    store["params"] = np.array(json.dumps(params))
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **store)
    print("losses.npz:", len(params), "cases")


def gen_sclite():
    """tests/sclite/* of the reference -> token tensors + expected per-utt / total.

    Follows command_line.py:1074-1147 (eos=-1, padding=-2, NIST costs 3/3/4,
    error_rate(norm=False), per-utt = errs / len(ref))."""
    d = os.path.join(REF_TESTS, "sclite")
    tok2id = {}
    with open(os.path.join(d, "token2id.txt")) as f:
        for line in f:
            t, i = line.split()
            tok2id[t] = int(i)

    def read_trn(fn):
        out = {}
        with open(fn) as f:
            for line in f:
                line = line.strip()
                if not line:
                    continue
                words, utt = line.rsplit("(", 1)
                out[utt.rstrip(")")] = [tok2id[w] for w in words.split()]
        return out

    refs, hyps = read_trn(os.path.join(d, "ref.trn")), read_trn(os.path.join(d, "hyp.trn"))
    utts = sorted(refs)
    eos, padding = -1, -2
    ref = torch.nn.utils.rnn.pad_sequence(
        [torch.tensor(refs[u] + [eos]) for u in utts], padding_value=padding)
    hyp = torch.nn.utils.rnn.pad_sequence(
        [torch.tensor(hyps[u] + [eos]) for u in utts], padding_value=padding)
    ers = F.error_rate(ref, hyp, eos=eos, include_eos=False, ins_cost=3.0, del_cost=3.0,
                       sub_cost=4.0, norm=False, warn=False)
    per_utt = {}
    with open(os.path.join(d, "per_utt.txt")) as f:
        for line in f:
            u, v = line.split()
            per_utt[u] = float(v)
    with open(os.path.join(d, "total.txt")) as f:
        total = float(f.read().strip())
    ref_lens = np.array([len(refs[u]) for u in utts], np.int64)
    # the fixture is only worth committing if the reference reproduces sclite here
    for k, u in enumerate(utts):
        assert "{:.03f}".format(ers[k].item() / ref_lens[k]) == "{:.03f}".format(per_utt[u]), u
    assert "{:.03f}".format(ers.sum().item() / ref_lens.sum()) == "{:.03f}".format(total)
    np.savez_compressed(
        os.path.join(HERE, "sclite.npz"),
        ref=ref.numpy(), hyp=hyp.numpy(), ref_lens=ref_lens,
        errs=ers.numpy(), per_utt=np.array([per_utt[u] for u in utts]),
        total=np.array(total),
    )
    print("sclite.npz:", len(utts), "utterances, total", total)


def gen_seqlp_cases():
    """sequence_log_probs, tensor path (_decoding.py:1516-1548): outputs and gradients of the
    reference for every float dtype, several step axes, with and without eos, padding tokens
    outside [0, V), plus the reference's own known-answer test (tests/test_decoding.py:844-876
    builds its expectation the same way: masked gather of a log_softmax)."""
    g = np.random.default_rng(20240917)
    store, params = {}, {}
    shapes = [((6,), 0), ((5, 3), 0), ((4, 7), 1), ((3, 9, 4), 1), ((2, 5, 3, 2), -2), ((1, 1), 0),
              ((33, 5), 0)]
    k = 0
    for shape, dim in shapes:
        for V in (7, 40):
            for eos in (None, 0, 3):
                for dt in ("float32", "float64", "bfloat16", "float16"):
                    if (k % 4) and dt in ("float64", "float16"):  # thin the grid a little
                        k += 1
                        continue
                    k += 1
                    name = f"q{len(params)}"
                    hyp = g.integers(-2, V + 2, size=shape)
                    tdt = getattr(torch, dt)
                    logits = torch.tensor(g.standard_normal(shape + (V,)) * 3.0).to(tdt)
                    logits.requires_grad_(True)
                    out = F.sequence_log_probs(logits, torch.tensor(hyp), dim, eos)
                    go = torch.tensor(g.standard_normal(tuple(out.shape))).to(tdt)
                    (out * go).sum().backward()
                    # (every dtype but float64 is exactly representable in float32)
                    sdt = torch.float64 if dt == "float64" else torch.float32
                    store[name + ".logits"] = logits.detach().to(sdt).numpy()
                    store[name + ".hyp"] = hyp.astype(np.int16)
                    store[name + ".out"] = out.detach().to(sdt).numpy()
                    store[name + ".grad_out"] = go.to(sdt).numpy()
                    store[name + ".grad"] = logits.grad.to(sdt).numpy()
                    params[name] = dict(dim=dim, eos=eos, dtype=dt, V=V)
    store["params"] = np.array(json.dumps(params))
    np.savez_compressed(os.path.join(HERE, "seqlp.npz"), **store)
    print("seqlp.npz:", len(params), "cases")


def gen_ctc_cases():
    """ctc_greedy_search (_decoding.py:507-560): max_, paths (whole tensor: positions past
    out_lens keep the raw arg max) and out_lens of the reference; small class counts so that
    blanks and repeats are frequent.  No gradients: the reference's backward raises (its
    in-place masked_scatter_ invalidates the arg max saved for max's backward)."""
    g = np.random.default_rng(20240918)
    store, params = {}, {}
    for (T, N, V) in ((1, 1, 2), (7, 3, 3), (20, 5, 4), (45, 4, 6), (70, 2, 3), (33, 6, 40), (12, 3, 300)):
        for batch_first in (False, True):
            for lens_kind in ("none", "ragged"):
                for is_probs in (False, True):
                    for blank_idx in (-1, 0):
                        for dt in ("float32", "float64"):
                            if dt == "float64" and (len(params) % 3):
                                continue
                            if T * N * V > 1000 and (lens_kind == "none" or blank_idx == 0 or dt == "float64"):
                                continue  # the wide-row cases only in a few combinations
                            name = f"c{len(params)}"
                            tdt = getattr(torch, dt)
                            # values exactly representable in float32: the fixture stores float32
                            x = torch.tensor(g.standard_normal((T, N, V)) * 2.0).float().to(tdt)
                            if is_probs:
                                x = x.softmax(2).float().to(tdt)
                            if batch_first:
                                x = x.transpose(0, 1).contiguous()
                            lens = None if lens_kind == "none" else torch.tensor(g.integers(0, T + 3, N))
                            max_, paths, out_lens = F.ctc_greedy_search(x, lens, blank_idx, batch_first, is_probs)
                            store[name + ".logits"] = x.float().numpy()
                            if lens is not None:
                                store[name + ".in_lens"] = lens.numpy().astype(np.int16)
                            store[name + ".max"] = max_.numpy()
                            store[name + ".paths"] = paths.numpy().astype(np.int16)
                            store[name + ".out_lens"] = out_lens.numpy().astype(np.int16)
                            params[name] = dict(batch_first=batch_first, is_probs=is_probs, blank_idx=blank_idx,
                                                dtype=dt, lens=lens is not None)
    store["params"] = np.array(json.dumps(params))
    np.savez_compressed(os.path.join(HERE, "ctc.npz"), **store)
    print("ctc.npz:", len(params), "cases")


if __name__ == "__main__":
    which = sys.argv[1:] or ["string", "loss", "sclite", "seqlp", "ctc"]
    if "string" in which:
        gen_string_cases()
    if "loss" in which:
        gen_loss_cases()
    if "sclite" in which:
        gen_sclite()
    if "seqlp" in which:
        gen_seqlp_cases()
    if "ctc" in which:
        gen_ctc_cases()
