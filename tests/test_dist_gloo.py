"""Multi-process host logic of the pair-sharded bulk scoring, world_size 2,
gloo backend on CPU.  Kernels run on the SIMT emulator (test infrastructure); what is under
test is the sharding (n-best groups never split) and the single all-reduce of
[sum(err), sum(ref_len), #pairs]."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_cases as PC
from oracle import oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ref, hyp, q):
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    from emu_backend import emulated_kernels

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with emulated_kernels():
            from b200lev import dist as D

            lo, hi = D.shard_bounds(hyp.shape[1], rank, world, group=4)
            er, totals = D.bulk_error_rate(torch.from_numpy(ref[:, lo:hi]),
                                           torch.from_numpy(hyp[:, lo:hi]), eos=-1)
            q.put((rank, lo, hi, er.numpy(), totals.numpy()))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_respect_groups():
    from b200lev.dist import shard_bounds

    for n, w, g in ((64, 2, 8), (72, 4, 8), (40, 3, 4), (8, 8, 1), (16, 5, 2)):
        edges = [shard_bounds(n, r, w, g) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for (a, b), (c, d) in zip(edges, edges[1:]):
            assert b == c
        assert all(a % g == 0 and b % g == 0 for a, b in edges)
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 2, 4)


def test_two_rank_bulk_error_rate():
    rng = np.random.default_rng(0)
    P = 40
    ref = PC.random_tokens(rng, 12, P, 9, -1, -2, min_len=3)
    hyp = PC.random_tokens(rng, 13, P, 9, -1, -2, min_len=3)
    exp = O.error_rate(ref, hyp, eos=-1, norm=False)
    ref_lens = (ref == -1).argmax(0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ref, hyp, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lo, hi, er, totals in got:
        assert np.array_equal(er, exp[lo:hi])
        # every rank holds the GLOBAL totals after the all-reduce
        assert totals.tolist() == [float(exp.sum()), float(ref_lens.sum()), float(P)]
