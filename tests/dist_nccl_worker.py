"""Worker of tests/test_dist_nccl.py: one process per GPU (torchrun), NCCL backend.
Checks, on every rank, the sharded bulk scoring of tensors and of corpora against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import parity_cases as PC  # noqa: E402
from b200lev import dist as D  # noqa: E402
from b200lev import scoring as S  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    try:
        # ---- tensors: 4-best groups never straddle ranks; totals are global on every rank -------
        rng = np.random.default_rng(0)  # the same data on every rank
        P = 4 * 3000
        ref = PC.random_tokens(rng, 31, P, 500, -1, -2, min_len=3)
        hyp = PC.random_tokens(rng, 31, P, 500, -1, -2, min_len=3)
        exp = O.error_rate(ref, hyp, eos=-1, norm=False)
        ref_lens = (ref == -1).argmax(0)
        lo, hi = D.shard_bounds(P, rank, world, group=4)
        assert lo % 4 == 0 and hi % 4 == 0
        er, totals = D.bulk_error_rate(torch.from_numpy(ref[:, lo:hi]).to(dev),
                                       torch.from_numpy(hyp[:, lo:hi]).to(dev), eos=-1)
        assert totals.is_cuda, "the all-reduce buffer must stay on the device (NCCL)"
        assert np.array_equal(er.cpu().numpy(), exp[lo:hi]), "per-pair errors differ from the oracle"
        assert totals.tolist() == [float(exp.sum()), float(ref_lens.sum()), float(P)], totals.tolist()
        # host tensors: same numbers, result on the host, totals on the device
        er_h, totals_h = D.bulk_error_rate(torch.from_numpy(ref[:, lo:hi]), torch.from_numpy(hyp[:, lo:hi]),
                                           eos=-1)
        assert er_h.device.type == "cpu" and np.array_equal(er_h.numpy(), exp[lo:hi])
        assert totals_h.tolist() == totals.tolist()
        # ---- corpora: utterance-sharded scoring --------------------------------------------------
        utts = [f"u{i:05d}" for i in range(2001)]
        rs = [rng.integers(0, 40, int(rng.integers(1, 25))) for _ in utts]
        hs = [rng.integers(0, 40, int(rng.integers(0, 25))) for _ in utts]
        rc, hc = S.TokenCorpus.from_sequences(utts, rs), S.TokenCorpus.from_sequences(utts, hs)
        clo, chi, errs, rl, ctot = D.score_corpora_sharded(rc, hc, quiet=True)
        want = np.array([len(r) if len(h) == 0 else float(np.asarray(O.edit_distance(r[:, None], h[:, None]))[0])
                         for r, h in zip(rs, hs)])
        assert np.array_equal(errs, want[clo:chi].astype(np.float32))
        assert ctot.tolist() == [float(want.sum()), float(sum(len(r) for r in rs)), float(len(utts))]
        print(f"rank {rank}/{world}: ok (pairs {lo}:{hi}, utterances {clo}:{chi})", flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
