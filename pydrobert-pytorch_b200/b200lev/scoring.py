"""Bulk scoring front end: token data directories -> one batched error-rate call.

The B200-side replacement for the body of ``compute-torch-token-data-dir-error-rates``
(reference ``command_line.py:1027-1147``) and for the directory reader behind it
(``_DirectoryDataset`` / ``_TranscriptDataSet``, ``command_line.py:394-466``; on-disk format
``_datasets.py:64-106``: one ``<prefix><utt><suffix>`` file per utterance holding a long
tensor of shape ``(L,)``, ``(L, 1)`` or ``(L, 3)`` = token / start / end).

The reference walks Python lists token by token (a ``defaultdict`` token->id map, one
``pad_sequence`` + ``error_rate`` per 100 utterances).  Here a corpus is two flat arrays
(tokens, offsets); ``--id2token`` / ``--replace`` / ``--ignore`` are applied once per
DISTINCT token and pushed through the corpus as a look-up table; the coded corpus goes to
the device flat (int16 when the vocabulary allows) and the padded ``(N, T)`` batches (eos
``-1``, padding ``-2``: codes are >= 0, so no token can collide with either) are built there
by ``lev_ragged.cu``; every batch is ONE ``error_rate`` call (the whole corpus when it fits
the cell budget, else length-sorted runs so that the padding stays small).  Outputs
(per-utterance lines, corpus rate) are formatted exactly as the reference prints them.
"""
import argparse
import os
import sys
import warnings
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Hashable, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np
import torch

from . import _host, _ops, config
from . import functional as F

__all__ = [
    "TokenCorpus",
    "load_token_data_dir",
    "parse_id2token",
    "align_utterances",
    "score_corpora",
    "compute_torch_token_data_dir_error_rates",
]

_EOS, _PAD = -1, -2
# padded tokens (ref + hyp) per error_rate call: 2^28 int32 = 1 GiB of host matrix
_CELL_BUDGET = 1 << 28


class TokenCorpus:
    """Utterances as ``tokens[offsets[n]:offsets[n + 1]]`` (int64), ids sorted as the
    reference's directory listing is (``command_line.py:399-403``)."""

    def __init__(self, utt_ids: Sequence[str], tokens: np.ndarray, offsets: np.ndarray, name: str = ""):
        self.utt_ids = list(utt_ids)
        self.tokens = np.ascontiguousarray(tokens, dtype=np.int64)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.name = name
        if self.offsets.shape != (len(self.utt_ids) + 1,) or (
            len(self.utt_ids) and int(self.offsets[-1]) != self.tokens.shape[0]
        ):
            raise ValueError("offsets do not describe tokens")

    def __len__(self) -> int:
        return len(self.utt_ids)

    @classmethod
    def from_sequences(cls, utt_ids: Sequence[str], seqs: Iterable[Sequence[int]], name: str = ""):
        arrs = [np.asarray(s, dtype=np.int64).reshape(-1) for s in seqs]
        lens = np.array([a.shape[0] for a in arrs], dtype=np.int64)
        off = np.zeros(len(arrs) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        flat = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.int64)
        return cls(utt_ids, flat, off, name)

    def select(self, keep: np.ndarray) -> "TokenCorpus":
        """The corpus restricted to the utterances ``keep`` (ascending indices)."""
        keep = np.asarray(keep, dtype=np.int64)
        lens = np.diff(self.offsets)[keep]
        off = np.zeros(keep.shape[0] + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        idx = _ragged_arange(self.offsets[:-1][keep], lens)
        return TokenCorpus([self.utt_ids[i] for i in keep], self.tokens[idx], off, self.name)


def _ragged_arange(starts: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """concat(arange(s, s + l) for s, l in zip(starts, lens)) without a Python loop."""
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    nz = lens > 0
    starts, lens = starts[nz], lens[nz]
    step = np.ones(total, dtype=np.int64)
    heads = np.zeros(lens.shape[0], dtype=np.int64)
    np.cumsum(lens[:-1], out=heads[1:])
    step[heads[0]] = starts[0]
    step[heads[1:]] = starts[1:] - (starts[:-1] + lens[:-1] - 1)
    return np.cumsum(step)


def _load_one(path: str) -> np.ndarray:
    tok = torch.load(path)
    if tok.dim() == 2:  # (L, 1) or (L, 3): the token id is column 0 (_parsing.py:884-891)
        tok = tok[:, 0] if tok.shape[1] else tok.reshape(-1)
    return tok.reshape(-1).to(torch.int64).numpy()


def load_token_data_dir(dir_: str, file_prefix: str = config.DEFT_FILE_PREFIX,
                        file_suffix: str = config.DEFT_FILE_SUFFIX, num_workers: int = 0) -> TokenCorpus:
    """Read every ``<prefix><utt><suffix>`` of ``dir_`` (timing columns dropped)."""
    fpl, fsl = len(file_prefix), len(file_suffix)
    utt_ids = sorted(x[fpl:len(x) - fsl] for x in os.listdir(dir_)
                     if x.startswith(file_prefix) and x.endswith(file_suffix))
    paths = [os.path.join(dir_, file_prefix + u + file_suffix) for u in utt_ids]
    if num_workers > 0:
        with ThreadPoolExecutor(num_workers) as ex:
            arrs = list(ex.map(_load_one, paths, chunksize=256))
    else:
        arrs = [_load_one(p) for p in paths]
    return TokenCorpus.from_sequences(utt_ids, arrs, dir_)


def parse_id2token(file, swap: bool = False) -> Dict[int, str]:
    """``<id> <token>`` lines (``<token> <id>`` with ``swap``); same checks, messages and
    last-one-wins behaviour as ``_parse_token2id`` (command_line.py:265-289) as the bulk
    scorer calls it (:998-999)."""
    ids, toks = dict(), dict()
    name = getattr(file, "name", "<id2token>")
    for line_no, line in enumerate(file):
        line = line.strip()
        if not line:
            continue
        ls = line.split()
        if len(ls) != 2 or not ls[int(swap)].lstrip("-").isdigit():
            raise ValueError(f"Cannot parse line {line_no + 1} of {name}")
        id_, tok = (int(ls[1]), ls[0]) if swap else (int(ls[0]), ls[1])
        for key, seen in ((tok, toks), (id_, ids)) if swap else ((id_, ids), (tok, toks)):
            if key in seen:
                warnings.warn(f'{name} line {line_no + 1}: "{key}" already exists. Mapping will be ambiguous')
        ids[id_] = tok
        toks[tok] = id_
    return ids


def align_utterances(ref: TokenCorpus, hyp: TokenCorpus, warn_missing: bool = False
                     ) -> Tuple[TokenCorpus, TokenCorpus]:
    """Drop (``warn_missing``) or reject utterances only one side has
    (command_line.py:1040-1071: same messages, same first offender)."""
    ri, hi, kr, kh = 0, 0, [], []
    R, H = ref.utt_ids, hyp.utt_ids
    while ri < len(R) or hi < len(H):
        if hi < len(H) and (ri == len(R) or H[hi] < R[ri]):
            tup, hi = (hyp.name, H[hi], ref.name), hi + 1
        elif ri < len(R) and (hi == len(H) or R[ri] < H[hi]):
            tup, ri = (ref.name, R[ri], hyp.name), ri + 1
        else:
            kr.append(ri)
            kh.append(hi)
            ri, hi = ri + 1, hi + 1
            continue
        msg = 'Directory "{}" contains utterance "{}" which directory "{}" does not contain'.format(*tup)
        if not warn_missing:
            raise ValueError(msg)
        warnings.warn(msg + ". Skipping")
    if len(kr) == len(R) and len(kh) == len(H):
        return ref, hyp
    return ref.select(np.array(kr, dtype=np.int64)), hyp.select(np.array(kh, dtype=np.int64))


def _check_known(c: TokenCorpus, id2token: Dict[int, str]) -> None:
    """The reference rejects an id without a token while it loads the directory
    (command_line.py:437-441): first utterance, first such id."""
    known = np.fromiter(id2token.keys(), np.int64, len(id2token))
    bad = np.flatnonzero(~np.isin(c.tokens, known))
    if bad.size:
        utt = int(np.searchsorted(c.offsets, int(bad[0]), side="right")) - 1
        raise ValueError(f"Utterance '{c.utt_ids[utt]}': ID '{int(c.tokens[bad[0]])}' could not be found "
                         "in id2token")


# ids spanning at most this many values are ranked through a dense table (two linear passes)
# instead of a sort
_DENSE_SPAN = 1 << 26


_CHUNK = 1 << 22  # tokens per host task: numpy releases the GIL inside these loops
_POOL = None


def _chunked(fn, *arrays):
    """fn over aligned slices of the arrays, on a few host threads when there is enough work."""
    global _POOL
    n = arrays[0].shape[0]
    if n <= _CHUNK:
        return [fn(*arrays)] if n else []
    if _POOL is None:
        _POOL = ThreadPoolExecutor(min(8, os.cpu_count() or 1))
    cuts = range(0, n, _CHUNK)
    return list(_POOL.map(lambda c: fn(*(a[c:c + _CHUNK] for a in arrays)), cuts))


def _recode(ref: TokenCorpus, hyp: TokenCorpus, id2token: Optional[Dict[int, str]],
            replace: Optional[Dict[Hashable, Hashable]], ignore: Optional[Set[Hashable]]):
    """Token ids -> ((ref, hyp) codes >= 0, (ref, hyp) keep masks or None, #codes).  Equal codes <=> equal tokens
    after ``id2token`` and ``replace`` -- all the edit distance sees of
    command_line.py:1081-1108's token2id map.  The map is evaluated once per DISTINCT id
    and applied as a look-up table; without any of the three options the ids only need
    shifting to be non-negative."""
    if id2token is not None:
        _check_known(ref, id2token)
        _check_known(hyp, id2token)
    sides = (ref.tokens, hyp.tokens)
    if ref.tokens.size + hyp.tokens.size == 0:
        z = np.zeros(0, dtype=np.int32)
        return (z, z), (None, None), 0
    ranges = [r for t in sides for r in _chunked(lambda a: (int(a.min()), int(a.max())), t)]
    lo, hi = min(r[0] for r in ranges), max(r[1] for r in ranges)
    span = hi - lo + 1
    if id2token is None and not replace and not ignore and span < (1 << 31):
        # (t - lo) evaluated in the narrow type: exact modulo 2^16 / 2^32, and the result fits
        dt = np.int16 if span < (1 << 15) else np.int32
        lo_n = np.int64(lo).astype(dt)  # wrapped, like the tokens the narrow subtraction reads
        outs = tuple(np.empty(t.shape, dtype=dt) for t in sides)
        for t, o in zip(sides, outs):
            _chunked(lambda a, b: np.subtract(a, lo_n, out=b, dtype=dt, casting="unsafe"), t, o)
        return outs, (None, None), span
    if span <= _DENSE_SPAN:
        present = np.zeros(span, dtype=bool)
        rel = tuple(t - lo for t in sides)
        for r in rel:
            present[r] = True
        uniq = np.flatnonzero(present)
        slot = np.zeros(span, dtype=np.int32)
        slot[uniq] = np.arange(uniq.shape[0], dtype=np.int32)
        inv, uniq = tuple(slot[r] for r in rel), uniq + lo
    else:
        uniq, both = np.unique(np.concatenate(sides), return_inverse=True)
        both = both.reshape(-1)
        inv = (both[:ref.tokens.size], both[ref.tokens.size:])
    code = np.zeros(uniq.shape[0], dtype=np.int32)
    keep = np.ones(uniq.shape[0], dtype=bool)
    codes: Dict[Hashable, int] = dict()
    replace = replace or dict()
    ignore = ignore or set()
    for k, id_ in enumerate(uniq.tolist()):
        tok = id2token[id_] if id2token is not None else id_
        tok = replace.get(tok, tok)
        if tok in ignore:
            keep[k] = False
        else:
            code[k] = codes.setdefault(tok, len(codes))
    keeps = (None, None) if keep.all() else tuple(keep[i] for i in inv)
    return tuple(code[i] for i in inv), keeps, len(codes)


def _filtered(c: TokenCorpus, code: np.ndarray, keep: Optional[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    if keep is None:
        return code, c.offsets
    csum = np.zeros(code.shape[0] + 1, dtype=np.int64)
    np.cumsum(keep, out=csum[1:])
    return code[keep], csum[c.offsets]


def _to_device(a: np.ndarray) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a))
    return _host.Placement(t).to_dev(t)  # raises without a CUDA device: there is no CPU path


def score_corpora(ref: TokenCorpus, hyp: TokenCorpus, id2token: Optional[Dict[int, str]] = None,
                  replace: Optional[Dict[Hashable, Hashable]] = None,
                  ignore: Optional[Set[Hashable]] = None,
                  costs: Sequence[float] = (config.DEFT_INS_COST, config.DEFT_DEL_COST, config.DEFT_SUB_COST),
                  quiet: bool = False, cell_budget: int = _CELL_BUDGET) -> Tuple[np.ndarray, np.ndarray]:
    """Errors (fp32 counts, as ``error_rate(norm=False)`` returns them) and reference
    lengths (after ``ignore``) per utterance of two ALIGNED corpora."""
    if ref.utt_ids != hyp.utt_ids:
        raise ValueError("corpora are not aligned (see align_utterances)")
    N = len(ref)
    code, keep, ncodes = _recode(ref, hyp, id2token, replace, ignore)
    rflat, roff = _filtered(ref, code[0], keep[0])
    hflat, hoff = _filtered(hyp, code[1], keep[1])
    rlen, hlen = np.diff(roff), np.diff(hoff)
    errors = np.zeros(N, dtype=np.float32)
    if N == 0:
        return errors, rlen
    # the corpus goes to the device flat (int16 codes when the vocabulary allows: a quarter of
    # the int64 bytes); the padded batches are built there (lev_ragged.cu)
    rdev, hdev = _to_device(rflat), _to_device(hflat)
    roff_d, hoff_d = _to_device(roff), _to_device(hoff)
    if (int(rlen.max()) + int(hlen.max()) + 2) * N <= cell_budget:
        order = None  # one call holds the corpus: rows in corpus order
    else:
        order = np.argsort(np.maximum(rlen, hlen), kind="stable")  # short ones together
    a = 0
    while a < N:
        # longest run of utterances (in `order`) whose two padded matrices fit the budget
        rl, hl = (rlen[a:], hlen[a:]) if order is None else (rlen[order[a:]], hlen[order[a:]])
        width = np.maximum.accumulate(rl + 1) + np.maximum.accumulate(hl + 1)
        fits = np.flatnonzero(width * np.arange(1, N - a + 1) <= cell_budget)
        n = int(fits[-1]) + 1 if fits.size else 1
        sel = None if order is None else _to_device(order[a:a + n])
        rm = _ops.ragged_to_padded(rdev, roff_d, sel, a, n, int(rl[:n].max()) + 1, _EOS, _PAD)
        hm = _ops.ragged_to_padded(hdev, hoff_d, sel, a, n, int(hl[:n].max()) + 1, _EOS, _PAD)
        er = F.error_rate(rm, hm, eos=_EOS, include_eos=False, norm=False, batch_first=True,
                          ins_cost=costs[0], del_cost=costs[1], sub_cost=costs[2], warn=not quiet)
        errors[slice(a, a + n) if order is None else order[a:a + n]] = er.cpu().numpy()
        a += n
    return errors, rlen


def _as_dir(val: str) -> str:
    if not os.path.isdir(val):
        raise argparse.ArgumentTypeError(f"'{val}' is not a directory")
    return val


def _as_nat(val: str) -> int:
    v = int(val)
    if v < 1:
        raise argparse.ArgumentTypeError(f"{val} is not a natural number")
    return v


def _typed(x: str, as_int: bool, fname: str):
    if not as_int:
        return x
    try:
        return int(x)
    except ValueError:
        raise ValueError(f'If --id2token is not set, all elements in "{fname}" must be integers')


def compute_torch_token_data_dir_error_rates(args: Optional[Sequence[str]] = None) -> Optional[int]:
    """Same command line and output as the reference's command (command_line.py:858-1147).

    ``--batch-size`` is accepted and ignored: batches are sized by the padded-cell budget."""
    parser = argparse.ArgumentParser(description=compute_torch_token_data_dir_error_rates.__doc__)
    parser.add_argument("dir", type=_as_dir)
    parser.add_argument("hyp", nargs="?", type=_as_dir, default=None)
    parser.add_argument("out", nargs="?", type=argparse.FileType("w"), default=sys.stdout)
    parser.add_argument("--id2token", type=argparse.FileType("r"), default=None)
    parser.add_argument("--replace", type=argparse.FileType("r"), default=None)
    parser.add_argument("--ignore", type=argparse.FileType("r"), default=None)
    parser.add_argument("--file-prefix", default=config.DEFT_FILE_PREFIX)
    parser.add_argument("--file-suffix", default=config.DEFT_FILE_SUFFIX)
    parser.add_argument("--swap", action="store_true", default=False)
    parser.add_argument("--warn-missing", action="store_true", default=False)
    parser.add_argument("--distances", action="store_true", default=False)
    parser.add_argument("--per-utt", action="store_true", default=False)
    parser.add_argument("--batch-size", type=_as_nat, default=100)
    parser.add_argument("--num-workers", type=int, default=0)  # threads lose to the GIL in torch.load
    parser.add_argument("--quiet", action="store_true", default=False)
    group = parser.add_mutually_exclusive_group()
    group.add_argument("--costs", nargs=3, type=float, metavar=("INS", "DEL", "SUB"),
                       default=(config.DEFT_INS_COST, config.DEFT_DEL_COST, config.DEFT_SUB_COST))
    group.add_argument("--nist-costs", action="store_true", default=False)
    try:
        options = parser.parse_args(args)
    except SystemExit as ex:
        return ex.code
    if options.nist_costs:
        options.costs = (3.0, 3.0, 4.0)
    if options.hyp:
        ref_dir, hyp_dir = options.dir, options.hyp
    else:
        ref_dir, hyp_dir = os.path.join(options.dir, "ref"), os.path.join(options.dir, "hyp")
    for d in (ref_dir, hyp_dir):
        if not os.path.isdir(d):
            print(f'"{d}" is not a directory', file=sys.stderr)
            return 1
    id2token = parse_id2token(options.id2token, options.swap) if options.id2token else None
    replace = dict()
    if options.replace:
        for line in options.replace:
            replaced, replacement = line.strip().split()
            replace[_typed(replaced, id2token is None, options.replace.name)] = _typed(
                replacement, id2token is None, options.replace.name)
    ignore = set()
    if options.ignore:
        ignore = {_typed(x, id2token is None, options.ignore.name)
                  for x in options.ignore.read().strip().split()}
    ref = load_token_data_dir(ref_dir, options.file_prefix, options.file_suffix, options.num_workers)
    hyp = load_token_data_dir(hyp_dir, options.file_prefix, options.file_suffix, options.num_workers)
    if id2token is not None:  # before the alignment, the reference directory first
        _check_known(ref, id2token)
        _check_known(hyp, id2token)
    ref, hyp = align_utterances(ref, hyp, options.warn_missing)
    errors, rlen = score_corpora(ref, hyp, id2token, replace, ignore, options.costs, options.quiet)
    errs = errors.astype(np.float64)  # er.item(): the fp32 count as a Python double
    if options.per_utt:
        lines: List[str] = []
        for utt_id, e, n in zip(ref.utt_ids, errs.tolist(), rlen.tolist()):
            lines.append("{} {}\n".format(utt_id, e / (1 if options.distances else n)))
        options.out.write("".join(lines))
    else:
        if not options.distances and (rlen == 0).any():
            # the reference divides per utterance before it looks at --per-utt (:1131-1133)
            raise ZeroDivisionError("float division by zero")
        tot, denom = float(errs.sum()), (len(ref) if options.distances else float(rlen.sum()))
        options.out.write("{}\n".format(tot / denom))
    options.out.flush()  # the reference leaves this to the file object's finaliser
    return None  # as the reference does on success (a console-script exit status of 0)
