"""The eight functionals of the reference's string-matching family, same signatures.

Mirrors ``pydrobert.torch.functional`` (functional.py:49-58 of the reference) for the
names backed by ``pydrobert/torch/_string.py`` ("SM" below): argument meaning,
defaults, shapes, dtypes, error messages and the three data-dependent warnings are
the reference's; the numbers come from the sm_100a kernels behind ``torch.ops.b200lev``.

Host tensors: the reference runs wherever its inputs live.  Here a CPU tensor is
copied to the current CUDA device (asynchronously when it is pinned), the kernels run
there and the result is copied back -- the ``e2e`` path of ``bench.py``.  Large host
batches of the final / prefix modes are cut into blocks of the batch axis that flow
through three streams (copy in, kernels, copy out), so that the PCIe transfers of both
directions and the kernels overlap.  Without a CUDA device the call raises; there is
no CPU implementation.
"""
from __future__ import annotations

import warnings
from typing import Optional, Tuple

import torch

from . import _abi, _ops, config

__all__ = [
    "edit_distance",
    "error_rate",
    "fill_after_eos",
    "hard_optimal_completion_distillation_loss",
    "minimum_error_rate_loss",
    "optimal_completion",
    "prefix_edit_distances",
    "prefix_error_rates",
    "sequence_log_probs",
    "ctc_greedy_search",
]


def _offload(*tensors):
    """Move host tensors to the current CUDA device; returns (tensors, back) where
    ``back`` maps a result to the device the caller's inputs were on."""
    first = next(t for t in tensors if t is not None)
    if first.device.type == "cuda" or _abi.EMULATED:
        return tensors, (lambda x: x)
    if not torch.cuda.is_available():
        raise _abi.B200LevError(
            "b200lev needs a CUDA device: the string-matching kernels have no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    moved = tuple(None if t is None else t.to(dev, non_blocking=True) for t in tensors)

    def back(x):
        # D2H into page-locked memory (torch's caching host allocator recycles the block):
        # one DMA, no staging copy through a pageable buffer
        host = torch.empty(x.shape, dtype=x.dtype, device="cpu", pin_memory=True)
        host.copy_(x, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return host

    return moved, back


# ---- host tensors, large batches: three-stream pipeline over blocks of the batch ----------
_PIPE_MIN_BYTES = 16 << 20  # below this one copy each way is as fast
_PIPE_BLOCKS = 16  # measured on cfg2 (212 MB in): 4 -> 4.35 ms, 8 -> 4.58, 16 -> 4.29, 32 -> 4.79 (CPU-bound)
_pipe_streams = {}


def _pipe_streams_for(dev: torch.device):
    if dev.index not in _pipe_streams:
        _pipe_streams[dev.index] = tuple(torch.cuda.Stream(dev) for _ in range(3))
    return _pipe_streams[dev.index]


def _pipe_plan(ref, hyp, batch_first, ref_group):
    """Block boundaries (in reference sequences) or None when the single-copy path is the
    right one (small batch, odd shapes that the op itself must reject, emulation)."""
    if _abi.EMULATED or ref.device.type != "cpu" or hyp.device.type != "cpu":
        return None
    if not torch.cuda.is_available() or ref.dim() != 2 or hyp.dim() != 2:
        return None
    bdim = 0 if batch_first else 1
    nr, n = ref.shape[bdim], hyp.shape[bdim]
    if nr * ref_group != n or ref.shape[1 - bdim] == 0 or hyp.shape[1 - bdim] == 0:
        return None
    if ref.dtype not in _ops._INT_DTYPES or hyp.dtype not in _ops._INT_DTYPES:
        return None
    nbytes = ref.numel() * ref.element_size() + hyp.numel() * hyp.element_size()
    if nbytes < _PIPE_MIN_BYTES or nr < 2 * 256:
        return None
    blocks = min(_PIPE_BLOCKS, nr // 256)
    step = -(-nr // blocks)
    step = -(-step // 32) * 32
    return [(a, min(a + step, nr)) for a in range(0, nr, step)]


def _copy_block(lib, dev_t, host_t, batch_first, a, b, to_device, stream):
    """One DMA between rows/columns [a, b) of the batch axis of a host matrix (inner stride
    1) and the contiguous device matrix of that block."""
    es = host_t.element_size()
    if host_t.dim() == 1:
        base, pitch, width, height = a * es, (b - a) * es, (b - a) * es, 1
    elif batch_first:
        base, pitch, width, height = a * host_t.stride(0) * es, host_t.stride(0) * es, \
            host_t.shape[1] * es, b - a
    else:
        base, pitch, width, height = a * es, host_t.stride(0) * es, (b - a) * es, host_t.shape[0]
    hp, dp = host_t.data_ptr() + base, dev_t.data_ptr()
    if to_device:
        _abi.check(lib.b200lev_copy2d_async(dp, width, hp, pitch, width, height, 1, stream))
    else:
        _abi.check(lib.b200lev_copy2d_async(hp, pitch, dp, width, width, height, 0, stream))


def _string_matching_pipelined(plan, ref, hyp, op_args, batch_first, prefix, exclude_last,
                               ref_group):
    """Blocks of the batch through (H2D, kernels, D2H) on three streams; same numbers as one
    call on the whole batch (pairs are independent).  Returns (host result, flags)."""
    lib = _abi.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    s_in, s_run, s_out = _pipe_streams_for(dev)
    if ref.dim() == 2 and ref.stride(1) != 1:
        ref = ref.contiguous()
    if hyp.stride(1) != 1:
        hyp = hyp.contiguous()
    bdim = 0 if batch_first else 1
    n = hyp.shape[bdim]
    hout = hyp.shape[1 - bdim] + (0 if exclude_last else 1)
    if not prefix:
        shape = (n,)
    else:
        shape = (n, hout) if batch_first else (hout, n)
    host_out = torch.empty(shape, dtype=torch.float32, device="cpu", pin_memory=True)
    flags = None
    keep = []  # blocks stay referenced until the last stream drains
    here = torch.cuda.current_stream(dev)
    s_in.wait_stream(here)
    for (a, b) in plan:
        ha, hb = a * ref_group, b * ref_group
        with torch.cuda.stream(s_in):
            rshape = (b - a, ref.shape[1]) if batch_first else (ref.shape[0], b - a)
            hshape = (hb - ha, hyp.shape[1]) if batch_first else (hyp.shape[0], hb - ha)
            ref_d = torch.empty(rshape, dtype=ref.dtype, device=dev)
            hyp_d = torch.empty(hshape, dtype=hyp.dtype, device=dev)
            _copy_block(lib, ref_d, ref, batch_first, a, b, True, s_in.cuda_stream)
            _copy_block(lib, hyp_d, hyp, batch_first, ha, hb, True, s_in.cuda_stream)
            arrived = s_in.record_event()
        with torch.cuda.stream(s_run):
            s_run.wait_event(arrived)
            out_d, f = _ops.string_matching_fast(ref_d, hyp_d, *op_args)
            flags = f if flags is None else flags.bitwise_or_(f)
            done = s_run.record_event()
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            _copy_block(lib, out_d, host_out, batch_first, ha, hb, False, s_out.cuda_stream)
        keep.append((ref_d, hyp_d, out_d, f))
    s_out.synchronize()
    s_run.synchronize()
    del keep
    return host_out, flags


def _warn_flags(flags: torch.Tensor, eos: Optional[int], include_eos: bool, norm: bool,
                prefix: bool) -> None:
    """The data-dependent warnings of SM:202-217, 361-366, 398-404 (one 4-byte D2H)."""
    if torch.jit.is_tracing():
        return
    f = int(flags.item())
    if eos is not None and include_eos:
        for bit, name in ((_abi.FLAG_REF_NO_EOS, "ref"), (_abi.FLAG_HYP_NO_EOS, "hyp")):
            if f & bit:
                warnings.warn(
                    "include_eos=True, but a transcription in {} did not "
                    "contain the eos symbol ({}). To suppress this "
                    "warning, set warn=False".format(name, eos)
                )
    if norm and (f & _abi.FLAG_EMPTY_REF):
        if prefix:
            warnings.warn(
                "ref contains empty transcripts. Error rates will be "
                "0 for prefixes of length 0, 1 otherwise. To suppress "
                "this warning, set warn=False"
            )
        else:
            warnings.warn(
                "ref contains empty transcripts. Error rates for entries "
                "will be 1 if any insertion and 0 otherwise. To suppress "
                "this warning, set warn=False"
            )


def _string_matching(ref, hyp, eos, include_eos, batch_first, ins_cost, del_cost, sub_cost, warn,
                     norm=False, return_prf_dsts=False, exclude_last=False,
                     padding=config.INDEX_PAD_VALUE, return_mistakes=False, ref_group=1):
    """SM:146-406 for the final and prefix modes (the mask mode is folded into
    :func:`optimal_completion`)."""
    if ref.dim() != 2 or hyp.dim() != 2:
        raise RuntimeError("ref and hyp must be 2 dimensional")
    uniform = ins_cost == del_cost == sub_cost > 0.0
    if not uniform and return_mistakes and warn:  # SM:175-180
        warnings.warn(
            "The behaviour for non-uniform error rates has changed after v0.3.0. "
            "Please switch to edit_distance functions for old behaviour. Set "
            "warn=False to suppress this warning"
        )
    op_args = (eos, include_eos, batch_first, float(ins_cost), float(del_cost), float(sub_cost),
               norm, return_prf_dsts, exclude_last, int(padding), return_mistakes, ref_group)
    plan = _pipe_plan(ref, hyp, batch_first, ref_group)
    if plan is not None:
        out, flags = _string_matching_pipelined(plan, ref, hyp, op_args, batch_first,
                                                return_prf_dsts, exclude_last, ref_group)
        if warn:
            _warn_flags(flags, eos, include_eos, norm, return_prf_dsts)
        return out
    (ref_d, hyp_d), back = _offload(ref, hyp)
    out, flags = _ops.string_matching_fast(ref_d, hyp_d, *op_args)
    if warn:
        _warn_flags(flags, eos, include_eos, norm, return_prf_dsts)
    return back(out)


def error_rate(
    ref: torch.Tensor,
    hyp: torch.Tensor,
    eos: Optional[int] = None,
    include_eos: bool = False,
    norm: bool = True,
    batch_first: bool = False,
    ins_cost: float = config.DEFT_INS_COST,
    del_cost: float = config.DEFT_DEL_COST,
    sub_cost: float = config.DEFT_SUB_COST,
    warn: bool = True,
) -> torch.Tensor:
    """Functional version of ErrorRate (SM:409-434)."""
    return _string_matching(ref, hyp, eos, include_eos, batch_first, ins_cost, del_cost,
                            sub_cost, warn, norm=norm, return_mistakes=True)


def edit_distance(
    ref: torch.Tensor,
    hyp: torch.Tensor,
    eos: Optional[int] = None,
    include_eos: bool = False,
    norm: bool = False,
    batch_first: bool = False,
    ins_cost: float = config.DEFT_INS_COST,
    del_cost: float = config.DEFT_DEL_COST,
    sub_cost: float = config.DEFT_SUB_COST,
    warn: bool = True,
) -> torch.Tensor:
    """Functional version of EditDistance (SM:437-461)."""
    return _string_matching(ref, hyp, eos, include_eos, batch_first, ins_cost, del_cost,
                            sub_cost, warn, norm=norm)


def prefix_error_rates(
    ref: torch.Tensor,
    hyp: torch.Tensor,
    eos: Optional[int] = None,
    include_eos: bool = True,
    norm: bool = True,
    batch_first: bool = False,
    ins_cost: float = config.DEFT_INS_COST,
    del_cost: float = config.DEFT_DEL_COST,
    sub_cost: float = config.DEFT_SUB_COST,
    padding: int = config.INDEX_PAD_VALUE,
    exclude_last: bool = False,
    warn: bool = True,
) -> torch.Tensor:
    """Functional version of PrefixErrorRates (SM:520-550)."""
    return _string_matching(ref, hyp, eos, include_eos, batch_first, ins_cost, del_cost,
                            sub_cost, warn, norm=norm, return_prf_dsts=True,
                            exclude_last=exclude_last, padding=padding, return_mistakes=True)


def prefix_edit_distances(
    ref: torch.Tensor,
    hyp: torch.Tensor,
    eos: Optional[int] = None,
    include_eos: bool = True,
    norm: bool = False,
    batch_first: bool = False,
    ins_cost: float = config.DEFT_INS_COST,
    del_cost: float = config.DEFT_DEL_COST,
    sub_cost: float = config.DEFT_SUB_COST,
    padding: int = config.INDEX_PAD_VALUE,
    exclude_last: bool = False,
    warn: bool = True,
) -> torch.Tensor:
    """Functional version of PrefixEditDistances (SM:553-583)."""
    return _string_matching(ref, hyp, eos, include_eos, batch_first, ins_cost, del_cost,
                            sub_cost, warn, norm=norm, return_prf_dsts=True,
                            exclude_last=exclude_last, padding=padding, return_mistakes=False)


def optimal_completion(
    ref: torch.Tensor,
    hyp: torch.Tensor,
    eos: Optional[int] = None,
    include_eos: bool = True,
    batch_first: bool = False,
    ins_cost: float = config.DEFT_INS_COST,
    del_cost: float = config.DEFT_DEL_COST,
    sub_cost: float = config.DEFT_SUB_COST,
    padding: int = config.INDEX_PAD_VALUE,
    exclude_last: bool = False,
    warn: bool = True,
) -> torch.Tensor:
    """Functional version of OptimalCompletion (SM:464-517)."""
    if ref.dim() != 2 or hyp.dim() != 2:
        raise RuntimeError("ref and hyp must be 2 dimensional")
    (ref_d, hyp_d), back = _offload(ref, hyp)
    out, flags = _ops.optimal_completion_fast(ref_d, hyp_d, eos, include_eos, batch_first,
                                              float(ins_cost), float(del_cost), float(sub_cost),
                                              int(padding), exclude_last)
    if warn:
        _warn_flags(flags, eos, include_eos, False, False)
    return back(out)


def hard_optimal_completion_distillation_loss(
    logits: torch.Tensor,
    ref: torch.Tensor,
    hyp: torch.Tensor,
    eos: Optional[int] = None,
    include_eos: bool = True,
    batch_first: bool = False,
    ins_cost: float = config.DEFT_INS_COST,
    del_cost: float = config.DEFT_DEL_COST,
    sub_cost: float = config.DEFT_SUB_COST,
    weight: Optional[torch.Tensor] = None,
    reduction: str = "mean",
    ignore_index: int = -2,
    warn: bool = True,
) -> torch.Tensor:
    """Functional version of HardOptimalCompletionDistillationLoss (SM:1188-1251)."""
    if logits.dim() != 3:
        raise RuntimeError("logits must be 3 dimensional")
    if logits.shape[:-1] != hyp.shape:
        raise RuntimeError("first two dims of logits must match hyp shape")
    if include_eos:
        if eos is not None and ((eos < 0) or (eos >= logits.size(-1))):
            raise RuntimeError(f"If include_eos=True, eos ({eos}) must be a class idx")
        if eos is not None and eos == ignore_index:
            raise RuntimeError(f"If include_eos=True, eos cannot equal ignore_index ({eos}")
    if reduction not in _abi.REDUCE:
        raise RuntimeError(f"'{reduction}' is not a valid value for reduction")
    (logits_d, ref_d, hyp_d, weight_d), back = _offload(logits, ref, hyp, weight)
    optimals, flags = _ops.optimal_completion_fast(ref_d, hyp_d, eos, include_eos, batch_first,
                                                   float(ins_cost), float(del_cost),
                                                   float(sub_cost), int(ignore_index),
                                                   True)  # SM:1216-1228
    if warn:
        _warn_flags(flags, eos, include_eos, False, False)
    loss, _, _ = _ops.ocd_loss(logits_d, optimals, weight_d, int(ignore_index),
                               _abi.REDUCE[reduction], 1 if batch_first else 0)
    return back(loss)


def minimum_error_rate_loss(
    log_probs: torch.Tensor,
    ref: torch.Tensor,
    hyp: torch.Tensor,
    eos: Optional[int] = None,
    include_eos: bool = True,
    sub_avg: bool = True,
    batch_first: bool = False,
    norm: bool = True,
    ins_cost: float = config.DEFT_INS_COST,
    del_cost: float = config.DEFT_DEL_COST,
    sub_cost: float = config.DEFT_SUB_COST,
    reduction: str = "mean",
    warn: bool = True,
) -> torch.Tensor:
    """Functional version of MinimumErrorRateLoss (SM:1400-1472).

    A 2-D ``ref`` is NOT physically repeated over the samples (SM:1426, 1439): the
    kernels read reference column ``pair // samples`` instead."""
    if log_probs.dim() != 2:
        raise RuntimeError("log_probs must be 2 dimensional")
    if hyp.dim() != 3:
        raise RuntimeError("hyp must be 3 dimensional")
    if ref.dim() not in (2, 3):
        raise RuntimeError("ref must be 2 or 3 dimensional")
    if batch_first:
        batch_size, samples, max_hyp_steps = hyp.shape
        rshape = tuple(ref.shape[:2]) if ref.dim() == 3 else (ref.shape[0], samples)
    else:
        max_hyp_steps, batch_size, samples = hyp.shape
        rshape = tuple(ref.shape[1:]) if ref.dim() == 3 else (ref.shape[1], samples)
    if rshape != (batch_size, samples) or rshape != tuple(log_probs.shape):
        raise RuntimeError("ref and hyp batch_size and sample dimensions must match")
    if samples < 2:
        raise RuntimeError(f"Batch must have at least two samples, got {samples}")
    if reduction not in _abi.REDUCE:
        raise RuntimeError(f"'{reduction}' is not a valid value for reduction")
    group = samples if ref.dim() == 2 else 1
    if batch_first:
        hyp2 = hyp.reshape(-1, max_hyp_steps)
        ref2 = ref if ref.dim() == 2 else ref.reshape(-1, ref.size(-1))
    else:
        hyp2 = hyp.reshape(max_hyp_steps, -1)
        ref2 = ref if ref.dim() == 2 else ref.reshape(ref.size(0), -1)
    (lp_d, ref_d, hyp_d), back = _offload(log_probs, ref2, hyp2)
    uniform = ins_cost == del_cost == sub_cost > 0.0
    if not uniform and warn:  # SM:175-180 via error_rate
        warnings.warn(
            "The behaviour for non-uniform error rates has changed after v0.3.0. "
            "Please switch to edit_distance functions for old behaviour. Set "
            "warn=False to suppress this warning"
        )
    er, flags = _ops.string_matching_fast(ref_d, hyp_d, eos, include_eos, batch_first,
                                          float(ins_cost), float(del_cost), float(sub_cost), norm,
                                          False, False, 0, True, group)  # SM:1451-1462
    if warn:
        _warn_flags(flags, eos, include_eos, norm, False)
    loss = _ops.mwer_loss(er, lp_d, sub_avg, _abi.REDUCE[reduction])
    return back(loss)


def fill_after_eos(
    tokens: torch.Tensor,
    eos: int,
    dim: int = 0,
    fill: Optional[float] = None,
    value: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Functional version of FillAfterEndOfSequence (SM:30-42)."""
    out = tokens if value is None else value
    fill_ = float(eos) if fill is None else fill
    (tok_d, out_d), back = _offload(tokens, out)
    fill_mask = _ops.after_eos_mask(tok_d, int(eos), int(dim))
    return back(out_d.masked_fill(fill_mask, fill_))


def sequence_log_probs(logits, hyp: torch.Tensor, dim: int = 0, eos: Optional[int] = None) -> torch.Tensor:
    """Functional version of SequenceLogProbabilities, tensor path (_decoding.py:1516-1548,
    1579-1633): joint log-probability of the token sequences in ``hyp`` (``(A*, T, B*)``, step
    axis ``dim``) under the categorical distributions ``logits`` (``(A*, T, B*, V)``).  Tokens
    outside ``[0, V)`` are padding; with ``eos`` the first eos step is the last one counted.
    Differentiable with respect to ``logits``.

    Only tensors: the reference's PackedSequence branch (_decoding.py:1551-1576) is a host-side
    re-packing of the same sum and is not part of the hot-path scope."""
    if not isinstance(logits, torch.Tensor):
        raise RuntimeError("logits must be a Tensor (the PackedSequence path is not implemented "
                           "by b200lev)")
    hyp_dim = hyp.dim()
    if dim < -hyp_dim or dim > hyp_dim - 1:  # _decoding.py:1521-1525
        raise RuntimeError(
            "Dimension out of range (expected to be in range of [{}, {}], but "
            "got {})".format(-hyp_dim, hyp_dim - 1, dim)
        )
    dim = (hyp_dim + dim) % hyp_dim
    if logits.dim() != hyp_dim + 1 or tuple(logits.shape[:-1]) != tuple(hyp.shape):
        raise RuntimeError(
            "logits must have the shape of hyp plus a class axis: got {} and {}".format(
                tuple(logits.shape), tuple(hyp.shape)))
    if not logits.is_floating_point():
        raise RuntimeError("logits must be floating point")
    (logits_d, hyp_d), back = _offload(logits, hyp)
    outer = 1
    for d in hyp.shape[:dim]:
        outer *= d
    inner = 1
    for d in hyp.shape[dim + 1:]:
        inner *= d
    T, V = hyp.shape[dim], logits.shape[-1]
    out, _, _ = _ops.sequence_log_probs_fast(
        logits_d.contiguous().view(outer, T, inner, V),
        hyp_d.to(torch.long).contiguous().view(outer, T, inner), eos)
    return back(out.view(tuple(hyp.shape[:dim]) + tuple(hyp.shape[dim + 1:])))


def ctc_greedy_search(logits: torch.Tensor, in_lens: Optional[torch.Tensor] = None, blank_idx: int = -1,
                      batch_first: bool = False, is_probs: bool = False
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Functional version of CTCGreedySearch (_decoding.py:507-560): per step the most likely
    class of ``logits`` (``(T, N, V)``, or ``(N, T, V)`` if ``batch_first``); returns the path
    score ``max_ (N,)`` (sum of the chosen log-probabilities; product of the chosen
    probabilities with ``is_probs``), ``paths`` (``(T, N)`` / ``(N, T)`` long: blanks and repeats
    removed, compacted to the front) and ``out_lens (N,)``.  Equal maxima resolve to the lowest
    class index.  ``max_`` is differentiable w.r.t. ``logits`` when ``is_probs`` is false."""
    if logits.dim() != 3:
        raise RuntimeError("logits must be 3-dimensional")
    V = logits.size(2)
    if blank_idx < -V or blank_idx > (V - 1):
        raise RuntimeError(
            "Blank index out of range (expected to be in the range of "
            f"[-{V},{V-1}], but got {blank_idx})"
        )
    if not logits.is_floating_point():
        raise RuntimeError("logits must be floating point")
    blank = (blank_idx + V) % V
    N = logits.size(0) if batch_first else logits.size(1)
    T = logits.size(1) if batch_first else logits.size(0)
    if in_lens is not None and tuple(in_lens.shape) != (N,):
        raise RuntimeError(f"in_lens must have shape ({N},), got {tuple(in_lens.shape)}")
    (logits_d, lens_d), back = _offload(logits, in_lens)
    outer, inner = (N, 1) if batch_first else (1, N)
    max_, paths, out_lens, _, _, _ = _ops.ctc_greedy_search_fast(
        logits_d.contiguous().view(outer, T, inner, V), lens_d, blank, is_probs)
    shape = (N, T) if batch_first else (T, N)
    return back(max_.view(N)), back(paths.view(shape)), back(out_lens.view(N))
