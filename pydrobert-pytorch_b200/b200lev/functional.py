"""The eight functionals of the reference's string-matching family, same signatures.

Mirrors ``pydrobert.torch.functional`` (functional.py:49-58 of the reference) for the
names backed by ``pydrobert/torch/_string.py`` ("SM" below): argument meaning,
defaults, shapes, dtypes, error messages and the three data-dependent warnings are
the reference's; the numbers come from the sm_100a kernels behind ``torch.ops.b200lev``.

Host tensors: the reference runs wherever its inputs live.  Here a CPU tensor is
copied to the current CUDA device (asynchronously when it is pinned), the kernels run
there and the result is copied back -- the ``e2e`` path of ``bench.py``.  Without a
CUDA device the call raises; there is no CPU implementation.
"""
from __future__ import annotations

import warnings
from typing import Optional

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
    (ref_d, hyp_d), back = _offload(ref, hyp)
    out, flags = _ops.string_matching_fast(ref_d, hyp_d, eos, include_eos, batch_first, float(ins_cost),
                                      float(del_cost), float(sub_cost), norm, return_prf_dsts,
                                      exclude_last, int(padding), return_mistakes, ref_group)
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
