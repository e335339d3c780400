"""Host-tensor transport for the registered ops.

The reference runs wherever its inputs live (_string.py:146 has no device requirement), so
every ``torch.ops.b200lev.*`` op accepts tensors on any device: the computation ALWAYS runs
on a CUDA device (there is no CPU implementation); tensors that live on the host are copied
to the current CUDA device first and the results are copied back to where the inputs came
from.  :class:`Placement` is that rule; :func:`string_matching_blocks` is the large-batch
form of it for the final / prefix modes -- blocks of the batch axis flow through three
streams (copy in, kernels, copy out) so the PCIe transfers of both directions overlap the
kernels, and the int64 tokens of a block are narrowed to the smallest integer type that
holds them before they cross the bus (``narrow``: the kernels read 2/4/8-byte tokens).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _abi


def compute_device(home: torch.device) -> torch.device:
    """The CUDA device the kernels run on for tensors that live on ``home``."""
    if home.type == "cuda":
        return home
    if home.type != "cpu" or not torch.cuda.is_available():
        raise _abi.B200LevError(
            f"b200lev kernels need a CUDA device (tensors on {home}); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Placement:
    """Where an op's tensors live (``home``) and where it computes (``dev``)."""

    __slots__ = ("home", "dev", "moved")

    def __init__(self, *tensors: Optional[Tensor]):
        first = next(t for t in tensors if t is not None)
        self.home = first.device
        for t in tensors:
            if t is not None and t.device != self.home:
                raise RuntimeError(f"expected all tensors on {self.home}, got one on {t.device}")
        self.dev = compute_device(self.home)
        self.moved = self.dev != self.home

    def to_dev(self, t: Optional[Tensor]) -> Optional[Tensor]:
        if t is None or not self.moved:
            return t
        return t.to(self.dev, non_blocking=True)

    def back(self, t: Tensor) -> Tensor:
        if not self.moved:
            return t
        # D2H into page-locked memory (torch's caching host allocator recycles the block):
        # one DMA, no staging copy through a pageable buffer
        host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return host


# ---- host tensors, large batches: three-stream pipeline over blocks of the batch ----------
PIPE_MIN_BYTES = 16 << 20  # below this one copy each way is as fast
PIPE_BLOCKS = 16  # measured on cfg2 (212 MB in): 4 -> 4.35 ms, 8 -> 4.58, 16 -> 4.29, 32 -> 4.79
_INT_DTYPES = (torch.int64, torch.int32, torch.int16, torch.int8)
_pipe_streams = {}


def _streams_for(dev: torch.device):
    if dev.index not in _pipe_streams:
        _pipe_streams[dev.index] = tuple(torch.cuda.Stream(dev) for _ in range(3))
    return _pipe_streams[dev.index]


def block_plan(ref: Tensor, hyp: Tensor, batch_first: bool, ref_group: int
               ) -> Optional[List[Tuple[int, int]]]:
    """Block boundaries (in reference sequences) or None when the single-copy path is the
    right one (device tensors, small batch, odd shapes that the op itself must reject)."""
    if ref.device.type != "cpu" or hyp.device.type != "cpu" or not torch.cuda.is_available():
        return None
    if ref.dim() != 2 or hyp.dim() != 2:
        return None
    bdim = 0 if batch_first else 1
    nr, n = ref.shape[bdim], hyp.shape[bdim]
    if nr * ref_group != n or ref.shape[1 - bdim] == 0 or hyp.shape[1 - bdim] == 0:
        return None
    if ref.dtype not in _INT_DTYPES or hyp.dtype not in _INT_DTYPES:
        return None
    nbytes = ref.numel() * ref.element_size() + hyp.numel() * hyp.element_size()
    if nbytes < PIPE_MIN_BYTES or nr < 2 * 256:
        return None
    blocks = min(PIPE_BLOCKS, nr // 256)
    step = -(-nr // blocks)
    step = -(-step // 32) * 32
    return [(a, min(a + step, nr)) for a in range(0, nr, step)]


def _copy_block(lib, dev_t: Tensor, host_t: Tensor, batch_first: bool, a: int, b: int,
                to_device: bool, stream: int) -> None:
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


def string_matching_blocks(plan: Sequence[Tuple[int, int]], ref: Tensor, hyp: Tensor,
                           run: Callable[[Tensor, Tensor], Tuple[Tensor, Tensor]],
                           batch_first: bool, prefix: bool, exclude_last: bool, ref_group: int
                           ) -> Tuple[Tensor, Tensor]:
    """Blocks of the batch through (H2D, kernels, D2H) on three streams; same numbers as one
    call on the whole batch (pairs are independent).  ``run(ref_block, hyp_block)`` is the
    device-side call.  Returns (page-locked host result, host flags)."""
    lib = _abi.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    s_in, s_run, s_out = _streams_for(dev)
    if ref.stride(1) != 1:
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
            out_d, f = run(ref_d, hyp_d)
            flags = f if flags is None else flags.bitwise_or_(f)
            done = s_run.record_event()
        with torch.cuda.stream(s_out):
            s_out.wait_event(done)
            _copy_block(lib, out_d, host_out, batch_first, ha, hb, False, s_out.cuda_stream)
        keep.append((ref_d, hyp_d, out_d, f))
    s_out.synchronize()
    s_run.synchronize()
    host_flags = flags.cpu()
    del keep
    return host_out, host_flags
