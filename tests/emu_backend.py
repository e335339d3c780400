"""Load the SIMT-emulated build of the kernel sources into the host layer (CPU tests).

TEST INFRASTRUCTURE ONLY -- see tests/emu/emu_cuda.h.  The product has no seam for this:
the context manager below monkeypatches, from the outside, the four places where the host
layer touches CUDA (the loaded library handle, the compute-device rule, the stream handle
and the device guard).  With it active the Python host code (b200lev.functional / modules /
_ops) runs unchanged on CPU tensors and every kernel executes, thread by thread, on the
host, so the `-m "not gpu"` suite checks kernel logic + host logic against the oracle
without a GPU.  The GPU suite (`-m gpu`) never imports this module.
"""
import contextlib
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

ACTIVE = False


@contextlib.contextmanager
def emulated_kernels():
    global ACTIVE
    import build_emu
    from b200lev import _abi, _host, _ops

    path = build_emu.build()
    saved = (_abi._lib, _host.compute_device, _ops._stream, _ops._DeviceGuard)
    _abi._lib = _abi._bind(ctypes.CDLL(path))
    _host.compute_device = lambda home: home
    _ops._stream = lambda dev: 0
    _ops._DeviceGuard = lambda dev: contextlib.nullcontext()
    ACTIVE = True
    try:
        yield
    finally:
        _abi._lib, _host.compute_device, _ops._stream, _ops._DeviceGuard = saved
        ACTIVE = False
