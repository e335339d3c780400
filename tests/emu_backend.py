"""Load the SIMT-emulated build of the kernel sources into the host layer (CPU tests).

TEST INFRASTRUCTURE ONLY -- see tests/emu/emu_cuda.h.  With this backend active the
Python host code (b200lev.functional / modules / _ops) runs unchanged on CPU tensors
and every kernel executes, thread by thread, on the host, so the `-m "not gpu"` suite
checks kernel logic + host logic against the oracle without a GPU.  The GPU suite
(`-m gpu`) never imports this module.
"""
import contextlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))


@contextlib.contextmanager
def emulated_kernels():
    import build_emu
    from b200lev import _abi

    path = build_emu.build()
    _abi._set_library_for_tests(path)
    try:
        yield
    finally:
        _abi._set_library_for_tests(None)
