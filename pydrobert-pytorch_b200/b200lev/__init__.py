"""b200lev -- B200-native (sm_100a) batched Levenshtein hot path.

A drop-in for the string-matching family of ``pydrobert.torch`` (its ``_string.py``):
``b200lev.functional`` and ``b200lev.modules`` carry the reference's names and
signatures; ``b200lev.install()`` rebinds them inside an installed ``pydrobert.torch``.
All computation happens in ``libb200lev.so`` (hand-written CUDA behind the C ABI of
``include/b200lev.h``); there is no CPU fallback.
"""
from . import config, functional, modules, scoring  # noqa: F401
from ._abi import B200LevError  # noqa: F401
from .install import install, uninstall  # noqa: F401

__version__ = "0.1.0"
