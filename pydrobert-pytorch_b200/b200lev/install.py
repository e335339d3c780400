"""Rebind the reference's string-matching names to the B200 kernels.

``install()`` patches an importable ``pydrobert.torch`` in place -- ``_string``,
``_decoding``, ``functional`` and ``modules`` (functional.py:49-58, modules.py:115-124 of
the reference) -- so that existing user code and the reference's own
``tests/test_string.py`` run on ``libb200lev.so`` without edits.  Rebinding
``_decoding.beam_search_advance`` / ``random_walk_advance`` also puts the step kernels under
the reference's own ``BeamSearch`` / ``RandomWalk`` modules, whose loops call them by name.
"""
from __future__ import annotations

import importlib

from . import functional as _F
from . import modules as _M

_NOT_REBOUND = ()  # (every functional and module of the family is rebound)
_FUNCS = tuple(n for n in _F.__all__ if n not in _NOT_REBOUND)
_CLASSES = tuple(n for n in _M.__all__ if n not in _NOT_REBOUND)
_saved = {}


def install() -> bool:
    """Returns False if ``pydrobert.torch`` is not importable (nothing to patch)."""
    try:
        mods = [importlib.import_module(f"pydrobert.torch.{m}")
                for m in ("_string", "_decoding", "functional", "modules")]
    except ImportError:
        return False
    for mod in mods:
        for name in _FUNCS:
            if hasattr(mod, name):
                _saved.setdefault((mod.__name__, name), getattr(mod, name))
                setattr(mod, name, getattr(_F, name))
        for name in _CLASSES:
            if hasattr(mod, name):
                _saved.setdefault((mod.__name__, name), getattr(mod, name))
                setattr(mod, name, getattr(_M, name))
    return True


def uninstall() -> None:
    for (modname, name), obj in list(_saved.items()):
        setattr(importlib.import_module(modname), name, obj)
    _saved.clear()
