"""The constructor checks the string-matching modules use.

Written against the reference's observable contract (``pydrobert/torch/argcheck.py``):
``is_<kind>(val, name=None, allow_none=False)`` returns ``val`` converted to the canonical
Python type of the kind, or raises ``ValueError("<what> is not a[n] <kind>")`` where
``<what>`` is ``name (val)``, just ``val`` without a name, strings quoted, one-element
tensors shown by value and larger tensors by name only; ``is_in`` raises
``ValueError("<what> is not one of <collection>")``.
"""
from typing import Any, Collection, Optional, Tuple

import numpy as np
import torch


def _describe(val: Any, name: Optional[str]) -> str:
    """How a rejected value is shown in the message."""
    if isinstance(val, torch.Tensor) and val.numel() != 1:
        return "tensor" if name is None else name
    if isinstance(val, torch.Tensor):
        shown = str(val.item())
    elif isinstance(val, str):
        shown = "'" + val + "'"
    else:
        shown = str(val)
    return shown if name is None else "{} ({})".format(name, shown)


class _Kind:
    """One accepted kind of constructor argument: the canonical type it is returned as and
    the other types that convert to it."""

    def __init__(self, canonical: type, also: Tuple[type, ...] = ()):
        self.canonical = canonical
        self.accepted = (canonical,) + tuple(also)
        label = canonical.__name__
        self.phrase = ("an " if label[0] in "aeiou" else "a ") + label

    def __call__(self, val: Any, name: Optional[str] = None, allow_none: bool = False) -> Any:
        if val is None and allow_none:
            return None
        if not isinstance(val, self.accepted):
            raise ValueError("{} is not {}".format(_describe(val, name), self.phrase))
        if type(val) is not self.canonical:
            val = self.canonical(val)
        return val


is_int = _Kind(int, (np.integer,))
is_bool = _Kind(bool)
is_float = _Kind(float, (int, np.integer, np.floating))
is_tensor = _Kind(torch.Tensor)


def is_in(val: Any, collection: Collection, name: Optional[str] = None, allow_none: bool = False) -> Any:
    if val is None and allow_none:
        return None
    if val in collection:
        return val
    raise ValueError("{} is not one of {}".format(_describe(val, name), collection))
