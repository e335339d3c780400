"""The constructor checks the string-matching modules use.

Same behaviour and messages as the reference's ``pydrobert/torch/argcheck.py``
(``_type_check_factory`` :202-223, ``is_in`` :316-325): a value of an accepted type is
returned converted to the canonical type, anything else raises
``ValueError("<name> (<val>) is not a[n] <type>")``.
"""
from typing import Any, Collection, Optional

import numpy as np
import torch


def _nv(name: Optional[str], val: Any) -> str:
    if isinstance(val, torch.Tensor):
        if val.numel() == 1:
            return f"{val.item()}" if name is None else f"{name} ({val.item()})"
        return name if name is not None else "tensor"
    if isinstance(val, str):
        val = f"'{val}'"
    return f"{val}" if name is None else f"{name} ({val})"


def _type_check(t, *ts):
    ts = (t,) + ts

    def check(val, name=None, allow_none=False):
        if val is None and allow_none:
            return val
        if isinstance(val, ts):
            return val if (type(val) is t) else t(val)
        tname = t.__name__
        x = "n" if tname.startswith(("a", "e", "i", "o", "u")) else ""
        raise ValueError(f"{_nv(name, val)} is not a{x} {tname}")

    return check


is_int = _type_check(int, np.integer)
is_bool = _type_check(bool)
is_float = _type_check(float, int, np.integer, np.floating)
is_tensor = _type_check(torch.Tensor)


def is_in(val, collection: Collection, name=None, allow_none=False):
    if allow_none and val is None:
        return None
    if val not in collection:
        raise ValueError(f"{_nv(name, val)} is not one of {collection}")
    return val
