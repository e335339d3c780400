"""Constants the hot path reads as default arguments.

Mirrors the five values of the reference's ``pydrobert/torch/config.py`` that the
string-matching signatures use (config.py:55, 156-162).
"""

INDEX_PAD_VALUE = -100
"""The value to pad index-based tensors with (config.py:55)"""

DEFT_INS_COST = 1.0
"""Default insertion cost in error rate/distance computations (config.py:156)"""

DEFT_DEL_COST = 1.0
"""Default deletion cost in error rate/distance computations (config.py:159)"""

DEFT_SUB_COST = 1.0
"""Default substitution cost in error rate/distance computations (config.py:162)"""

DEFT_FILE_PREFIX = ""
"""Default prefix of a torch data file in a data directory (config.py:86)"""

DEFT_FILE_SUFFIX = ".pt"
"""Default suffix of a torch data file in a data directory (config.py:89)"""
