"""B200-native eaQHM analysis+synthesis hot path (drop-in for the reference's
eaQHMAnalysisAndSynthesis adaptation loop, functions.py:163-402).

Python host code calls hand-written sm_100a CUDA kernels through the C ABI of
libeaqhm_b200.so (include/eaqhm_b200.h).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import EaqhmError, build  # noqa: F401
from .structs import Deterministic, Frame  # noqa: F401
