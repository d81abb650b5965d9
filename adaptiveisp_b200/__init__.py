"""B200-native (sm_100a) implementation of AdaptiveISP's differentiable ISP filter chain.

The product is ``csrc/libaisp_b200.so`` (C ABI: ``include/aisp_b200.h``); this package is the thin
Python host that mirrors the reference's ``isp/filters.py`` class API over it.
"""
from . import _lib  # noqa: F401
from ._lib import AispError, build, lib  # noqa: F401

__version__ = "0.1.0"
