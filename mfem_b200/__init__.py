"""mfem_b200 -- B200-native (sm_100a) partial-assembly operator path behind MFEM's operator API.

The product is the C-ABI library mfem_b200/lib/libb200fem.so (sources: mfem_b200/csrc, header:
include/b200fem.h) and the C++ adapter in adapter/.  This package is the thin Python mirror
of the reference's host interface used by tests/ and bench.py.  There is no CPU fallback:
without the built library (or without a CUDA device) the compute entry points raise.
"""
from . import capi  # noqa: F401
from .fem import *  # noqa: F401,F403
