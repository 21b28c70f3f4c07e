"""folve_b200 -- B200 (sm_100a) convolution engine behind folve's SoundProcessor.

The product is folve_b200/libfolve_b200.so (C ABI, include/folve_b200.h) and
the C++ host layer in folve_b200/host/.  This Python package only carries the
ctypes plumbing used by tests/ and bench.py.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
