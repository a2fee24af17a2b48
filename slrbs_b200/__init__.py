"""slrbs_b200 -- B200-native (sm_100a) per-step rigid-body pipeline behind a C ABI.

The product is the shared library built from slrbs_b200/csrc (CUDA kernels + C ABI,
include/slrbs_b200.h) and the C++ mirror of the reference's class surface in
slrbs_b200/host.  This Python package is a thin ctypes binding used by the tests and
bench.py; it never falls back to a CPU implementation.
"""
from .capi import Context, SlrbsError, lib_path, load_library  # noqa: F401
