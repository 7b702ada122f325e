"""B200-native correlated progressive photon mapping: Python face of libcpm_b200.so.

The product is the CUDA library + C ABI (include/cpm_b200.h) and the C++ host mirror of the
reference's Inviwo processors (host/).  This package is the ctypes binding used by tests and
benchmarks; it holds no compute of its own.
"""
from .capi import *  # noqa: F401,F403
from .capi import Context, CpmError, TraceParams, Volume, lib, declared_symbols, make_trace_params  # noqa: F401
