"""t2b200: B200-native DVB-T2 demodulation + FEC hot path (CUDA behind a C-ABI).

The product is sdr_receiver_dvb_t2_b200/libt2b200.so (include/t2b200.h).  This package is the thin
Python binding used by tests/ and bench.py; it never computes anything on the CPU and fails loudly
when the library or a GPU is missing.
"""
from .engine import Engine, T2Error, lib, lib_path  # noqa: F401

__all__ = ['Engine', 'T2Error', 'lib', 'lib_path']
