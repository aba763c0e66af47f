"""B200-native implementation of mi-fieldcalc's FieldCalculations hot path.

The product is the native library mi-fieldcalc_b200/lib/libfcb200.so (hand-written sm_100a CUDA
kernels behind the C-ABI of include/fcb200.h) and the C++ drop-in shim libmi-fieldcalc.so.0 built
on top of it.  This Python package only holds the build script, a ctypes binding used by the
tests and the benchmark, and the multi-GPU sharding helpers.  The directory name carries a hyphen:
import it with importlib.import_module("mi-fieldcalc_b200").
"""
from . import capi  # noqa: F401
from .capi import ALL_DEFINED, NONE_DEFINED, SOME_DEFINED, Fcb200, load  # noqa: F401
