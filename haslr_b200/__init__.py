"""haslr_b200 — B200-native hot path of HASLR's haslr_assemble (compact long reads, backbone edge table, batched POA).

The product is the CUDA library haslr_b200/libhaslr_b200.so (C ABI in include/haslr_b200.h); this package is a thin
ctypes binding used by the tests and bench.py. There is no CPU fallback: importing `haslr_b200.ffi` without the built
library, or creating a context without a B200-class GPU, raises.
"""
from .ffi import Context, HgpuError, lib_path, load  # noqa: F401
