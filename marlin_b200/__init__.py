"""marlin_b200 - B200-native spectral time-step path behind Marlin's operator API.

The product is libmarlin_b200.so (hand-written sm_100a CUDA kernels + the C ABI declared in
include/marlin_b200.h).  `marlin_b200.capi` is a thin ctypes binding used by the tests and by
bench.py; the MOOSE-facing host objects live in C++ under marlin_b200/host/.
"""
