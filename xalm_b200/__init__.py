"""xalm_b200 — B200-native (sm_100a) `-d cuda` backend for the Xalm transformer forward pass.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI declared in
include/xalm_cuda.h + the C++ host mirror of the reference's Model/InferenceState surface) and a
thin ctypes layer the tests and bench.py drive it through.  See DESIGN.md.
"""
__version__ = "0.1.0"
