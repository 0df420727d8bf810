"""CPU-side checks of the boundary: the library loads, exports every symbol the header declares, and refuses to
compute without a GPU (no silent fallback).  No kernels are launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from xalm_b200 import capi, build
from xalm_b200 import types as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        return capi.device_count() > 0
    except capi.XalmError:
        return False


def test_library_exports_every_symbol_in_the_header():
    hdr = open(os.path.join(ROOT, "include", "xalm_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(xalm_cuda_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    assert declared == set(capi.SYMBOLS), f"binding and header disagree: {declared ^ set(capi.SYMBOLS)}"
    L = C.CDLL(build.build_cuda())
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/xalm_cuda.h but not exported"
    assert capi.lib().xalm_cuda_abi_version() == 1


def test_type_table_matches_reference_sizes():
    L = capi.lib()
    for t in T.ALL:
        b, s = C.c_int(), C.c_int()
        assert L.xalm_cuda_type_info(t.id, C.byref(b), C.byref(s)) == 0
        assert (b.value, s.value) == (t.block, t.bytes), t.name
    assert L.xalm_cuda_type_info(12345, None, None) != 0
    assert b"invalid type" in L.xalm_cuda_last_error()


def test_argument_validation_needs_no_gpu():
    x = np.zeros(48, np.float32)
    with pytest.raises(capi.XalmError) as e:
        capi.matmul(x, np.zeros(48 * 32, np.uint8), T.Q8.id, 48, 32)      # n % 32 (infer.cpp:110)
    assert e.value.status == 1
    with pytest.raises(capi.XalmError):
        capi.dequant(T.Q4_0.id, np.zeros(18, np.uint8), 31)               # not a whole block
    with pytest.raises(capi.XalmError):
        capi.rmsnorm(x, np.zeros(96, np.uint8), T.F16.id, 1e-5)           # F32/BF16 only (infer.cpp:248-249)


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_a_device():
    x = np.ones(32, np.float32)
    with pytest.raises(capi.XalmError) as e:
        capi.matmul(x, np.zeros((32, 32), np.float16).view(np.uint8), T.F16.id, 32, 32)
    assert e.value.status == 3   # XALM_ERR_CUDA
    with pytest.raises(capi.XalmError):
        capi.dequant(T.F16.id, np.zeros(64, np.uint8), 32)
    from xalm_b200.model import Model, InferenceState
    from xalm_b200 import synth
    c = synth.model_config("tiny")
    m = Model.from_tensors({**c, "act": 1}, [])
    with pytest.raises(RuntimeError):
        m.forward(InferenceState(c), 0, 0)                                 # not on a device: no CPU path
    with pytest.raises(RuntimeError):
        m.prefill([1, 2, 3])                                               # nor for the batched path
    with pytest.raises(capi.XalmError) as e:
        capi.gemm(np.ones((4, 32), np.float32), np.zeros((32, 32), np.float16).view(np.uint8), T.F16.id, 32, 32)
    assert e.value.status == 3
    with pytest.raises(capi.XalmError):
        capi.bench_gemm(128, 256, 64)
    with pytest.raises(capi.XalmError):
        m.cuda()
