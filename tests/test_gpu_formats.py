"""Dequantised weight values must be BIT-EXACT (north star): the dequant entry point, and the weights the hot
matvec kernel effectively multiplies by (extracted with one-hot activations), against the CPU oracle / quants.py."""
import os

import numpy as np
import pytest

from oracle import oracle
from xalm_b200 import capi
from xalm_b200 import types as T

from gpu_util import bits, random_raw

pytestmark = pytest.mark.gpu

BLOCK = ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0", "tq1_0"]


@pytest.mark.parametrize("fmt", BLOCK)
def test_dequant_block_formats_bit_exact_vs_quants_py(golden_dir, fmt):
    g = np.load(os.path.join(golden_dir, f"blocks_{fmt}.npz"))
    t = T.parse(fmt)
    for raw_key, deq_key in (("qbytes", "deq"), ("raw", "raw_deq")):
        got = capi.dequant(t.id, g[raw_key], g[deq_key].size).reshape(g[deq_key].shape)
        assert np.array_equal(bits(got), bits(g[deq_key])), f"{fmt}/{raw_key}"


@pytest.mark.parametrize("t", [T.F8_E2M5, T.F8_E3M4, T.F8_E4M3, T.F8_E5M2, T.Q8, T.QI8, T.U8], ids=lambda t: t.name)
def test_dequant_every_byte_code(t):
    codes = np.arange(256, dtype=np.uint8)
    assert np.array_equal(bits(capi.dequant(t.id, codes, 256)), bits(oracle.dequant(t.id, codes, 256)))


def test_dequant_fp8_against_torch_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "fp8_codes.npz"))
    codes = np.arange(256, dtype=np.uint8)
    assert np.array_equal(bits(capi.dequant(T.F8_E4M3.id, codes, 256)), bits(g["e4m3_ref"]))
    assert np.array_equal(bits(capi.dequant(T.F8_E5M2.id, codes, 256)), bits(g["e5m2_ref"]))


def test_dequant_wide_types_and_sizes():
    rng = np.random.default_rng(0)
    for n in (32, 4096, 1 << 20):
        x = rng.standard_normal(n).astype(np.float32)
        assert np.array_equal(capi.dequant(T.F32.id, x.view(np.uint8), n), x)
        h = x.astype(np.float16)
        assert np.array_equal(capi.dequant(T.F16.id, h.view(np.uint8), n), h.astype(np.float32))
        bf = (x.view(np.uint32) >> 16).astype(np.uint16)
        assert np.array_equal(capi.dequant(T.BF16.id, bf.view(np.uint8), n), oracle.dequant(T.BF16.id, bf.view(np.uint8), n))
    assert capi.dequant(T.F16.id, np.zeros(0, np.uint8), 0).size == 0          # empty input


# (rows, cols): small/few-row shapes take the K-split config, >= 6144 rows the many-row config
SHAPES = [(64, 256), (32, 512), (128, 1024), (6144, 256)]   # rows of <= 64 pieces also take the 16-row short-row tiles


@pytest.mark.parametrize("t", T.MATMUL_TYPES, ids=lambda t: t.name)
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_matvec_effective_weights_bit_exact(t, shape):
    """y = W e_j is column j of the weights as the HOT kernel decodes them: must equal the oracle's dequantised
    column bit for bit (the fused d * sum(q*x) arithmetic is exact for one-hot x)."""
    d, n = shape
    raw = random_raw(t, d, n, seed=11)
    W = oracle.dequant(t.id, raw, d * n).reshape(d, n)
    for j in (0, 1, 15, 16, 31, 32, n // 2 + 5, n - 1):
        e = np.zeros(n, np.float32)
        e[j] = 1.0
        y = capi.matmul(e, raw, t.id, n, d)
        # value equality: a weight of -0.0 comes back as +0.0 from a sum of signed zeros, nothing else may differ
        assert np.array_equal(y, W[:, j]) and np.all(np.isfinite(y)), f"{t.name} column {j}"


def test_matvec_fp8_tensor_with_nan_codes_takes_exact_path():
    for t in (T.F8_E4M3, T.F8_E5M2):
        raw = np.random.default_rng(5).integers(0, 256, size=64 * 256, dtype=np.uint8)   # includes 0x7F / 0x7C.. codes
        W = oracle.dequant(t.id, raw, raw.size).reshape(64, 256)
        assert np.all(np.isfinite(W))
        x = np.random.default_rng(6).standard_normal(256).astype(np.float32)
        y = capi.matmul(x, raw, t.id, 256, 64)
        ref = W.astype(np.float64) @ x.astype(np.float64)
        assert np.max(np.abs(y - ref) / (np.abs(ref) + 1.0)) < 1e-5


@pytest.mark.parametrize("t", T.MATMUL_TYPES, ids=lambda t: t.name)
def test_matvec_random_x_vs_oracle(t):
    for d, n in [(96, 768), (6144, 512), (4096, 4096)]:
        if t is T.TQ1_0 and n % 256:
            continue
        raw = random_raw(t, d, n, seed=3)
        x = np.random.default_rng(4).standard_normal(n).astype(np.float32)
        y = capi.matmul(x, raw, t.id, n, d)
        ref = oracle.matmul(x, raw, t.id, n, d, acc_mode=2)          # fp64 accumulate = noise-free reference
        scale = np.sqrt(n) * np.abs(oracle.dequant(t.id, raw[: t.nbytes(n)], n)).max()
        assert np.max(np.abs(y - ref)) <= 2e-6 * scale + 1e-6, f"{t.name} {d}x{n}"


@pytest.mark.parametrize("t", [T.Q8_0, T.Q4_0, T.Q4_1, T.Q5_0, T.Q5_1, T.Q8], ids=lambda t: t.name)
def test_matvec_tensor_core_shapes_vs_oracle(t):
    """The tensor-core matvec (matvec_mma.cuh) on shapes with more 16-row tiles than SMs (150 / 300 / 151), rows that are not a
    multiple of a ring stage (33 and 129 block columns), and run-to-run bit identity."""
    for d, n in [(2400, 1056), (4800, 2048), (2432, 4128)]:
        raw = random_raw(t, d, n, seed=5)
        x = np.random.default_rng(8).standard_normal(n).astype(np.float32)
        y = capi.matmul(x, raw, t.id, n, d)
        ref = oracle.matmul(x, raw, t.id, n, d, acc_mode=2)
        scale = np.sqrt(n) * np.abs(oracle.dequant(t.id, raw[: t.nbytes(n)], n)).max()
        assert np.max(np.abs(y - ref)) <= 2e-6 * scale + 1e-6, f"{t.name} {d}x{n}"
        assert np.array_equal(bits(y), bits(capi.matmul(x, raw, t.id, n, d))), "not bit-reproducible run to run"


def test_matvec_rejects_what_the_reference_rejects():
    x = np.zeros(64, np.float32)
    with pytest.raises(capi.XalmError) as e:
        capi.matmul(x, np.zeros(64 * 64, np.uint8), T.U8.id, 64, 64)     # unsupported dtype (infer.cpp:211-214)
    assert e.value.status == 2
    with pytest.raises(capi.XalmError):
        capi.matmul(x, np.zeros(64 * 54 // 4, np.uint8), T.TQ1_0.id, 64, 64)   # row not a multiple of the 256 block
