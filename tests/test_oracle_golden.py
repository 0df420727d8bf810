"""Pin the CPU oracle against the fixtures the REFERENCE's own Python produced
(tests/golden/make_golden.py: quants.py, torch fp8 casts, convert.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle
from xalm_b200 import types as T

BLOCK_FORMATS = ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0", "tq1_0"]


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("fmt", BLOCK_FORMATS)
def test_block_dequant_matches_quants_py_bit_exact(golden_dir, fmt):
    g = np.load(os.path.join(golden_dir, f"blocks_{fmt}.npz"))
    t = T.parse(fmt)
    assert (int(g["block"]), int(g["type_size"])) == (t.block, t.bytes)
    for raw_key, deq_key in (("qbytes", "deq"), ("raw", "raw_deq")):
        raw, want = g[raw_key], g[deq_key]
        got = oracle.dequant(t.id, raw, want.size).reshape(want.shape)
        assert np.array_equal(_bits(got), _bits(want)), f"{fmt}/{raw_key}: oracle dequant differs from quants.py"


def test_fp8_decode_matches_torch_and_types_h_rule(golden_dir):
    g = np.load(os.path.join(golden_dir, "fp8_codes.npz"))
    codes = np.arange(256, dtype=np.uint8)
    for name, t in (("e4m3", T.F8_E4M3), ("e5m2", T.F8_E5M2)):
        got = oracle.dequant(t.id, codes, 256)
        torch_vals = g[f"{name}_torch"]
        finite = np.isfinite(torch_vals)
        # every code torch decodes to a finite value: bit-identical (incl. -0.0 and subnormals)
        assert np.array_equal(_bits(got)[finite], _bits(torch_vals)[finite])
        # NaN/Inf codes: the reference decodes them as ordinary finite numbers (types.h:302-314)
        assert np.array_equal(_bits(got), _bits(g[f"{name}_ref"]))
        assert np.all(np.isfinite(got))
    assert (~np.isfinite(g["e4m3_torch"])).sum() == 2 and (~np.isfinite(g["e5m2_torch"])).sum() == 8
    assert oracle.dequant(T.F8_E4M3.id, np.array([0x7F], np.uint8), 1)[0] == 480.0
    assert oracle.dequant(T.F8_E5M2.id, np.array([0x7C], np.uint8), 1)[0] == 65536.0


def test_scalar_formats():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(512) * 0.3).astype(np.float32)
    assert np.array_equal(oracle.dequant(T.F32.id, x.view(np.uint8), 512), x)
    h = x.astype(np.float16)
    assert np.array_equal(oracle.dequant(T.F16.id, h.view(np.uint8), 512), h.astype(np.float32))
    bf = (x.view(np.uint32) >> 16).astype(np.uint16)
    assert np.array_equal(oracle.dequant(T.BF16.id, bf.view(np.uint8), 512), (bf.astype(np.uint32) << 16).view(np.float32))
    q = np.arange(-128, 128, dtype=np.int8)
    # types.h:423-424: (1.f/100.f) * q — NOT q / 100 (70 of 256 codes differ)
    want = np.float32(1.0) / np.float32(100.0) * q.astype(np.float32)
    assert np.array_equal(oracle.dequant(T.Q8.id, q.view(np.uint8), 256), want)
    assert (want != q.astype(np.float32) / np.float32(100.0)).sum() > 0
    u = np.arange(256, dtype=np.uint8)
    assert np.array_equal(oracle.dequant(T.QI8.id, u, 256), u.astype(np.float32) / np.float32(127.5) - np.float32(1.0))
    # e2m5 / e3m4 (types.h:317-318): same shift-and-scale rule, bias 1 / 3
    for t, E, M in ((T.F8_E2M5, 2, 5), (T.F8_E3M4, 3, 4)):
        got = oracle.dequant(t.id, u, 256)
        bias = (1 << (E - 1)) - 1
        e = (u & 0x7F) >> M
        m = (u & ((1 << M) - 1)).astype(np.float64)
        mag = np.where(e > 0, (1 + m / (1 << M)) * 2.0 ** (e.astype(np.float64) - bias), m / (1 << M) * 2.0 ** (1 - bias))
        want = np.where(u & 0x80, -mag, mag).astype(np.float32)
        assert np.array_equal(_bits(got), _bits(want))


def test_matmul_modes_agree_and_follow_definition():
    rng = np.random.default_rng(1)
    n, d = 256, 64
    w = (rng.standard_normal((d, n)) * 0.05).astype(np.float16)
    x = rng.standard_normal(n).astype(np.float32)
    ref = (w.astype(np.float64) @ x.astype(np.float64))
    for mode in (0, 1, 2):
        got = oracle.matmul(x, w.view(np.uint8), T.F16.id, n, d, mode)
        assert np.max(np.abs(got - ref)) < 1e-5
    with pytest.raises(ValueError):
        oracle.matmul(x[:48], w.view(np.uint8), T.F16.id, 48, d)       # n % 32 != 0 (infer.cpp:110)
    with pytest.raises(ValueError):
        oracle.matmul(x, w.view(np.uint8), T.U8.id, n, d)               # unsupported dtype (infer.cpp:211-214)


def test_rmsnorm_rope_softmax_definitions():
    rng = np.random.default_rng(2)
    x = rng.standard_normal(128).astype(np.float32)
    w = (1 + 0.1 * rng.standard_normal(128)).astype(np.float32)
    got = oracle.rmsnorm(x, w.view(np.uint8), T.F32.id, 1e-5)
    want = x / np.sqrt(np.mean(x.astype(np.float64) ** 2) + 1e-5) * w
    assert np.max(np.abs(got - want)) < 1e-5
    with pytest.raises(ValueError):
        oracle.rmsnorm(x, w.astype(np.float16).view(np.uint8), T.F16.id, 1e-5)   # F32/BF16 only (infer.cpp:248)
    # rope: interleaved pairs, freq = theta^-(j/rotary_dim), untouched beyond rotary_dim (infer.cpp:305-322)
    v = rng.standard_normal(64).astype(np.float32)
    r = oracle.rope(v, 32, 7, 10000.0, 16)
    for i in range(0, 64, 2):
        j = i % 32
        f = 0.0 if j >= 16 else 1.0 / 10000.0 ** (j / 16)
        c, s = np.cos(7 * f), np.sin(7 * f)
        assert abs(r[i] - (v[i] * c - v[i + 1] * s)) < 1e-5 and abs(r[i + 1] - (v[i] * s + v[i + 1] * c)) < 1e-5


def test_sampler_flt_min_quirk():
    # sampler.cpp:6,22 seed the running max with FLT_MIN (+1.18e-38), not lowest: all-negative logits -> argmax 0
    assert oracle.sample_argmax(np.array([-3.0, -1.0, -2.0], np.float32)) == 0
    assert oracle.sample_argmax(np.array([-3.0, 1.0, 2.0], np.float32)) == 2
    lg = np.array([0.5, 1.5, -0.5], np.float32)
    p = np.exp(lg - lg.max()); p /= p.sum()
    assert abs(oracle.sample_prob(lg, 1) - p[1]) < 1e-6
