"""Op-level parity (the functions model.h:286-316 exposes "for tests"): rmsnorm, rope, mha, ffn vs the CPU oracle."""
import numpy as np
import pytest

from oracle import oracle
from xalm_b200 import capi, synth
from xalm_b200 import types as T

from gpu_util import random_raw

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [32, 128, 4096, 8192])
def test_rmsnorm(size):
    rng = np.random.default_rng(size)
    x = rng.standard_normal(size).astype(np.float32) * 3
    w = (1 + 0.2 * rng.standard_normal(size)).astype(np.float32)
    wb = synth.quantize(T.BF16, w)
    for wt, raw in ((T.F32, w.view(np.uint8)), (T.BF16, wb.view(np.uint8))):
        got = capi.rmsnorm(x, raw, wt.id, 1e-5)
        want = oracle.rmsnorm(x, raw, wt.id, 1e-5)
        assert np.max(np.abs(got - want)) <= 2e-6 * np.abs(want).max()


@pytest.mark.parametrize("pos", [0, 1, 7, 4095, 32767, 100000])
def test_rope(pos):
    rng = np.random.default_rng(pos)
    for head_dim, rotary_dim, theta in ((128, 128, 1e6), (64, 32, 10000.0), (32, 32, 5e5)):
        v = rng.standard_normal(4 * head_dim).astype(np.float32)
        got = capi.rope(v, head_dim, pos, theta, rotary_dim)
        want = oracle.rope(v, head_dim, pos, theta, rotary_dim)
        # same host-side powf table; device sincosf vs glibc cosf/sinf: a few ulp of |v|
        assert np.max(np.abs(got - want)) <= 4e-6 * np.abs(v).max()
        if rotary_dim < head_dim:   # untouched beyond rotary_dim (infer.cpp:310-312)
            idx = np.arange(v.size) % head_dim >= rotary_dim
            assert np.array_equal(got[idx], v[idx])


MHA_CASES = [
    # head_dim, n_heads, n_kv_heads, max_seq_len, kv_len
    (128, 32, 8, 4096, 1), (128, 32, 8, 4096, 2), (128, 32, 8, 4096, 127), (128, 32, 8, 4096, 129),
    (128, 32, 8, 4096, 1000), (128, 32, 8, 4096, 4096), (128, 8, 1, 2048, 2048), (128, 64, 8, 512, 300),
    (64, 4, 2, 128, 77), (32, 4, 2, 64, 64), (32, 4, 4, 64, 5), (256, 2, 1, 96, 96), (64, 8, 1, 300, 299),
]


@pytest.mark.parametrize("case", MHA_CASES, ids=lambda c: "hd%d_h%d_kv%d_T%d_len%d" % c)
def test_mha_vs_oracle(case):
    hd, nh, nkv, T_, kv_len = case
    rng = np.random.default_rng(kv_len * 7 + hd)
    q = rng.standard_normal(nh * hd).astype(np.float32)
    kb = (rng.standard_normal((T_, nkv * hd)) * 0.7).astype(np.float16)
    vb = rng.standard_normal((T_, nkv * hd)).astype(np.float16)
    kb[kv_len:] = np.float16(np.nan)     # slots past kv_len must never be read
    vb[kv_len:] = np.float16(np.nan)
    got, att = capi.mha(q, kb.view(np.uint16), vb.view(np.uint16), hd, kv_len, T_, nh, nkv, want_att=True)
    want, watt = oracle.mha(q, kb.view(np.uint16), vb.view(np.uint16), hd, kv_len, T_, nh, nkv)
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - want)) <= 1e-5
    att, watt = att.reshape(nh, T_)[:, :kv_len], watt.reshape(nh, T_)[:, :kv_len]
    assert np.max(np.abs(att - watt)) <= 1e-6
    # deterministic: same input, same bits
    again = capi.mha(q, kb.view(np.uint16), vb.view(np.uint16), hd, kv_len, T_, nh, nkv)
    assert np.array_equal(got, again)


def test_mha_large_scores_are_stable():
    hd, nh, nkv, T_ = 128, 8, 2, 512
    rng = np.random.default_rng(1)
    q = (rng.standard_normal(nh * hd) * 30).astype(np.float32)
    kb = (rng.standard_normal((T_, nkv * hd)) * 30).astype(np.float16)
    vb = rng.standard_normal((T_, nkv * hd)).astype(np.float16)
    got = capi.mha(q, kb.view(np.uint16), vb.view(np.uint16), hd, T_, T_, nh, nkv)
    want, _ = oracle.mha(q, kb.view(np.uint16), vb.view(np.uint16), hd, T_, T_, nh, nkv)
    assert np.all(np.isfinite(got)) and np.max(np.abs(got - want)) <= 1e-4


@pytest.mark.parametrize("t", [T.F16, T.Q8_0, T.Q4_0, T.F8_E4M3, T.Q5_1], ids=lambda t: t.name)
@pytest.mark.parametrize("act", [capi.GELU, capi.SILU])
def test_ffn_vs_oracle(t, act):
    dim, hidden = 256, 768
    w1, w3 = random_raw(t, hidden, dim, 1), random_raw(t, hidden, dim, 3)
    w2 = random_raw(t, dim, hidden, 2)
    x = np.random.default_rng(9).standard_normal(dim).astype(np.float32)
    got = capi.ffn(x, w1, w2, w3, t.id, hidden, dim, act)
    L = oracle.lib()
    a = oracle.matmul(x, w1, t.id, dim, hidden, 2)
    b = oracle.matmul(x, w3, t.id, dim, hidden, 2)
    f = L.orc_gelu if act == capi.GELU else L.orc_silu
    hb = np.array([f(float(v)) for v in a], dtype=np.float32) * b
    want = oracle.matmul(hb, w2, t.id, hidden, dim, 2)
    assert np.max(np.abs(got - want)) <= 2e-5 * max(1.0, np.abs(want).max())
