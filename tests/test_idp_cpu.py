"""numpy model of the integer-dot core of the decode matvec (xalm_b200/csrc/idp.cuh), run on the CPU.

The kernel holds each 32-element activation block as one power-of-two scale dx and three int8 limbs per element
(xq_store_block), multiplies limbs with weights in integers (dp4a) and recombines (idp_combine).  This file restates
that arithmetic step by step and checks the three claims the header of idp.cuh makes:
  * the limbs reconstruct round(x / dx) exactly, and |x - dx * X| <= dx / 2 <= 2^-23 of the block maximum;
  * a one-hot x returns the dequantised weight bit for bit (what the per-format GPU parity tests rely on);
  * against an fp64 dot product the result is no worse than the reference's sequential fp32 loop (infer.cpp:104-135).
The GPU side of the same checks is tests/test_gpu_formats.py.
"""
import numpy as np
import pytest

F32 = np.float32


def xq_block(x):
    """xq_store_block: x (32 fp32) -> (limbs [3, 32] as int64 values, limb sums [3], dx)."""
    x = np.asarray(x, dtype=F32)
    mb = int(np.max(x.view(np.uint32) & 0x7FFFFFFF))
    eb = mb >> 23
    live = 40 <= eb < 255
    if not live:
        return np.zeros((3, 32), np.int64), np.zeros(3, np.int64), F32(0)
    sc = np.array([(276 - eb) << 23], np.uint32).view(F32)[0]
    dx = np.array([(eb - 22) << 23], np.uint32).view(F32)[0]
    X = np.minimum(np.rint((x * sc).astype(F32)).astype(np.int64), 8388607)   # __float2int_rn, ties to even like np.rint
    u = X & 0xFFFFFF                                                         # 24-bit two's complement
    l2, l1, l0 = u & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF
    l0 = np.where(l0 >= 128, l0 - 256, l0)                                   # the top limb is the signed one
    limbs = np.stack([l0, l1, l2])
    return limbs, limbs.sum(axis=1), dx


def idp_block(q_stored, bias, d, limbs, sums, dx, m=None):
    """IdpFmt<...>::block for one 32-weight block: stored (unsigned) quants, bias folded through the limb sums."""
    acc = [int(-bias * sums[k] + np.dot(q_stored.astype(np.int64), limbs[k])) for k in range(3)]
    assert all(abs(a) < 2 ** 31 for a in acc)                                # int32 accumulators in the kernel
    low = F32((acc[1] << 8) + acc[2])                                        # __int2float_rn of the 29-bit sum
    comb = F32(np.float64(F32(acc[0])) * 65536.0 + np.float64(low))          # fmaf: one rounding
    y = F32(np.float64(comb) * np.float64(F32(d * dx)))                      # fmaf(comb, d*dx, 0)
    if m is not None:
        sx = F32(F32(np.float64(F32(sums[0])) * 65536.0 + np.float64(F32((int(sums[1]) << 8) + int(sums[2])))) * dx)
        y = F32(np.float64(F32(m)) * np.float64(sx) + np.float64(y))
    return y


@pytest.mark.parametrize("scale", [1e-20, 1e-3, 1.0, 37.5, 1e20])
def test_limbs_reconstruct_the_rounded_activation(scale):
    rng = np.random.default_rng(1)
    for _ in range(50):
        x = (rng.standard_normal(32) * scale).astype(F32)
        limbs, sums, dx = xq_block(x)
        X = 65536 * limbs[0] + 256 * limbs[1] + limbs[2]
        assert limbs[0].min() >= -128 and limbs[0].max() <= 127 and limbs[1:].min() >= 0 and limbs[1:].max() <= 255
        err = np.abs(x.astype(np.float64) - np.float64(dx) * X)
        assert err.max() <= np.float64(dx) / 2 * (1 + 1e-12)
        assert np.float64(dx) / 2 <= 2.0 ** -23 * np.abs(x).max()            # dx = 2^(eb-149), max >= 2^(eb-127)
        assert np.array_equal(sums, limbs.sum(axis=1))


def test_dead_and_non_finite_blocks_are_dropped():
    for x in (np.zeros(32, F32), np.full(32, 1e-40, F32), np.array([np.inf] + [1.0] * 31, F32), np.array([np.nan] + [1.0] * 31, F32)):
        limbs, sums, dx = xq_block(x)
        assert dx == 0 and not limbs.any()


@pytest.mark.parametrize("bias,levels", [(128, 256), (8, 16), (16, 32), (0, 16), (0, 32)])
def test_one_hot_returns_the_dequantised_weight_bit_for_bit(bias, levels):
    rng = np.random.default_rng(2)
    for _ in range(40):
        q = rng.integers(0, levels, size=32)
        d = F32(np.float16(rng.uniform(1e-4, 0.2)))
        m = F32(np.float16(rng.uniform(-1, 1))) if bias == 0 else None
        e = int(rng.integers(0, 32))
        x = np.zeros(32, F32)
        x[e] = 1.0
        y = idp_block(q, bias, d, *xq_block(x), m=m)
        want = F32(d * F32(q[e] - bias)) if m is None else F32(np.float64(d) * float(q[e]) + np.float64(m))   # quants.py: d*q (+m), one rounding
        assert y.tobytes() == want.tobytes()


def test_closer_to_fp64_than_sequential_fp32():
    """A 4096-long row (128 blocks): integer block sums + one fp32 accumulation per block vs the reference's element loop."""
    rng = np.random.default_rng(3)
    worse = 0
    e_idp, e_seq = [], []
    for trial in range(20):
        n = 4096
        q = rng.integers(0, 256, size=n)
        d = np.float16(rng.uniform(1e-3, 2e-2, size=n // 32)).astype(F32)
        x = rng.standard_normal(n).astype(F32)
        w = (np.repeat(d, 32) * (q - 128).astype(F32)).astype(F32)
        exact = float(np.dot(w.astype(np.float64), x.astype(np.float64)))
        y = F32(0)
        for b in range(n // 32):
            sl = slice(32 * b, 32 * b + 32)
            y = F32(y + idp_block(q[sl], 128, d[b], *xq_block(x[sl])))
        seq = F32(0)
        for i in range(n):
            seq = F32(seq + F32(w[i] * x[i]))
        e_idp.append(abs(float(y) - exact))
        e_seq.append(abs(float(seq) - exact))
        worse += e_idp[-1] > e_seq[-1]
    assert np.mean(e_idp) < np.mean(e_seq), (np.mean(e_idp), np.mean(e_seq))
    assert np.max(e_idp) < 1e-4 * 4096 ** 0.5                                # rows of unit-variance x, |w| <= 2.6


def test_accumulators_stay_inside_int32():
    # worst case: all weights 255 (stored), all limbs at their extremes
    limbs = np.stack([np.full(32, -128), np.full(32, 255), np.full(32, 255)])
    sums = limbs.sum(axis=1)
    for bias in (0, 8, 16, 128):
        for k in range(3):
            assert abs(-bias * sums[k] + 255 * limbs[k].sum()) < 2 ** 21
