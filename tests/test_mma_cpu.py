"""numpy model of the tensor-core decode matvec (xalm_b200/csrc/matvec_mma.cuh) and of its weight layout
(xalm_b200/csrc/frag_layout.cuh), run on the CPU.

The kernel feeds mma.sync.m16n8k32.u8.u8.s32 with the STORED weight bytes (quant + bias) as A and, as B, the three unsigned
limbs of the OFFSET activations X' = X + 2^23 plus a column of ones; the accumulators start at -bias * (limb sums).  This file
restates that arithmetic in integers and checks the claims of the kernel's header:
  * 65536 * v(l0', ones) + v(l1', l2') == sum_k (u_k - bias) * X_k exactly, for every bias the formats use;
  * every intermediate fits the int32 accumulators of the mma;
  * a one-hot x gives back d * q (+ m) bit for bit;
and for the layout: the (row, element) -> byte map of a record is a bijection that puts the mma's A fragment of lane t at
bytes [16 t, 16 t + 16) (8-bit) / [8 t, 8 t + 8) (4-bit), with the two rows of a GLU pair in the same lane.
The GPU side of the same checks is tests/test_gpu_formats.py (one-hot bit exactness, random x against the oracle).
"""
import numpy as np
import pytest

F32 = np.float32


def stage_block(x):
    """xm_store_group8 over one 32-element block: offset limbs l0' l1' l2' (each 0..255), their sums, dx."""
    x = np.asarray(x, dtype=F32)
    mb = int(np.max(x.view(np.uint32) & 0x7FFFFFFF))
    eb = mb >> 23
    live = 40 <= eb < 255
    sc = np.array([(276 - eb) << 23], np.uint32).view(F32)[0] if live else F32(0)
    dx = np.array([(eb - 22) << 23], np.uint32).view(F32)[0] if live else F32(0)
    with np.errstate(invalid="ignore"):
        prod = np.nan_to_num((x * sc).astype(F32), nan=0.0)            # __float2int_rn(NaN) = 0 (inf * 0 in a dropped block)
    X = np.minimum(np.rint(prod).astype(np.int64), 8388607)
    Xo = X + 8388608
    assert Xo.min() >= 0 and Xo.max() < 1 << 24
    limbs = np.stack([(Xo >> 16) & 255, (Xo >> 8) & 255, Xo & 255])
    return X, limbs, limbs.sum(axis=1), dx


def mma_block(u, bias, limbs, sums):
    """One record column of the mma for one row: accumulators of the four used B columns -> (v01, v23)."""
    u = u.astype(np.int64)
    c_l2 = -bias * sums[2] + int(np.dot(u, limbs[2]))
    c_l1 = -bias * sums[1] + int(np.dot(u, limbs[1]))
    c_l0 = -bias * (sums[0] - 4096) + int(np.dot(u, limbs[0]))
    c_one = int(u.sum())
    for c in (c_l2, c_l1, c_l0, c_one):
        assert abs(c) < 2 ** 31
    v01 = 256 * c_l1 + c_l2          # lane tig = 0: columns (l2', l1')
    v23 = c_l0 - 128 * c_one         # lane tig = 1: columns (l0', ones); 128 * 65536 = 2^23 removes the offset
    assert abs(v01) < 2 ** 31 and abs(v23) < 2 ** 31
    return v01, v23


@pytest.mark.parametrize("bias,qmax", [(128, 255), (8, 15), (16, 31), (0, 15), (0, 31)])
@pytest.mark.parametrize("scale", [1e-12, 1.0, 250.0])
def test_offset_limb_mma_is_the_exact_integer_dot_product(bias, qmax, scale):
    rng = np.random.default_rng(bias * 7 + qmax)
    for trial in range(40):
        x = (rng.standard_normal(32) * scale).astype(F32)
        if trial == 0:
            x[:] = np.abs(x).max()                      # every element at the block maximum: the largest sums
        if trial == 1:
            x[:] = -np.abs(x).max()
        X, limbs, sums, dx = stage_block(x)
        u = rng.integers(0, qmax + 1, 32)
        if trial < 2:
            u[:] = qmax
        v01, v23 = mma_block(u, bias, limbs, sums)
        assert 65536 * v23 + v01 == int(np.dot(u.astype(np.int64) - bias, X))


def test_dead_and_nonfinite_blocks_contribute_nothing():
    for x in (np.zeros(32, F32), np.full(32, 1e-30, F32), np.array([np.inf] + [1.0] * 31, F32)):
        X, limbs, sums, dx = stage_block(x)
        assert dx == 0 and np.all(X == 0)
        v01, v23 = mma_block(np.full(32, 200), 128, limbs, sums)
        assert v01 == 0 and v23 == 0                    # and dx = 0 multiplies whatever is left


@pytest.mark.parametrize("bias,qmax,has_min", [(128, 255, False), (8, 15, False), (16, 31, False), (0, 15, True), (0, 31, True)])
def test_one_hot_returns_the_dequantised_weight_bit_for_bit(bias, qmax, has_min):
    rng = np.random.default_rng(3)
    for _ in range(200):
        j = int(rng.integers(0, 32))
        x = np.zeros(32, F32)
        x[j] = 1.0
        X, limbs, sums, dx = stage_block(x)
        u = rng.integers(0, qmax + 1, 32)
        d = F32(np.float16(rng.standard_normal() * 0.1))
        m = F32(np.float16(rng.standard_normal())) if has_min else None
        v01, v23 = mma_block(u, bias, limbs, sums)
        # the two lanes accumulate float(v) * (d * dx [* 65536]) separately, the quad reduction adds them
        y1 = F32(np.float64(F32(v23)) * np.float64(F32(d * F32(dx * F32(65536.0)))))
        y0 = F32(np.float64(F32(v01)) * np.float64(F32(d * dx)))
        if has_min:
            tot = (int(sums[0]) - 4096) * 65536 + int(sums[1]) * 256 + int(sums[2])
            sx = F32(F32(tot) * dx)
            assert sx == 1.0
            y0 = F32(np.float64(m) * np.float64(sx) + np.float64(y0))
        y = F32(y0 + y1)
        want = F32(F32(d * F32(int(u[j]) - bias)) + (m if has_min else F32(0)))   # quants.py: d * q (+ m), one rounding
        assert y.view(np.uint32) == want.view(np.uint32) or (y == 0 and want == 0)


# ---- layout (frag_layout.cuh) -------------------------------------------------------------------------------------------
def rec_offset_8bit(row, kk):
    g, hi = row % 8, row // 8
    tig, e, j = (kk % 16) // 4, kk % 4, (kk // 16) * 2 + hi
    return (g * 4 + tig) * 16 + j * 4 + e


def rec_offset_4bit(row, kk):
    g, hi = row % 8, row // 8
    tig, e, half = (kk % 16) // 4, kk % 4, kk // 16
    return (g * 4 + tig) * 8 + half * 4 + e, hi     # (byte, nibble)


def test_record_layout_is_the_mma_a_fragment():
    seen8, seen4 = set(), set()
    for row in range(16):
        for kk in range(32):
            o = rec_offset_8bit(row, kk)
            assert 0 <= o < 512 and o not in seen8
            seen8.add(o)
            b, nib = rec_offset_4bit(row, kk)
            assert 0 <= b < 256 and (b, nib) not in seen4
            seen4.add((b, nib))
            # PTX m16n8k32 A fragment: byte i of lane t = 4 * groupID + threadID_in_group holds
            #   row = groupID (+ 8 for i in 4..7, 12..15), col = 4 * threadID_in_group + (i & 3) (+ 16 for i >= 8)
            t, i = o // 16, o % 16
            assert row == t // 4 + (8 if (i // 4) % 2 else 0) and kk == 4 * (t % 4) + (i & 3) + (16 if i >= 8 else 0)
            t4, i4 = b // 8, b % 8
            assert row % 8 == t4 // 4 and row // 8 == nib and kk == 4 * (t4 % 4) + (i4 & 3) + (16 if i4 >= 4 else 0)
    assert len(seen8) == 512 and len(seen4) == 512


def frag_virtual_row(glu_half, r):
    if not glu_half:
        return r
    o = r if r < glu_half else r - glu_half
    return (o // 8) * 16 + (0 if r < glu_half else 8) + o % 8


def test_glu_pairs_share_a_lane():
    H = 64
    vr = [frag_virtual_row(H, r) for r in range(2 * H)]
    assert sorted(vr) == list(range(2 * H))                       # a permutation of the stacked gate|up rows
    for o in range(H):
        a, b = vr[o], vr[H + o]
        assert a // 16 == b // 16 and b - a == 8                  # same tile, rows g and g + 8: registers (c0,c1) and (c2,c3) of one lane
        assert a // 16 * 8 + a % 8 == o                           # the epilogue's output index row0 / 2 + lane
