"""Host-side format writers (csrc/host/quantize.cpp) against what the reference's quants.py produced
(tests/golden/blocks_*.npz: `x` -> `qbytes`), byte for byte.  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle
from xalm_b200 import synth
from xalm_b200 import types as T


@pytest.mark.parametrize("fmt", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0", "tq1_0"])
def test_quantizer_is_byte_identical_to_quants_py(golden_dir, fmt):
    g = np.load(os.path.join(golden_dir, f"blocks_{fmt}.npz"))
    got = synth.quantize(T.parse(fmt), g["x"])
    assert got.dtype == np.uint8 and got.shape == g["qbytes"].shape
    assert np.array_equal(got, g["qbytes"])


def test_scalar_writers():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((4, 64)) * 0.7).astype(np.float32)
    assert np.array_equal(synth.quantize(T.F16, x).view(np.float16), x.astype(np.float16))
    bf = synth.quantize(T.BF16, x)
    u = x.view(np.uint32).astype(np.uint64)
    want = ((u + (0x7FFF + ((u >> 16) & 1))) >> 16).astype(np.uint16)          # quants.py:265-274
    assert np.array_equal(bf, want)
    q8 = synth.quantize(T.Q8, x).view(np.int8)                                  # types.h:458-463
    assert np.array_equal(q8, np.clip(np.round(x * np.float32(100)), -128, 127).astype(np.int8))
    import torch
    assert np.array_equal(synth.quantize(T.F8_E4M3, x), torch.from_numpy(x).to(torch.float8_e4m3fn).view(torch.uint8).numpy())


def test_normal_generator_is_seeded_and_gaussian():
    a = synth.normal(7, 11, 200001, 0.02)
    b = synth.normal(7, 11, 200001, 0.02)
    assert np.array_equal(a, b)
    assert not np.array_equal(a[:1000], synth.normal(7, 12, 1000, 0.02))
    assert abs(a.mean()) < 3e-4 and abs(a.std() - 0.02) < 3e-4
    # prefix property: element i depends only on (seed, stream, i)
    assert np.array_equal(a[:1000], synth.normal(7, 11, 1000, 0.02))


def test_synthetic_checkpoint_roundtrip(tmp_path):
    from xalm_b200 import xalm_file as X
    p = str(tmp_path / "tiny.q8_0.xalm")
    c = synth.write_checkpoint(p, "tiny", "q8_0", seed=5)
    f = X.XalmFile(p)
    cfg = X.parse_config(f.metadata)
    assert cfg["dim"] == c["dim"] and cfg["max_seq_len"] == 128 and cfg["act"] == 1
    f.verify_hashes()
    for name, shape in X.expected_tensors(cfg).items():
        assert f.tensors[name].shape == shape
    assert f.tensors["l.1.mlp.down.weight"].type is T.Q8_0 and f.tensors["output.norm.weight"].type is T.F32
    # dequantised weights look like N(0, 0.02^2)
    w = oracle.dequant(T.Q8_0.id, f.raw("l.0.attn.q.weight"), 256 * 256)
    assert abs(w.std() - 0.02) < 2e-3
    assert len(f.tokens()) == cfg["vocab_size"] and f.tokens()[3 + 65] == b"<0x41>"
    f.close()
