"""C++ host side without a device: checkpoint verification (xxh3 of every tensor against the header, the check the reference
parses and never performs — xalm.h:90-192, convert.py:265-266) and Tensor::convert_to for the block formats (tensor.cpp:226-296
only knows the scalar pairs; the block quantisers are quants.py's, byte for byte)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from xalm_b200 import build as B
from xalm_b200 import synth
from xalm_b200 import types as T
from xalm_b200 import xalm_file as X

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def xalm_main():
    if not os.path.exists(B.MAIN):
        B.build_host()
    return B.MAIN


def run(exe, *args):
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=120)


@pytest.mark.parametrize("name", ["tiny_f16", "tiny_bf16", "tiny_q8_0", "tiny_q4_0", "tiny_f8_e4m3"])
def test_verify_accepts_convert_py_checkpoints(xalm_main, name):
    r = run(xalm_main, os.path.join(GOLDEN, f"{name}.xalm"), "-m", "verify")
    assert r.returncode == 0, r.stderr
    assert "verified 22 tensors (22 with a hash)" in r.stdout and "NOT checked" not in r.stdout


def test_verify_rejects_a_flipped_payload_bit(xalm_main, tmp_path):
    src = os.path.join(GOLDEN, "tiny_q8_0.xalm")
    f = X.XalmFile(src)
    ti = f.tensors["l.1.mlp.gate.weight"]
    f.close()
    bad = tmp_path / "bad.xalm"
    shutil.copy(src, bad)
    with open(bad, "r+b") as fh:
        fh.seek(ti.offset + ti.size // 2)
        b = fh.read(1)
        fh.seek(ti.offset + ti.size // 2)
        fh.write(bytes([b[0] ^ 0x10]))
    r = run(xalm_main, str(bad), "-m", "verify")
    assert r.returncode == 1 and "hash mismatch for tensor l.1.mlp.gate.weight" in r.stderr
    # the Python mirror agrees
    with pytest.raises(ValueError):
        X.XalmFile(str(bad)).verify_hashes()


@pytest.mark.parametrize("target", ["q8_0", "q4_0", "q4_1", "q5_0", "q5_1", "bf16", "f32"])
def test_convert_to_block_formats_matches_the_python_quantiser(xalm_main, target):
    import xxhash
    path = os.path.join(GOLDEN, "tiny_f16.xalm")
    r = run(xalm_main, path, "-m", "verify", "-t", target)
    assert r.returncode == 0, r.stderr
    got = {}
    for line in r.stdout.splitlines():
        p = line.split()
        if len(p) == 4 and p[1] == target.upper():
            got[p[0]] = (int(p[2]), int(p[3]))
    f = X.XalmFile(path)
    t = T.parse(target)
    n = 0
    for name, ti in f.tensors.items():
        if len(ti.shape) != 2 or ti.type is not T.F16:
            continue
        x = f.raw(name).view(np.float16).astype(np.float32).reshape(ti.shape)
        want = np.ascontiguousarray(synth.quantize(t, x)).view(np.uint8).reshape(-1)    # pinned to quants.py by test_quantize.py
        assert got[name] == (want.size, xxhash.xxh3_64(want.data).intdigest()), name
        n += 1
    assert n >= 15
    f.close()
