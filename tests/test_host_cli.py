"""The C++ host executable (csrc/host: the reference's CLI surface over the C ABI)."""
import os
import subprocess

import numpy as np
import pytest

from xalm_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _main():
    build.build_cuda()
    build.build_host()
    return build.MAIN


def _has_gpu():
    from xalm_b200 import capi
    try:
        return capi.device_count() > 0
    except capi.XalmError:
        return False


def test_cli_usage_and_device_switch():
    exe = _main()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr
    r = subprocess.run([exe, os.path.join(GOLD, "tiny_f16.xalm"), "-d", "cpu"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU forward" in r.stderr          # no CPU fallback in this backend
    r = subprocess.run([exe, os.path.join(GOLD, "tiny_f16.xalm"), "-m", "nonsense"], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr
    r = subprocess.run([exe, "/nonexistent.xalm"], capture_output=True, text=True)
    assert r.returncode == 1


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a GPU")
def test_cli_fails_loudly_without_a_gpu():
    r = subprocess.run([_main(), os.path.join(GOLD, "tiny_q8_0.xalm"), "-n", "4"], capture_output=True, text=True)
    assert r.returncode == 1 and "model.cuda()" in r.stderr


def _oracle_completion(path, prompt, n, context=0):
    from oracle import oracle
    from xalm_b200 import xalm_file as X
    from xalm_b200.model import Tokenizer
    f = X.XalmFile(path)
    cfg = X.parse_config(f.metadata, context)
    om = oracle.OracleModel(cfg, {k: (ti.type.id, f.raw(k)) for k, ti in f.tensors.items()})
    tok = Tokenizer(f)
    enc = tok.encode(prompt, True)
    om.forward(0, 0, 1)
    lg = None
    for pos, t in enumerate(enc):
        lg = om.forward(t, pos, 1 if pos + 1 == len(enc) else 0)
    text = b""
    logits = []
    for _ in range(n):
        t = oracle.sample_argmax(lg)
        text += tok.decode_one(enc[-1], t)
        enc.append(t)
        if t in (tok.eos_id, tok.eot_id):
            break
        lg = om.forward(t, len(enc) - 1, 1)
    return enc, text, om, tok, cfg, f


@pytest.mark.gpu
@pytest.mark.parametrize("ckpt,context", [("tiny_q8_0.xalm", 0), ("tiny_f16.xalm", 0), ("tiny_q4_0.xalm", 16)])
def test_cli_completion_matches_oracle(ckpt, context):
    path = os.path.join(GOLD, ckpt)
    prompt = "Q: What is the meaning of life? A:"
    enc, text, om, tok, cfg, f = _oracle_completion(path, prompt, 24, context)
    args = [_main(), path, "-d", "cuda", "-m", "completion", "-n", "24", "-i", prompt]
    if context:
        args += ["-T", str(context)]
    r = subprocess.run(args, capture_output=True)
    assert r.returncode == 0, r.stderr.decode(errors="replace")
    out = r.stdout
    assert b"Using CUDA" in out and b"Generation stats:" in out and b"bandwidth:" in out
    # the generated text sits between the encoding debug line / stats and must equal the oracle's greedy decode
    assert text in out, (text, out[-600:])
    assert f"{len(enc)} tokens".encode() in out


@pytest.mark.gpu
def test_cli_perplexity_matches_oracle():
    from oracle import oracle
    path = os.path.join(GOLD, "tiny_bf16.xalm")
    prompt = "the meaning of life is in the the of a"
    enc, _, om, tok, cfg, f = _oracle_completion(path, prompt, 0)
    enc = tok.encode(prompt, True)
    s = 0.0
    for pos in range(len(enc) - 1):
        lg = om.forward(enc[pos], pos, 1)
        s += np.log(oracle.sample_prob(lg, enc[pos + 1]))
    ppl = float(np.exp(-s / (len(enc) - 1)))
    r = subprocess.run([_main(), path, "-m", "perp", "-i", prompt], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    line = [l for l in r.stdout.splitlines() if "perplexity:" in l][0]
    got = float(line.split("perplexity:")[1].split()[0])
    assert abs(got - ppl) / ppl < 2e-3, (got, ppl)


@pytest.mark.gpu
def test_cli_passkey_mode_runs_past_the_context_window():
    r = subprocess.run([_main(), os.path.join(GOLD, "tiny_f16.xalm"), "-m", "passkey", "-n", "3", "-T", "32", "-l", "1"],
                       capture_output=True)          # bytes: byte-fallback tokens print raw bytes
    assert r.returncode == 0, r.stderr
    assert b"Passkey test:" in r.stdout and b"What is the pass key?" in r.stdout
