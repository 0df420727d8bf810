"""Batched prefill / perplexity on the tcgen05 path (BASELINE config 3) against the oracle and the token-at-a-time path.

* the GEMM alone, every weight format: out = a . W^T with W bit-exactly dequantised (capi.dequant is pinned to the oracle
  in test_gpu_formats.py) and both operands rounded to fp16 exactly as the kernel rounds them -> tight tolerance;
* the whole forward: logits of every position vs the CPU oracle run token by token (north star: max-abs 1e-2),
  KV cache contents, perplexity probabilities (Sampler::sample_prob), chunked prefill, prefill followed by decode.
"""
import numpy as np
import pytest

from oracle import oracle
from xalm_b200 import capi, synth
from xalm_b200 import types as T
from xalm_b200.model import InferenceState, Sampler

from gpu_util import LOGIT_TOL, random_raw, synth_pair

pytestmark = pytest.mark.gpu

ALL = ["f32", "f16", "bf16", "f8_e2m5", "f8_e3m4", "f8_e4m3", "f8_e5m2", "q8", "qi8", "q8_0", "q4_0", "q4_1", "q5_0", "q5_1", "tq1_0"]


def _fp16(x):
    return x.astype(np.float16).astype(np.float32)


@pytest.mark.parametrize("wtype", ALL)
@pytest.mark.parametrize("split", [1, 2, 3])
def test_gemm_every_format(wtype, split):
    t = T.parse(wtype)
    Tn, K, N = 150, 512, 320           # ragged: T not a multiple of 128, N not a multiple of 256; K = 2 units (TMA layout) for block formats
    raw = random_raw(t, N, K, seed=11)
    w = capi.dequant(t.id, raw, N * K).reshape(N, K)
    a = synth.normal(5, 77, Tn * K, 1.0).reshape(Tn, K)
    out = capi.gemm(a, raw, t.id, K, N, split)
    if split == 3:   # hi+lo on both operands: fp32-grade products (what is dropped: lo x lo ~ 2^-22 and the lo planes' own rounding)
        exact = a.astype(np.float64) @ w.astype(np.float64).T
        assert np.max(np.abs(out - exact)) <= 3e-6 * float(np.abs(exact).max()) + 1e-7, f"{wtype}: {np.max(np.abs(out - exact))}"
        return
    wh = _fp16(w).astype(np.float64)
    if split == 1:
        ref = _fp16(a).astype(np.float64) @ wh.T
    else:
        hi = _fp16(a)
        ref = (hi.astype(np.float64) + _fp16(a - hi).astype(np.float64)) @ wh.T
    scale = float(np.abs(ref).max())
    assert np.max(np.abs(out - ref)) <= 2e-5 * scale + 1e-6, f"{wtype}: {np.max(np.abs(out - ref))} vs scale {scale}"
    if split == 2:   # the hi+lo pair keeps the activations to ~fp32: what is left is the fp16 rounding of the weights
        exact = a.astype(np.float64) @ wh.T
        assert np.max(np.abs(out - exact)) <= 2e-5 * scale + 1e-6


@pytest.mark.parametrize("shape", [(1, 256, 256), (128, 64, 256), (129, 96, 32), (300, 2816, 1024)])
def test_gemm_shapes(shape):
    Tn, K, N = shape
    t = T.parse("f16")
    raw = random_raw(t, N, K, seed=3)
    w = capi.dequant(t.id, raw, N * K).reshape(N, K)
    a = synth.normal(6, 78, Tn * K, 1.0).reshape(Tn, K)
    out = capi.gemm(a, raw, t.id, K, N, 1)
    ref = _fp16(a).astype(np.float64) @ w.astype(np.float64).T
    assert np.max(np.abs(out - ref)) <= 2e-5 * float(np.abs(ref).max()) + 1e-6


def _oracle_all_logits(om, tokens):
    return np.stack([om.forward(int(tok), pos, 1).copy() for pos, tok in enumerate(tokens)])


@pytest.mark.parametrize("wtype,shape,n", [("f16", "tiny", 100), ("q8_0", "tiny", 128), ("q4_0", "tiny", 37), ("q5_1", "tiny", 70),
                                           ("q8_0", "small", 200), ("f16", "small", 130)])
@pytest.mark.parametrize("split", [1, 2, 3])
def test_prefill_logits_match_oracle(wtype, shape, n, split):
    capi.tune("prefill_split", split)
    try:
        config, om, gm = synth_pair(shape, wtype, seed=7, std=0.06 if shape == "tiny" else 0.03)
        rng = np.random.default_rng(1)
        tokens = rng.integers(3, config["vocab_size"], size=n).astype(np.int32)
        ref = _oracle_all_logits(om, tokens)
        got = gm.prefill(tokens, 0, want_logits=2)
        diff = float(np.max(np.abs(got - ref)))
        assert diff <= (LOGIT_TOL if split < 3 else 1e-3), f"{wtype}/{shape}: prefill logits differ from the oracle by {diff}"
        # the KV cache the prefill leaves behind is the one the token loop would have written (fp16-operand rounding aside)
        for layer in range(config["n_layers"]):
            for which in (0, 1):
                kvd = config["n_kv_heads"] * config["head_dim"]
                a = gm.read_kv(layer, which).view(np.float16).astype(np.float32)[: n * kvd]
                b = om.kv(layer, which).view(np.float16).astype(np.float32)[: n * kvd]
                assert np.max(np.abs(a - b)) <= 1e-2 * max(1.0, float(np.abs(b).max()))
        gm.close(); om.close()
    finally:
        capi.tune("prefill_split", 3)


def test_prefill_then_decode_and_last_logits():
    """Prompt hydrate through prefill (want 1 = last position only), then the greedy loop on the decode kernels."""
    config, om, gm = synth_pair("tiny", "q8_0", seed=9, std=0.06)
    prompt = [1] + list(np.random.default_rng(2).integers(3, config["vocab_size"], size=60))
    for pos, tok in enumerate(prompt):
        lg_o = om.forward(int(tok), pos, 1 if pos + 1 == len(prompt) else 0)
    lg = gm.prefill(np.array(prompt, dtype=np.int32), 0, want_logits=1)
    assert np.max(np.abs(lg - lg_o)) <= LOGIT_TOL
    state, sampler = InferenceState(config), Sampler(config)
    state.logits()[:] = lg
    otoks, gtoks = [], []
    for i in range(16):
        to, tg = oracle.sample_argmax(lg_o), sampler.sample_argmax(state)
        otoks.append(to); gtoks.append(tg)
        pos = len(prompt) + i
        lg_o = om.forward(to, pos, 1)
        gm.forward(state, tg, pos, 1)
        assert np.max(np.abs(state.logits() - lg_o)) <= LOGIT_TOL
    assert otoks == gtoks
    gm.close(); om.close()


@pytest.mark.parametrize("shape,n,cut", [("tiny", 120, 50), ("small", 300, 130)])   # head_dim 64: mma.sync attention; 128: tcgen05 attention
def test_chunked_prefill_and_hydrate_mode(shape, n, cut):
    config, om, gm = synth_pair(shape, "f16", seed=10, std=0.06 if shape == "tiny" else 0.03)
    tokens = np.random.default_rng(3).integers(3, config["vocab_size"], size=n).astype(np.int32)
    ref = _oracle_all_logits(om, tokens)
    assert gm.prefill(tokens[:cut], 0, want_logits=0) is None          # hydrate only
    got = gm.prefill(tokens[cut:], cut, want_logits=2)                 # second chunk attends to the first through the cache
    assert np.max(np.abs(got - ref[cut:])) <= LOGIT_TOL
    with pytest.raises(capi.XalmError):                                 # would wrap the ring: not a prefill job
        gm.prefill(tokens, config["max_seq_len"] - 10, want_logits=0)
    gm.close(); om.close()


def test_perplexity_probabilities():
    """main.cpp:244-258: logprob = log(sample_prob(next token)) at every position."""
    config, om, gm = synth_pair("tiny", "q8_0", seed=12, std=0.06)
    tokens = np.random.default_rng(4).integers(3, config["vocab_size"], size=101).astype(np.int32)
    ref_lp = []
    for pos in range(100):
        lg = om.forward(int(tokens[pos]), pos, 1)
        ref_lp.append(np.log(oracle.sample_prob(lg, int(tokens[pos + 1]))))
    _, probs = gm.prefill(tokens[:-1], 0, want_logits=2, targets=tokens[1:])
    lp = np.log(probs)
    assert np.max(np.abs(lp - np.array(ref_lp))) <= 2e-2
    ppl_ref, ppl = np.exp(-np.mean(ref_lp)), np.exp(-np.mean(lp))
    assert abs(ppl - ppl_ref) <= 2e-3 * ppl_ref
    gm.close(); om.close()


def test_prefill_gelu_tied_embeddings_partial_rotary():
    """The config switches of model.h:44-90 through the batched path: gelu GLU, classifier tied to the embedding, rotary_dim < head_dim."""
    config, om, gm = synth_pair("tiny", "f16", seed=8, std=0.06, act_type="gelu", tie_word_embeddings=True, rotary_dim=32)
    tokens = np.random.default_rng(5).integers(3, config["vocab_size"], size=90).astype(np.int32)
    ref = _oracle_all_logits(om, tokens)
    got = gm.prefill(tokens, 0, want_logits=2)
    assert np.max(np.abs(got - ref)) <= 1e-3
    gm.close(); om.close()


def test_prefill_single_token_and_full_cache():
    """Edge sizes: one position; a batch that fills the KV cache exactly (max_seq_len = 128 for the tiny shape)."""
    config, om, gm = synth_pair("tiny", "q8_0", seed=13, std=0.06)
    tokens = np.random.default_rng(6).integers(3, config["vocab_size"], size=config["max_seq_len"]).astype(np.int32)
    ref = _oracle_all_logits(om, tokens)
    one = gm.prefill(tokens[:1], 0, want_logits=1)
    assert np.max(np.abs(one - ref[0])) <= 1e-3
    got = gm.prefill(tokens, 0, want_logits=2)
    assert np.max(np.abs(got - ref)) <= 1e-3
    with pytest.raises(capi.XalmError):
        gm.prefill(tokens[:2], config["max_seq_len"] - 1, want_logits=0)
    gm.close(); om.close()


def test_prefill_4096_tokens_vs_oracle():
    """BASELINE config 3 at its full length: 4096 positions in one pass through attn_tc_kernel (head_dim 128, GQA 8:2) and the
    tcgen05 GEMMs, against the ORACLE run token by token (not against the repo's own decode kernels): logits at 40 positions
    spread over the sequence plus the last 8, max-abs 1e-2, and Sampler::sample_prob at every position."""
    config, om, gm = synth_pair("small", "q8_0", seed=3, std=0.03, n_layers=2, max_seq_len=4096)
    assert config["max_seq_len"] == 4096 and config["head_dim"] == 128
    n = 4096
    toks = np.random.default_rng(17).integers(3, config["vocab_size"], size=n + 1).astype(np.int32)
    check = sorted(set(list(range(0, n, 100)) + list(range(n - 8, n)) + [1, 127, 128, 129, 2047, 2048]))
    want = {}
    probs_o = np.zeros(n, np.float32)
    for pos in range(n):
        lg = om.forward(int(toks[pos]), pos, 1)
        probs_o[pos] = oracle.sample_prob(lg, int(toks[pos + 1]))
        if pos in check:
            want[pos] = lg.copy()
    lg_all, probs = gm.prefill(toks[:-1], 0, want_logits=2, targets=toks[1:])
    worst = max(float(np.max(np.abs(lg_all[pos] - want[pos]))) for pos in check)
    assert worst <= LOGIT_TOL, f"4096-token prefill logits differ from the oracle by {worst}"
    assert np.max(np.abs(probs - probs_o)) <= 1e-3 * float(probs_o.max()) + 1e-6
    ppl_g, ppl_o = float(np.exp(-np.mean(np.log(probs)))), float(np.exp(-np.mean(np.log(probs_o))))
    assert abs(ppl_g - ppl_o) <= 1e-3 * ppl_o, (ppl_g, ppl_o)
    # and decoding continues from the prefilled cache: ring is full at 4096, so this token wraps (sinks active)
    st = InferenceState(config)
    gm.forward(st, int(toks[n]), n, 1)
    lg_o = om.forward(int(toks[n]), n, 1)
    assert np.max(np.abs(st.logits() - lg_o)) <= LOGIT_TOL
    gm.close(); om.close()
