"""Shared helpers for the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import os

import numpy as np

from oracle import oracle
from xalm_b200 import synth
from xalm_b200 import types as T
from xalm_b200 import xalm_file as X
from xalm_b200.model import InferenceState, Model, Sampler

LOGIT_TOL = 1e-2      # north star: logits within max-abs 1e-2 of the reference CPU path


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def random_raw(t: T.XType, rows: int, cols: int, seed: int, std: float = 0.05) -> np.ndarray:
    """Random weights in the on-disk layout of t (rows, cols elements), as a flat uint8 array."""
    x = synth.normal(seed, t.id, rows * cols, std).reshape(rows, cols)
    if t is T.QI8:
        x = x * 4
    if t in (T.F8_E2M5, T.F8_E3M4):
        rng = np.random.default_rng(seed)
        return rng.integers(0, 256, size=rows * cols, dtype=np.uint8)
    return np.ascontiguousarray(synth.quantize(t, x)).view(np.uint8).reshape(-1)


def load_pair(path=None, cfg_tensors=None, context=0, acc_mode=1, **cuda_kw):
    """(config, oracle model, cuda model) for a checkpoint file or a (config, tensor-iterator) pair."""
    if path is not None:
        f = X.XalmFile(path)
        model = Model.from_xalm(f, context)
        config = model.config
    else:
        config, it = cfg_tensors
        model = Model.from_tensors(config, it)
    otensors = {n: (t.id, raw) for n, (t, shape, raw) in model.tensors.items()}
    om = oracle.OracleModel(config, otensors, acc_mode)
    model.cuda(**cuda_kw)
    return config, om, model


def synth_pair(shape: str, wtype: str, seed=0, context=0, std=0.05, **over):
    c = synth.model_config(shape, **over)
    md = synth.metadata_strings(c)
    cfg = X.parse_config(md, context)
    t = T.parse(wtype)
    tensors = list(synth.iter_tensors(c, t, seed, std=std))
    return load_pair(cfg_tensors=(cfg, tensors))


def greedy_compare(config, om, gm, prompt, n_steps, check_logits=True):
    """Run the completion loop of main.cpp:94-115 on both paths, feeding each its own greedy tokens.
    Returns (oracle tokens, cuda tokens, max |logit diff| seen while the sequences agreed, min top-2 margin)."""
    state = InferenceState(config)
    sampler = Sampler(config)
    otoks, gtoks = list(prompt), list(prompt)
    maxdiff, margin = 0.0, np.inf
    lg_o = None
    for pos, tok in enumerate(prompt):
        last = pos + 1 == len(prompt)
        mode = 1 if last else 0
        lg_o = om.forward(tok, pos, mode)
        gm.forward(state, tok, pos, mode)
    for i in range(n_steps):
        lg_g = state.logits().copy()
        if otoks == gtoks:
            maxdiff = max(maxdiff, float(np.max(np.abs(lg_g - lg_o))))
            srt = np.sort(lg_o)
            margin = min(margin, float(srt[-1] - srt[-2]))
        to, tg = oracle.sample_argmax(lg_o), sampler.sample_argmax(state)
        otoks.append(to)
        gtoks.append(tg)
        lg_o = om.forward(to, len(otoks) - 1, 1)
        gm.forward(state, tg, len(gtoks) - 1, 1)
    return otoks, gtoks, maxdiff, margin
