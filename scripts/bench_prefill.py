"""Prefill / perplexity path timing on one B200: the tcgen05 GEMM alone on the model's shapes, then (with --model) the
whole 4k-token prefill of the m7 model against the token-at-a-time decode path.

    python scripts/bench_prefill.py [--T 4096] [--split 1] [--model m7 --wtype q8_0]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from xalm_b200 import capi, synth, types as T, xalm_file as X


def peak_tflops():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        return d["bf16_tflops"], d["bf16_tflops_sustained"]
    except Exception:
        return 1590.0, 1400.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=4096)
    ap.add_argument("--split", type=int, default=1)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--model", default="")
    ap.add_argument("--wtype", default="q8_0")
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--check", type=int, default=0, help="compare the last N positions' logits with the decode path")
    args = ap.parse_args()
    burst, sustained = peak_tflops()
    Tn = args.T
    shapes = (("qkv", 6144, 4096), ("wo", 4096, 4096), ("w13", 28672, 4096), ("w2", 4096, 14336), ("cls", 32000, 4096))
    tot_ms = 0.0
    for name, N, K in shapes:
        ms = capi.bench_gemm(Tn, N, K, args.split, args.iters)
        tf = 2.0 * Tn * N * K / (ms * 1e-3) / 1e12
        tot_ms += ms * (1 if name == "cls" else 32)
        print(f"gemm {name:4s} T={Tn} N={N} K={K} split={args.split}: {ms:8.3f} ms  {tf:7.1f} TFLOP/s  ({tf / burst:.2%} of measured burst {burst:.0f})", flush=True)
    print(f"sum over a 32-layer m7 forward (GEMM kernels only): {tot_ms:.1f} ms", flush=True)
    if not args.model:
        return
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench as B
    capi.tune("prefill_split", args.split)
    over = {"n_layers": args.layers} if args.layers else {}
    cfg_full = synth.model_config(args.model, **over)
    cfg = X.parse_config(synth.metadata_strings(cfg_full), 4096)
    t0 = time.time()
    model, _ = B.build_model_streaming(cfg_full, cfg, T.parse(args.wtype), 0, False, device=0)
    print(f"model built in {time.time() - t0:.0f}s", flush=True)
    rng = np.random.default_rng(5)
    toks = rng.integers(3, cfg["vocab_size"], size=Tn).astype(np.int32)
    for want in (2, 0):
        model.prefill_async(toks, 0, want); model.sync()
        t0 = time.perf_counter()
        model.prefill_async(toks, 0, want); model.sync()
        dt = time.perf_counter() - t0
        flops = 2.0 * Tn * (cfg["n_layers"] * (2 * 4096 * 4096 + 2 * 1024 * 4096 + 3 * 4096 * 14336) + (32000 * 4096 if want else 0))
        print(f"prefill T={Tn} want_logits={want}: {dt * 1e3:.1f} ms  {Tn / dt:.0f} tok/s  {flops / dt / 1e12:.1f} TFLOP/s (GEMM flops only)", flush=True)
    if args.check:
        from xalm_b200.model import InferenceState
        n = min(args.check, Tn)
        lg_all = model.prefill(toks, 0, want_logits=2)
        # the decode path over the same KV prefix: re-run the last n positions token by token (overwrites the same cache rows)
        st = InferenceState(cfg).cuda()
        worst = 0.0
        for pos in range(Tn - n, Tn):
            model.forward(st, int(toks[pos]), pos, 1)
            worst = max(worst, float(np.max(np.abs(st.logits() - lg_all[pos]))))
        print(f"max |prefill logits - decode logits| over the last {n} positions: {worst:.3e} (logit std {lg_all[-1].std():.3f})", flush=True)
        t0 = time.perf_counter()
        for pos in range(Tn - n, Tn):
            model.forward(st, int(toks[pos]), pos, 1)
        dt = (time.perf_counter() - t0) / n
        print(f"decode path: {dt * 1e3:.2f} ms/token -> {Tn * dt:.2f} s for {Tn} tokens token-at-a-time", flush=True)
    model.close()


if __name__ == "__main__":
    main()
