"""GPU experiment: the one-kernel-per-token decode path (decode_mega.cu) — parity against the oracle on small models for every
integer format, then device time per token on a full-size shape with the token kernel on and off, plus its phase timeline."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from xalm_b200 import capi, synth, types as T, xalm_file as X
from xalm_b200.model import Model
import bench

PROMPT = [1, 84, 61, 35, 90, 107, 100, 119]


def parity():
    from gpu_util import greedy_compare, synth_pair
    ok = True
    for shape, std, steps in (("tiny", 0.06, 24), ("small", 0.03, 16)):
        for wt in ("q8_0", "q4_0", "q4_1", "q5_0", "q5_1", "q8"):
            for mega in (1, 0):
                capi.tune("mega", mega)
                config, om, gm = synth_pair(shape, wt, seed=2, std=std)
                t0 = time.time()
                otoks, gtoks, maxdiff, margin = greedy_compare(config, om, gm, PROMPT, steps)
                good = otoks == gtoks and maxdiff <= 1e-2
                ok &= good
                print(f"parity {shape:5s} {wt:5s} mega={mega} launches={gm.last_launch_count()} tokens_match={otoks == gtoks} "
                      f"max_abs_logit={maxdiff:.3e} margin={margin:.2e} {'OK' if good else 'FAIL'} ({time.time()-t0:.1f}s)", flush=True)
                gm.close(); om.close()
    capi.tune("mega", 1)
    return ok


def perf(shape, wtname, layers=None):
    wtype = T.parse(wtname)
    cfg_full = synth.model_config(shape)
    if layers:
        cfg_full["n_layers"] = layers
    cfg = X.parse_config(synth.metadata_strings(cfg_full), 4096)
    st = torch.cuda.Stream()
    t0 = time.time()
    tensors = list(synth.iter_tensors(cfg_full, wtype, 0))
    print(f"{shape} {wtname}: generated in {time.time()-t0:.0f}s", flush=True)
    K = 64
    pos = bench.positions_for(K, 4096)
    ref_tok = None
    cfgs = [dict(kv.split("=") for kv in c.split(",")) for c in os.environ.get("CONFIGS", "mega=1;mega=0").split(";")]
    for cfgd in cfgs:
        for k, v in cfgd.items():
            capi.tune(k, int(v))
        mega = int(cfgd.get("mega", 0))
        capi.tune("mega", mega)
        capi.tune("mega_timeline", 1 if mega else 0)
        model = Model.from_tensors(cfg, tensors).cuda(device=0, stream=st.cuda_stream)
        for i in range(4):
            model.forward_async(5, pos[i], 1)
        model.sync()
        res = []
        for plist, label in ((pos, "spread"), (list(range(K)), "pos0-63")):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(st)
            for i in range(K):
                model.forward_async(5, plist[i], 1)
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            byts = np.mean([model.active_bytes(p) for p in plist])
            res.append(f"{label}: {ms/K:.3f} ms/tok {K/(ms/1e3):.0f} tok/s {byts*K/(ms/1e3)/1e9:.0f} GB/s")
        model.sync()
        lg = model.read_state(capi.S_LOGITS, cfg["vocab_size"])
        print(f"{shape} {wtname} {cfgd}: " + " | ".join(res) + f" launches/token {model.last_launch_count()} argmax {int(np.argmax(lg))}", flush=True)
        if ref_tok is None:
            ref_lg = lg
        else:
            print(f"   max |logit(mega) - logit(multi-kernel)| = {np.max(np.abs(lg - ref_lg)):.3e}", flush=True)
        ref_tok = 1
        if mega:
            model.forward_async(5, 4000, 1)
            model.sync()
            t4, arr = capi.mega_timeline(model._h)
            if t4 is not None:
                t4 = t4.astype(np.int64); arr = arr.astype(np.int64)
                nph = t4.shape[0]
                names = ["qkv", "attn", "wo", "w13", "w2"]
                agg = {}
                for ph in range(nph):
                    nm = names[ph % 5] if ph < nph - 1 else "cls"
                    a, b, c, d = t4[ph][:4]
                    skew = (arr[ph].max() - arr[ph].min()) / 1e3
                    ex = [(int(t4[ph][k]) - int(a)) / 1e3 if t4[ph][k] else -1.0 for k in (4, 5, 6, 7)]
                    agg.setdefault(nm, []).append(((b - a) / 1e3, (c - b) / 1e3, (d - c) / 1e3, skew, *ex))
                tot = (t4[-1, 3] - t4[0, 0]) / 1e3
                print(f"   token kernel at pos 4000: {tot:.1f} us total (CTA 0); per phase kind: mean us of [hand-off wait, staging, tiles] and arrival skew over CTAs")
                for nm, v in agg.items():
                    v = np.array(v)
                    print(f"     {nm:5s} n={len(v):3d} handoff {v[:,0].mean():6.2f}  stage {v[:,1].mean():6.2f}  tiles {v[:,2].mean():6.2f}  skew {v[:,3].mean():6.2f}"
                          f"   marks since phase entry: {v[:,4].mean():6.2f} {v[:,5].mean():6.2f} {v[:,6].mean():6.2f} {v[:,7].mean():6.2f}")
        model.close()
    capi.tune("mega", 1); capi.tune("mega_timeline", 0)


if __name__ == "__main__":
    torch.cuda.set_device(0)
    what = sys.argv[1:] or ["parity", "m7:q8_0"]
    for w in what:
        if w == "parity":
            print("PARITY", "OK" if parity() else "FAILED", flush=True)
        else:
            shape, wt = w.split(":")[:2]
            layers = int(w.split(":")[2]) if w.count(":") > 1 else None
            perf(shape, wt, layers)
