"""GPU probe: where does a decode step's time go?  CPU launch cost vs device time, graph/PDL on/off, and the
stand-alone matvec kernel at the model's shapes.  Usage: python scripts/probe_launch.py [shape] [wtype]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from xalm_b200 import capi, synth, types as T, xalm_file as X
import bench

shape = sys.argv[1] if len(sys.argv) > 1 else "m7"
wtype = T.parse(sys.argv[2] if len(sys.argv) > 2 else "q8_0")
cfg_full = synth.model_config(shape)
cfg = X.parse_config(synth.metadata_strings(cfg_full), 4096)
dim, hid = cfg["dim"], cfg["hidden_dim"]
qkv = (cfg["n_heads"] + 2 * cfg["n_kv_heads"]) * cfg["head_dim"]
bpw = wtype.bytes / wtype.block
print("== stand-alone matvec kernels (ms, GB/s), pdl on/off")
for pdl in (1, 0):
    capi.tune("pdl", pdl)
    for name, n, d, epi, norm in (("qkv", dim, qkv, 0, True), ("wo", dim, dim, 0, False), ("w13", dim, hid, 2, True), ("w2", hid, dim, 0, False),
                                  ("cls", dim, cfg["vocab_size"], 0, True)):
        rows = 2 * d if epi == 2 else d
        by = rows * n * bpw
        nbuf = max(2, int(np.ceil(600e6 / by)))
        ms = capi.bench_matvec(wtype.id, n, d, nbuf, 300, epi=epi, with_norm=norm)
        print(f"pdl={pdl} {name:4s} {rows}x{n}: {ms*1e3:8.2f} us  {by/ms/1e6:8.1f} GB/s")
capi.tune("pdl", 1)
if os.environ.get("PROBE_KERNELS_ONLY"):
    sys.exit(0)
torch.cuda.set_device(0)
st = torch.cuda.Stream()
for graph, pdl in ((1, 1), (1, 0), (0, 1), (0, 0)):
    capi.tune("graph", graph); capi.tune("pdl", pdl)
    model, _ = bench.build_model_streaming(cfg_full, cfg, wtype, 0, False, device=0, stream=st.cuda_stream)
    K = 64
    pos = bench.positions_for(K, 4096)
    for i in range(4):
        model.forward_async(5, pos[i], 1)
    model.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    t0 = time.perf_counter()
    for i in range(K):
        model.forward_async(5, pos[i], 1)
    t_cpu = time.perf_counter() - t0
    e1.record(st)
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    byts = np.mean([model.active_bytes(p) for p in pos])
    print(f"graph={graph} pdl={pdl}: cpu enqueue {t_cpu/K*1e3:.3f} ms/tok, device {ms/K:.3f} ms/tok ({K/(ms/1e3):.0f} tok/s, {byts*K/(ms/1e3)/1e9:.0f} GB/s), wall {t_all/K*1e3:.3f} ms/tok")
    # short-context variant: pos 0..K
    e0.record(st)
    for i in range(K):
        model.forward_async(5, i, 1)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"        pos 0..{K}: device {ms/K:.3f} ms/tok ({K/(ms/1e3):.0f} tok/s, {model.active_bytes(K//2)*K/(ms/1e3)/1e9:.0f} GB/s)")
    model.close()
