"""GPU: decode throughput of the m7 shape in several weight formats (device-timed, positions spread over 4k)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xalm_b200 import capi, synth, types as T, xalm_file as X
import bench
st = torch.cuda.Stream()
print("| format | bytes/weight | ms/token | tok/s | GB/s (active_bytes) | % of measured 6538 GB/s |")
print("|---|---|---|---|---|---|")
for w in sys.argv[1:] or ["f16", "bf16", "f8_e4m3", "q8_0", "q5_1", "q4_0"]:
    wtype = T.parse(w)
    cfg_full = synth.model_config("m7")
    cfg = X.parse_config(synth.metadata_strings(cfg_full), 4096)
    model, _ = bench.build_model_streaming(cfg_full, cfg, wtype, 0, False, device=0, stream=st.cuda_stream)
    K = 64
    pos = bench.positions_for(K, 4096)
    for i in range(4): model.forward_async(5, pos[i], 1)
    model.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(K): model.forward_async(5, pos[i], 1)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    by = np.mean([model.active_bytes(p) for p in pos])
    print(f"| {w} | {wtype.bytes/wtype.block:.4g} | {ms:.3f} | {1e3/ms:.0f} | {by/ms/1e6:.0f} | {100*by/ms/1e6/6538.3:.0f} % |", flush=True)
    model.close()
