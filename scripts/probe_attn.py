"""GPU probe: whole-token time vs attention knobs at long context (m7 q8_0, fewer layers to keep it quick)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xalm_b200 import capi, synth, types as T, xalm_file as X
from xalm_b200.model import Model
cfg_full = synth.model_config("m7"); cfg_full["n_layers"] = 8; cfg_full["vocab_size"] = 4096
cfg = X.parse_config(synth.metadata_strings(cfg_full), 4096)
tensors = list(synth.iter_tensors(cfg_full, T.Q8_0, 0))
st = torch.cuda.Stream()
for kn in [eval(a) for a in sys.argv[1:]]:
    for k, v in kn.items(): capi.tune(k, v)
    model = Model.from_tensors(cfg, tensors).cuda(device=0, stream=st.cuda_stream)
    out = []
    for P in (0, 511, 2047, 4095):
        for i in range(3): model.forward_async(5, P, 1)
        model.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(20): model.forward_async(5, P, 1)
        e1.record(st); torch.cuda.synchronize()
        out.append(f"pos{P}: {e0.elapsed_time(e1)/20/8*1e3:.1f}us/layer")
    print(kn, " ".join(out), flush=True)
    model.close()
