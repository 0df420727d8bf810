"""Run a few tokens of a depth-reduced m7 model through the megakernel (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xalm_b200 import capi, synth, types as T, xalm_file as X
from xalm_b200.model import Model
cfg_full = synth.model_config("m7")
cfg_full["n_layers"] = int(os.environ.get("LAYERS", "2"))
cfg_full["vocab_size"] = 4096
cfg = X.parse_config(synth.metadata_strings(cfg_full), 4096)
wtype = T.parse(os.environ.get("WTYPE", "q8_0"))
model = Model.from_tensors(cfg, synth.iter_tensors(cfg_full, wtype, 0)).cuda(device=0)
for i in range(6):
    model.forward_async(5, 100 + i, 1)
model.sync()
print("ok", model.last_launch_count())
if os.environ.get("SKEW"):
    capi.timeline_start(100000)
    model.forward_async(5, int(os.environ["SKEW"]), 1)
    model.sync()
    tl = capi.timeline_stop(100000).astype(np.int64)
    tl = tl[tl[:, 0] >= 1000]
    names = ["qkv", "attn", "wo", "w13", "w2"]
    prev_max = None
    for ph in sorted(set(tl[:, 0])):
        r = tl[tl[:, 0] == ph]
        t = r[:, 1]
        order = np.argsort(t)
        line = f"phase {ph-1000:3d} arrive-before-{names[(ph-1000) % 5]:4s}: min {0 if prev_max is None else (t.min()-prev_max)/1e3:7.2f} median {0 if prev_max is None else (np.median(t)-prev_max)/1e3:7.2f} max {0 if prev_max is None else (t.max()-prev_max)/1e3:7.2f} us after prev barrier complete; spread {(t.max()-t.min())/1e3:6.2f}; last CTAs {r[order[-4:], 2].tolist()} first {r[order[:3], 2].tolist()}"
        print(line)
        prev_max = t.max()
