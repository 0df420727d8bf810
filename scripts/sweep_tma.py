"""GPU sweep of the TMA matvec knobs on the model's matrix shapes: python scripts/sweep_tma.py <wtype> "dict(...)" ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xalm_b200 import capi, types as T
wtype = T.parse(sys.argv[1] if len(sys.argv) > 1 else "q8_0")
bpw = wtype.bytes / wtype.block
shapes = (("qkv", 4096, 6144, 0, True), ("wo", 4096, 4096, 0, False), ("w13", 4096, 14336, 2, True), ("w2", 14336, 4096, 0, False), ("cls", 4096, 32000, 0, True))
for kn in [eval(a) for a in sys.argv[2:]] or [dict()]:
    for k, v in kn.items():
        capi.tune(k, v)
    out = []
    for name, n, d, epi, norm in shapes:
        rows = 2 * d if epi == 2 else d
        by = rows * n * bpw
        nbuf = max(2, int(np.ceil(600e6 / by)))
        ms = capi.bench_matvec(wtype.id, n, d, nbuf, 200, epi=epi, with_norm=norm)
        out.append(f"{name} {ms*1e3:6.2f}us {by/ms/1e6:6.0f}GB/s")
    print(f"{wtype.name} {kn}: " + " | ".join(out), flush=True)
