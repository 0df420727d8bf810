"""Run only the stand-alone matvec kernel (for ncu): python scripts/kernel_only.py <wtype> <n> <d> <epi> <norm> [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xalm_b200 import capi, types as T
t = T.parse(sys.argv[1]); n = int(sys.argv[2]); d = int(sys.argv[3]); epi = int(sys.argv[4]); norm = bool(int(sys.argv[5]))
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 20
rows = 2 * d if epi == 2 else d
by = rows * n * t.bytes // t.block
nbuf = max(2, int(np.ceil(600e6 / by)))
ms = capi.bench_matvec(t.id, n, d, nbuf, iters, epi=epi, with_norm=norm)
print(f"{t.name} {rows}x{n} epi={epi} norm={norm}: {ms*1e3:.2f} us {by/ms/1e6:.1f} GB/s")
