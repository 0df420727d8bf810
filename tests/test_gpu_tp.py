"""Tensor-parallel parity on real GPUs (needs >= 2 visible devices; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

from xalm_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        return capi.device_count()
    except capi.XalmError:
        return 0


# "small": per-rank hidden 1408 is not a multiple of 256 -> LDG matvec + the stand-alone allreduce kernel;
# "tp": every per-rank dimension % 256 == 0 -> TMA matvec with the exchange fused into its epilogue / the next prologue
@pytest.mark.parametrize("wtype,shape", [("q8_0", "small"), ("f16", "small"), ("q8_0", "tp"), ("f16", "tp")])
def test_tp2_matches_oracle(wtype, shape):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "tests", "tp_gpu_worker.py"), wtype, shape],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "tokens match" in r.stdout
