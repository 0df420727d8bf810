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
# "tp" / "tp8": every per-rank dimension % 256 == 0 -> TMA / integer-dot matvec with the exchange fused into its epilogue and the
# next kernel's prologue.  The worker hydrates a 40-token prompt (HYDRATE_KV_CACHE: tp_drain_kernel receives each token's last
# exchange) and then decodes 16 greedy tokens against the CPU oracle on rank 0.
CASES = [(2, "q8_0", "small"), (2, "f16", "small"), (2, "q8_0", "tp"), (2, "f16", "tp"), (2, "q4_0", "tp"),
         (4, "q8_0", "tp"), (4, "f16", "tp"), (8, "q8_0", "tp8"), (8, "f16", "tp8"), (8, "q4_0", "tp8")]


@pytest.mark.parametrize("world,wtype,shape", CASES, ids=[f"tp{w}-{t}-{s}" for w, t, s in CASES])
def test_tp_matches_oracle(world, wtype, shape):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "tests", "tp_gpu_worker.py"), wtype, shape],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "tokens match" in r.stdout


@pytest.mark.parametrize("world,wtype,shape", [(2, "q8_0", "tp"), (2, "q4_0", "small"), (4, "q8_0", "tp")])
def test_tp_shard_aware_upload(world, wtype, shape):
    """Ranks > 0 generate, quantise and upload only their shard (xalm_cuda_shard_range / xalm_cuda_upload_tensor_shard)."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, XALM_TP_SHARD="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", "29613", os.path.join(ROOT, "tests", "tp_gpu_worker.py"), wtype, shape],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "tokens match" in r.stdout


def test_tp2_long_hydrate_prompt():
    """120 HYDRATE tokens before the first logits: 120 drained exchanges in a row (the protocol hole of round 1 was here)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, XALM_TP_PROMPT="120")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29612", os.path.join(ROOT, "tests", "tp_gpu_worker.py"), "q8_0", "tp"],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "tokens match" in r.stdout
