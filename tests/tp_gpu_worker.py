"""torchrun worker: tensor-parallel CUDA backend on WORLD_SIZE GPUs vs the CPU oracle (rank 0 checks).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/tp_gpu_worker.py [wtype] [shape]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from xalm_b200 import synth, tp
from xalm_b200 import types as T
from xalm_b200 import xalm_file as X
from xalm_b200.model import InferenceState, Model, Sampler


def main():
    wtype = sys.argv[1] if len(sys.argv) > 1 else "q8_0"
    shape = sys.argv[2] if len(sys.argv) > 2 else "small"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    c = synth.model_config(shape)
    cfg = X.parse_config(synth.metadata_strings(c))
    comm_id = tp.broadcast_comm_id(dist, rank, device="cuda")
    use_peer = os.environ.get("XALM_TP_PEER", "1") != "0"
    ipc = tp.make_ipc_exchange(dist, world) if use_peer else None
    if os.environ.get("XALM_TP_SHARD") == "1":
        # shard-aware upload: ranks > 0 generate and upload only what they keep (xalm_cuda_upload_tensor_shard); rank 0 builds the
        # whole checkpoint because it also feeds the CPU oracle
        import bench
        gm, host = bench.build_model_streaming(c, cfg, T.parse(wtype), 1, rank == 0, std=0.03, device=local, tp_rank=rank, tp_size=world,
                                               comm_id=comm_id, ipc_exchange=ipc)
        tensors = None
    else:
        tensors = list(synth.iter_tensors(c, T.parse(wtype), seed=1, std=0.03))
        gm = Model.from_tensors(cfg, tensors).cuda(device=local, tp_rank=rank, tp_size=world, comm_id=comm_id, ipc_exchange=ipc)
    state, sampler = InferenceState(cfg).cuda(), Sampler(cfg)
    prompt = [int(t) for t in np.random.default_rng(0).integers(3, cfg["vocab_size"], size=int(os.environ.get("XALM_TP_PROMPT", "40")))]
    om = None
    if rank == 0:
        from oracle import oracle
        om = oracle.OracleModel(cfg, host if tensors is None else {n: (t.id, np.ascontiguousarray(a).view(np.uint8).reshape(-1)) for n, t, a in tensors})
    lg_o = None
    for pos, tok in enumerate(prompt):
        mode = 1 if pos + 1 == len(prompt) else 0
        gm.forward(state, tok, pos, mode)
        if om:
            lg_o = om.forward(tok, pos, mode)
    seq_g, seq_o, maxdiff = list(prompt), list(prompt), 0.0
    for _ in range(16):
        tg = sampler.sample_argmax(state)      # full logits are all-gathered on every rank: every rank samples the same token
        if om:
            from oracle import oracle
            maxdiff = max(maxdiff, float(np.max(np.abs(state.logits() - lg_o))))
            to = oracle.sample_argmax(lg_o)
            seq_o.append(to)
            lg_o = om.forward(to, len(seq_o) - 1, 1)
        seq_g.append(tg)
        gm.forward(state, tg, len(seq_g) - 1, 1)
    ok = torch.tensor([1], device="cuda")
    if rank == 0:
        good = seq_g == seq_o and maxdiff <= 1e-2
        print(f"TP{world} {wtype} {shape}: tokens {'match' if seq_g == seq_o else 'DIVERGE'} maxdiff {maxdiff:.2e} launches/token {gm.last_launch_count()}")
        ok[0] = 1 if good else 0
    dist.broadcast(ok, 0)
    gm.close()
    dist.destroy_process_group()
    sys.exit(0 if int(ok[0]) == 1 else 1)


if __name__ == "__main__":
    main()
