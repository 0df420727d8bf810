"""Tensor-parallel plan and plumbing (host side of SURVEY.md §8e; the reference has no parallelism at all).

The C ABI takes FULL tensors and keeps this rank's shard; `shard_ranges` states which (rows, element columns) that is so
tests can check the plan and emulate it on the CPU.  `init_model_parallel` does the one exchange the backend needs from
the host: shipping rank 0's communicator id to every rank.
"""
from __future__ import annotations

import numpy as np


def shard_ranges(name: str, config: dict, rank: int, size: int):
    """((row0, row1), (col0, col1)) of tensor `name` (element coordinates) that rank `rank` of `size` keeps.
    q/k/v and gate/up: output rows by head / hidden slice; attn.down and mlp.down: input columns; classifier: vocab rows;
    norms and the embedding table: replicated."""
    c = config
    q_dim, kv_dim = c["n_heads"] * c["head_dim"], c["n_kv_heads"] * c["head_dim"]
    if c["n_heads"] % size or c["n_kv_heads"] % size or c["hidden_dim"] % (32 * size) or c["vocab_size"] % (32 * size):
        raise ValueError(f"tp_size {size} does not divide heads/hidden/vocab into 32-aligned shards")
    full = lambda r, cc: ((0, r), (0, cc))
    part = lambda n: (rank * (n // size), (rank + 1) * (n // size))
    if name.endswith("norm.weight"):
        return ((0, c["dim"]), None)
    if name == "embed.weight":
        return full(c["vocab_size"], c["dim"])
    if name == "output.weight":
        return (part(c["vocab_size"]), (0, c["dim"]))
    tail = name.split(".", 2)[2]
    if tail == "attn.q.weight":
        return (part(q_dim), (0, c["dim"]))
    if tail in ("attn.k.weight", "attn.v.weight"):
        return (part(kv_dim), (0, c["dim"]))
    if tail == "attn.down.weight":
        return ((0, c["dim"]), part(q_dim))
    if tail in ("mlp.gate.weight", "mlp.up.weight"):
        return (part(c["hidden_dim"]), (0, c["dim"]))
    if tail == "mlp.down.weight":
        return ((0, c["dim"]), part(c["hidden_dim"]))
    raise KeyError(name)


def slice_raw(t, shape, raw: np.ndarray, rows, cols) -> np.ndarray:
    """Cut ((r0,r1),(c0,c1)) out of a tensor stored in on-disk bytes of XType t (block formats: whole blocks only)."""
    if cols is None:
        return raw
    n_rows, n_cols = shape
    row_bytes = t.nbytes(n_cols)
    m = raw.reshape(n_rows, row_bytes)
    (r0, r1), (c0, c1) = rows, cols
    if c0 % t.block or c1 % t.block:
        raise ValueError("column slice splits a quantisation block")
    return np.ascontiguousarray(m[r0:r1, t.nbytes(c0): t.nbytes(c1)]).reshape(-1)


def broadcast_comm_id(dist, rank: int, device=None) -> bytes:
    """Rank 0 creates the backend's communicator id; everyone receives it through torch.distributed."""
    import torch
    from .model import Model
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        buf = torch.frombuffer(bytearray(Model.comm_unique_id()), dtype=torch.uint8).to(buf.device)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def make_ipc_exchange(dist, world: int, device="cuda"):
    """Returns the callable Model.cuda(ipc_exchange=...) wants: all-gather one 64-byte CUDA IPC handle per rank."""
    import torch

    def exchange(mine: bytes) -> bytes:
        t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(device)
        out = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(world)]
        dist.all_gather(out, t)
        return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)

    return exchange
