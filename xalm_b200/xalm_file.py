""".xalm checkpoint reader / writer.

Layout (reader: /root/reference/src/xalm.h:90-192; writer: /root/reference/convert.py:248-321):

    u64 LE  H            absolute offset of the data blob; 4096-aligned, >= 8 + len(json) + 128
    bytes   JSON         {"xalm": {"version": 1},
                          "<Arch>": {"config": {... all values are strings ...},
                                     "tensors": {name: {"type", "shape", "hash", "offset", "size"}}}}
    zeros                up to H
    blob                 tensor i at H + offset_i, offsets 32-byte aligned, order = sort_tensor_names
                         (convert.py:1057-1108); "tokenizer.tokens" (U8, NUL-separated) last.

Block-quantised tensors are stored as uint8 rows and carry their BYTE shape in the header
(`[rows, cols/32*34]` for q8_0; quants.py:79-83).  The reference reader cannot load those
(SURVEY.md §0.4); this one maps them back to element shapes through `types.XType.elem_shape`.
"""
from __future__ import annotations

import json
import mmap
import os
import re
import struct
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np

from . import types as T

SUPPORTED_ARCHITECTURES = ("MistralForCausalLM", "LlamaForCausalLM")  # xalm.h:141, convert.py:28-31

_LAYER_ORDER = ["attn.norm.weight", "mlp.norm.weight", "attn.q.weight", "attn.k.weight", "attn.v.weight",
                "attn.down.weight", "mlp.gate.weight", "mlp.down.weight", "mlp.up.weight"]


def align_offset(offset: int, alignment: int = 32) -> int:
    return (offset + alignment - 1) // alignment * alignment


def sort_tensor_names(names) -> list:
    """On-disk order (convert.py:1057-1108): embed, layers in numeric order (norms, q, k, v, down, gate,
    down, up), output.weight, output.norm.weight, tokenizer.tokens."""
    ordered, layers = [], {}
    pat = re.compile(r"l\.(\d+)\.")
    for key in names:
        if key == "embed.weight":
            ordered.insert(0, key)
        elif key in ("output.weight", "output.norm.weight", "tokenizer.tokens"):
            continue
        else:
            m = pat.search(key)
            if not m:
                raise ValueError(f"unexpected tensor name {key}")
            layers.setdefault(int(m.group(1)), []).append(key)
    for l in sorted(layers):
        def rank(k, l=l):
            tail = k.split(f"l.{l}.")[-1]
            return _LAYER_ORDER.index(tail) if tail in _LAYER_ORDER else 100
        ordered.extend(sorted(layers[l], key=rank))
    for tail in ("output.weight", "output.norm.weight", "tokenizer.tokens"):
        if tail in names:
            ordered.append(tail)
    return ordered


def write_xalm(path: str, arch: str, config: dict, tensors: "OrderedDict[str, tuple[str, np.ndarray]]") -> None:
    """Write a checkpoint byte-identical to what convert.py's save_xalm_binary would write for the
    same tensors.  `tensors`: name -> (type name as convert.py spells it, e.g. "q8_0", array whose
    raw bytes are the payload and whose shape is the header shape); dict order = header order."""
    import xxhash
    if arch not in SUPPORTED_ARCHITECTURES:
        raise ValueError(f"Architecture {arch} is not supported")
    names = sort_tensor_names(list(tensors.keys()))
    meta = OrderedDict((name, {"type": tname, "shape": list(arr.shape)}) for name, (tname, arr) in tensors.items())
    cur = 0
    for name in names:
        arr = np.ascontiguousarray(tensors[name][1])
        cur = align_offset(cur)
        raw = arr.view(np.uint8).reshape(-1)
        meta[name]["hash"] = xxhash.xxh3_64(raw.data).intdigest()
        meta[name]["offset"] = cur
        meta[name]["size"] = raw.size
        cur += raw.size
    header = {"xalm": {"version": 1}, arch: {"config": {k: str(v) for k, v in config.items()}, "tensors": meta}}
    blob = json.dumps(header).encode("utf-8")
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(blob)))
        f.write(blob)
        pad = align_offset(f.tell() + 128, 4096) - f.tell()
        f.write(b"\x00" * pad)
        start = f.tell()
        f.seek(0)
        f.write(struct.pack("<Q", start))
        f.seek(start)
        for name in names:
            pos = f.tell()
            f.write(b"\x00" * (align_offset(pos) - pos))
            f.write(np.ascontiguousarray(tensors[name][1]).view(np.uint8).reshape(-1).data)


@dataclass
class TensorInfo:
    name: str
    type: T.XType
    disk_shape: tuple      # as stored in the header (byte shape for block formats)
    shape: tuple           # element shape
    offset: int            # absolute file offset
    size: int              # bytes
    hash: int | None


class XalmFile:
    """Xalm::load (xalm.h:90-192) + file_info::load_tensor (xalm.h:26-47)."""

    def __init__(self, path: str):
        self.path = path
        file_size = os.path.getsize(path)
        self._f = open(path, "rb")
        (h,) = struct.unpack("<Q", self._f.read(8))
        if h == 0 or h > file_size - 8:
            raise ValueError(f"bad json size: {h} for file size: {file_size}")  # xalm.h:102-104
        raw = self._f.read(h - 8)
        header = json.loads(raw.split(b"\x00", 1)[0].decode("utf-8"))
        if "xalm" not in header:
            raise ValueError("invalid file format!")  # xalm.h:124-126
        ver = header["xalm"].get("version", 0)
        if ver != 1:
            raise ValueError(f"xalm version mismatch: {ver}")  # xalm.h:121-123
        self.data_offset = h
        self.arch = None
        self.metadata = None
        self.tensors: "OrderedDict[str, TensorInfo]" = OrderedDict()
        for key, val in header.items():
            if key == "xalm":
                continue
            if key not in SUPPORTED_ARCHITECTURES:
                raise ValueError(f"unsupported model architecture: {key}")  # xalm.h:186-188
            self.arch = key
            self.metadata = val["config"]
            for name, ti in val["tensors"].items():
                typ = T.parse(ti.get("type", "<missing>"))
                shape = tuple(int(s) for s in ti["shape"])
                if len(shape) > 4:
                    raise ValueError("shape exceeds 4 dimensions")  # xalm.h:149-151
                off, size = int(ti.get("offset", -1)), int(ti.get("size", -1))
                if off < 0:
                    raise ValueError("bad offset")
                if size < 0:
                    raise ValueError("bad size")
                if size == 0 or h + off + size > file_size:
                    raise ValueError("offset out of range")  # stricter than xalm.h:112,173
                eshape = typ.elem_shape(shape)
                n = 1
                for s in eshape:
                    n *= s
                if typ.nbytes(n) != size:
                    raise ValueError(f"size mismatch for {name}: {size} bytes vs shape {eshape} of {typ.name}")
                self.tensors[name] = TensorInfo(name, typ, shape, eshape, h + off, size, ti.get("hash"))
        if self.arch is None:
            raise ValueError("no model architecture in header")
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)

    def raw(self, name: str) -> np.ndarray:
        """The tensor's payload as a read-only uint8 view of the mapped file."""
        ti = self.tensors[name]
        return np.frombuffer(self._mm, dtype=np.uint8, count=ti.size, offset=ti.offset)

    def verify_hashes(self) -> None:
        import xxhash
        for name, ti in self.tensors.items():
            if ti.hash is not None and xxhash.xxh3_64(self.raw(name).data).intdigest() != int(ti.hash):
                raise ValueError(f"xxh3 mismatch for tensor {name}")

    def tokens(self) -> list:
        """NUL-separated vocab (tokenizer.cpp:33-43)."""
        raw = self.raw("tokenizer.tokens").tobytes()
        parts = raw.split(b"\x00")
        return parts[:-1] if raw.endswith(b"\x00") else parts

    def close(self):
        # numpy views of the map may still be alive; let the GC take the map in that case
        try:
            self._mm.close()
        except BufferError:
            pass
        self._f.close()


def parse_config(metadata: dict, context: int = 0) -> dict:
    """Config::from_xalm (model.h:44-90): every value is a string; max_seq_len is clamped to 4096
    unless `-T context` overrides it; defaults norm_eps 1e-5, act gelu, qkv_clip FLT_MAX."""
    c = {k: int(metadata[k]) for k in ("dim", "hidden_dim", "head_dim", "n_layers", "n_heads", "n_kv_heads", "vocab_size")}
    c["max_seq_len"] = min(int(metadata["max_seq_len"]), 4096)
    if context:
        c["max_seq_len"] = int(context)
    c["rope_theta"] = float(np.float32(float(metadata["rope_theta"])))
    c["rotary_dim"] = int(metadata["rotary_dim"])
    c["norm_eps"] = float(np.float32(float(metadata.get("norm_eps", "1e-5"))))
    act = metadata.get("act_type", "gelu")
    c["act"] = 1 if act == "silu" else 0       # unknown -> gelu (model.h:71-74)
    c["norm_type"] = 0
    c["qkv_clip"] = float(metadata["qkv_clip"]) if "qkv_clip" in metadata else float(np.finfo(np.float32).max)
    c["tie_word_embeddings"] = metadata["tie_word_embeddings"] == "True"
    c["bos_token_id"] = _first_int(metadata.get("bos_token_id", "-1"))
    c["eos_token_id"] = _first_int(metadata.get("eos_token_id", "-1"))
    return c


def _first_int(s: str) -> int:
    """tokenizer.cpp:4-21 parse_str: "[1, 2]" or "1" -> first element."""
    s = s.strip()
    if s.startswith("[") and s.endswith("]"):
        s = s[1:-1].split(",")[0]
    return int(s)


def expected_tensors(c: dict) -> "OrderedDict[str, tuple]":
    """Tensor names and element shapes Model::from_xalm asks for (model.cpp:83-114)."""
    q_dim, kv_dim = c["n_heads"] * c["head_dim"], c["n_kv_heads"] * c["head_dim"]
    out = OrderedDict()
    out["embed.weight"] = (c["vocab_size"], c["dim"])
    for i in range(c["n_layers"]):
        out[f"l.{i}.attn.norm.weight"] = (c["dim"],)
        out[f"l.{i}.mlp.norm.weight"] = (c["dim"],)
        out[f"l.{i}.attn.q.weight"] = (q_dim, c["dim"])
        out[f"l.{i}.attn.k.weight"] = (kv_dim, c["dim"])
        out[f"l.{i}.attn.v.weight"] = (kv_dim, c["dim"])
        out[f"l.{i}.attn.down.weight"] = (c["dim"], q_dim)
        out[f"l.{i}.mlp.gate.weight"] = (c["hidden_dim"], c["dim"])
        out[f"l.{i}.mlp.down.weight"] = (c["dim"], c["hidden_dim"])
        out[f"l.{i}.mlp.up.weight"] = (c["hidden_dim"], c["dim"])
    out["output.norm.weight"] = (c["dim"],)
    if not c["tie_word_embeddings"]:
        out["output.weight"] = (c["vocab_size"], c["dim"])
    return out
