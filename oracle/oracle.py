"""ctypes binding for liboracle.so — TEST INFRASTRUCTURE ONLY (see xalm_oracle.cpp header).

May be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs, nowhere else.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "xalm_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


class OrcConfig(C.Structure):
    # model.h:25-42
    _fields_ = [("dim", C.c_int), ("hidden_dim", C.c_int), ("head_dim", C.c_int), ("n_layers", C.c_int),
                ("n_heads", C.c_int), ("n_kv_heads", C.c_int), ("vocab_size", C.c_int), ("max_seq_len", C.c_int),
                ("rope_theta", C.c_float), ("rotary_dim", C.c_int), ("norm_eps", C.c_float), ("act", C.c_int),
                ("norm_type", C.c_int), ("qkv_clip", C.c_float), ("tie_word_embeddings", C.c_int)]


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    fp = C.POINTER(C.c_float)
    vp = C.c_void_p
    L.orc_type_nbytes.restype = C.c_longlong
    L.orc_type_nbytes.argtypes = [C.c_int, C.c_longlong]
    L.orc_dequant.argtypes = [C.c_int, vp, C.c_longlong, vp]
    L.orc_matmul.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_rmsnorm.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_float]
    L.orc_softmax.argtypes = [vp, vp, C.c_int]
    L.orc_rope.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
    L.orc_attn.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int]
    L.orc_mha.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    for f in ("orc_gelu", "orc_silu"):
        getattr(L, f).restype = C.c_float
        getattr(L, f).argtypes = [C.c_float]
    L.orc_clip.restype = C.c_float
    L.orc_clip.argtypes = [C.c_float, C.c_float]
    L.orc_model_create.restype = vp
    L.orc_model_create.argtypes = [C.POINTER(OrcConfig)]
    L.orc_model_destroy.argtypes = [vp]
    L.orc_model_set_acc_mode.argtypes = [vp, C.c_int]
    L.orc_model_set_tensor.argtypes = [vp, C.c_char_p, C.c_int, vp]
    L.orc_forward.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.orc_logits.restype = fp
    L.orc_logits.argtypes = [vp]
    L.orc_state.restype = fp
    L.orc_state.argtypes = [vp, C.c_int]
    L.orc_kv.restype = C.POINTER(C.c_uint16)
    L.orc_kv.argtypes = [vp, C.c_int, C.c_int]
    L.orc_sample_prob.restype = C.c_float
    L.orc_sample_prob.argtypes = [vp, C.c_int, C.c_int]
    L.orc_sample_argmax.argtypes = [vp, C.c_int]
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_set_threads.restype = None
    L.orc_num_threads.restype = C.c_int
    L.orc_active_bytes.restype = C.c_longlong
    L.orc_active_bytes.argtypes = [vp, C.c_longlong]
    _LIB = L
    return L


def set_threads(n: int) -> int:
    lib().orc_set_threads(int(n))
    return int(lib().orc_num_threads())


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def dequant(type_id: int, raw: np.ndarray, n_elems: int) -> np.ndarray:
    raw = np.ascontiguousarray(raw)
    out = np.empty(n_elems, dtype=np.float32)
    rc = lib().orc_dequant(type_id, _p(raw), n_elems, _p(out))
    if rc:
        raise ValueError(f"orc_dequant rc={rc}")
    return out


def matmul(x: np.ndarray, w_raw: np.ndarray, type_id: int, n: int, d: int, acc_mode: int = 1) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    w_raw = np.ascontiguousarray(w_raw)
    out = np.empty(d, dtype=np.float32)
    rc = lib().orc_matmul(_p(out), _p(x), _p(w_raw), type_id, n, d, acc_mode)
    if rc:
        raise ValueError(f"orc_matmul rc={rc}")
    return out


def rmsnorm(x, w_raw, wtype, eps):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    rc = lib().orc_rmsnorm(_p(out), _p(x), _p(np.ascontiguousarray(w_raw)), wtype, x.size, eps)
    if rc:
        raise ValueError(f"orc_rmsnorm rc={rc}")
    return out


def rope(vec, head_dim, pos, theta, rotary_dim):
    v = np.array(vec, dtype=np.float32, copy=True)
    lib().orc_rope(_p(v), v.size, head_dim, pos, theta, rotary_dim)
    return v


def mha(q, kb, vb, head_dim, kv_len, max_seq_len, n_heads, n_kv_heads):
    """mha_cpu (infer.cpp:498-517).  kb/vb: uint16 views of the fp16 caches (max_seq_len, n_kv_heads*head_dim)."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    out = np.zeros(n_heads * head_dim, dtype=np.float32)
    att = np.zeros(n_heads * max_seq_len, dtype=np.float32)
    lib().orc_mha(_p(out), _p(att), _p(np.ascontiguousarray(kb)), _p(np.ascontiguousarray(vb)), _p(q), head_dim, kv_len,
                  max_seq_len, n_heads, n_kv_heads)
    return out, att


class OracleModel:
    """Model + InferenceState of the reference (model.h:96-284) on the CPU oracle."""

    def __init__(self, cfg: dict, tensors: dict, acc_mode: int = 1):
        """cfg: dict with the Config fields; tensors: name -> (type_id, np.uint8 raw array)."""
        self.cfg = cfg
        c = OrcConfig(cfg["dim"], cfg["hidden_dim"], cfg["head_dim"], cfg["n_layers"], cfg["n_heads"], cfg["n_kv_heads"],
                      cfg["vocab_size"], cfg["max_seq_len"], cfg["rope_theta"], cfg["rotary_dim"], cfg["norm_eps"],
                      cfg["act"], 0, cfg.get("qkv_clip", 3.4028234663852886e38), int(cfg["tie_word_embeddings"]))
        self._L = lib()
        self._h = self._L.orc_model_create(C.byref(c))
        self._keep = []
        self._L.orc_model_set_acc_mode(self._h, acc_mode)
        for name, (tid, raw) in tensors.items():
            if name == "tokenizer.tokens":
                continue
            raw = np.ascontiguousarray(raw)
            self._keep.append(raw)
            rc = self._L.orc_model_set_tensor(self._h, name.encode(), tid, _p(raw))
            if rc:
                raise ValueError(f"oracle: unknown tensor {name}")

    def forward(self, token: int, pos: int, mode: int = 1):
        rc = self._L.orc_forward(self._h, token, pos, mode)
        if rc:
            raise RuntimeError(f"orc_forward rc={rc}")
        if mode == 0:
            return None
        return np.ctypeslib.as_array(self._L.orc_logits(self._h), shape=(self.cfg["vocab_size"],)).copy()

    def state(self, which: int, n: int) -> np.ndarray:
        return np.ctypeslib.as_array(self._L.orc_state(self._h, which), shape=(n,)).copy()

    def kv(self, layer: int, which: int) -> np.ndarray:
        n = self.cfg["max_seq_len"] * self.cfg["n_kv_heads"] * self.cfg["head_dim"]
        return np.ctypeslib.as_array(self._L.orc_kv(self._h, layer, which), shape=(n,)).copy()

    def active_bytes(self, pos: int) -> int:
        return int(self._L.orc_active_bytes(self._h, pos))

    def close(self):
        if self._h:
            self._L.orc_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sample_argmax(logits: np.ndarray) -> int:
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    return int(lib().orc_sample_argmax(_p(logits), logits.size))


def sample_prob(logits: np.ndarray, index: int) -> float:
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    return float(lib().orc_sample_prob(_p(logits), logits.size, index))
