"""Type registry shared by the host side, the tests and bench.

Ids 0..9 are the reference's `Type` ids (/root/reference/src/types.h:505-514); the block
formats take the values convert.py gives them in `XType` (/root/reference/convert.py:56-61),
since the reference C++ never assigned them an id.  `block`/`bytes` are
quants.py's GGML_QUANT_SIZES (/root/reference/quants.py:45-77).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class XType:
    id: int
    name: str          # upper-case, as Type::name() prints it (types.h:358-378)
    block: int         # elements per block
    bytes: int         # bytes per block

    @property
    def bytes_per_weight(self) -> float:
        return self.bytes / self.block

    def nbytes(self, n_elems: int) -> int:
        if n_elems % self.block:
            raise ValueError(f"{n_elems} elements is not a multiple of the {self.name} block ({self.block})")
        return n_elems // self.block * self.bytes

    def byte_shape(self, shape):
        """Element shape -> on-disk shape (quants.py:79-83: block formats store uint8 rows)."""
        if self.block == 1:
            return tuple(shape)
        return (*shape[:-1], shape[-1] // self.block * self.bytes)

    def elem_shape(self, shape):
        """On-disk shape -> element shape (quants.py:86-90)."""
        if self.block == 1:
            return tuple(shape)
        if shape[-1] % self.bytes:
            raise ValueError(f"bytes per row ({shape[-1]}) is not a multiple of the {self.name} type size ({self.bytes})")
        return (*shape[:-1], shape[-1] // self.bytes * self.block)


F32 = XType(1, "F32", 1, 4)
F16 = XType(2, "F16", 1, 2)
BF16 = XType(3, "BF16", 1, 2)
F8_E2M5 = XType(4, "F8_E2M5", 1, 1)
F8_E3M4 = XType(5, "F8_E3M4", 1, 1)
F8_E4M3 = XType(6, "F8_E4M3", 1, 1)
F8_E5M2 = XType(7, "F8_E5M2", 1, 1)
U8 = XType(8, "U8", 1, 1)
Q8 = XType(9, "Q8", 1, 1)
Q4_0 = XType(1007, "Q4_0", 32, 18)
Q4_1 = XType(1008, "Q4_1", 32, 20)
Q5_0 = XType(1009, "Q5_0", 32, 22)
Q5_1 = XType(1010, "Q5_1", 32, 24)
Q8_0 = XType(1011, "Q8_0", 32, 34)
TQ1_0 = XType(1012, "TQ1_0", 256, 54)
QI8 = XType(2007, "QI8", 1, 1)

ALL = [F32, F16, BF16, F8_E2M5, F8_E3M4, F8_E4M3, F8_E5M2, U8, Q8, Q4_0, Q4_1, Q5_0, Q5_1, Q8_0, TQ1_0, QI8]
BY_NAME = {t.name: t for t in ALL}
BY_ID = {t.id: t for t in ALL}
# every type a weight matrix may have (U8 is tokenizer bytes only)
MATMUL_TYPES = [t for t in ALL if t is not U8]
BLOCK_TYPES = [t for t in ALL if t.block > 1]


def parse(s: str) -> XType:
    """Type::parse (types.h:468-499): case-insensitive; extended with the block-format names."""
    t = BY_NAME.get(s.upper())
    if t is None:
        raise ValueError(f"invalid type: {s}")
    return t
