"""Codec registry mirroring the reference's `_Codecs[]` table (/root/reference/src/main.cpp:172-236).

The reference registers one row per (codec family, probability bits) holding function pointers of type
`decodeFunc` (main.cpp:149). Here each row carries the name the reference prints for that codec
(main.cpp:174-228), the C-ABI triple (family, state count, bits) and a `decode` callable with the reference's
argument meaning: (compressed bytes, out_capacity) -> (decoded_length or 0, output buffer).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

from . import capi


@dataclass(frozen=True)
class Codec:
    name: str           # the reference's display name, e.g. "rANS32x64 16w 12 (raw)"
    symbol: str         # the reference entry point this row replaces, e.g. "mt_rANS32x64_16w_decode_15"
    family: int
    state_count: int
    bits: int

    def decode(self, data, out_capacity: int, out=None):
        return capi.decode(self.family, self.state_count, self.bits, data, out_capacity, out)

    def capacity(self, n: int) -> int:
        return capi.capacity(self.family, self.state_count, n)


def _rows() -> List[Codec]:
    rows = []
    for n_states in (64, 32):
        for bits in (15, 14, 13, 12, 11, 10):
            rows.append(Codec(f"rANS32x{n_states} 16w {bits} (raw)", f"rANS32x{n_states}_16w_decode_scalar_{bits}",
                              capi.HSR_RAW, n_states, bits))
            rows.append(Codec(f"rANS32x{n_states} 16w {bits}", f"block_rANS32x{n_states}_16w_decode_{bits}",
                              capi.HSR_BLOCK, n_states, bits))
            rows.append(Codec(f"rANS32x{n_states} 16w {bits} mt", f"mt_rANS32x{n_states}_16w_decode_{bits}",
                              capi.HSR_MT, n_states, bits))
    for bits in (15, 14, 13, 12, 11, 10):  # main.cpp:216-228
        rows.append(Codec(f"rANS32x16 16w {bits} (raw)", f"rANS32x16_16w_decode_scalar_{bits}", capi.HSR_RAW, 16, bits))
    for bits in (15, 14, 13, 12, 11, 10):
        rows.append(Codec(f"rANS32x32 32blk 16w {bits} (raw)", f"rANS32x32_32blk_16w_decode_scalar_{bits}",
                          capi.HSR_RAW32BLK, 32, bits))
    return rows


CODECS: List[Codec] = _rows()


def find_codec(family: int, state_count: int, bits: int) -> Codec:
    for c in CODECS:
        if (c.family, c.state_count, c.bits) == (family, state_count, bits):
            return c
    raise KeyError((family, state_count, bits))
