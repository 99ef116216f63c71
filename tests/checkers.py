"""ctypes doors onto the CPU checkers under oracle/ — TEST INFRASTRUCTURE ONLY.

  oracle  = oracle/_build/libhsrans_oracle.so  (C restatement, oracle/hsrans_oracle.c)
  ref     = oracle/_ref/libhsrans_ref.so       (the unmodified reference compiled from /root/reference/src;
                                                present when it was built in the container — it travels to the GPU box)

`encode()` produces streams with the reference's own encoders when `ref` is available; otherwise raw streams come
from the oracle's scalar encoder twin and block_/mt_ streams from the committed golden fixtures only.
Nothing in the product imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libhsrans_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libhsrans_ref.so")
SYNTH_SO = os.path.join(ROOT, "oracle", "_build", "libhsr_synth.so")
RAW, BLOCK, MT, RAW32BLK = 0, 1, 2, 3
IMPL_SCALAR, IMPL_AVX2, IMPL_AVX512, IMPL_POOL = 0, 1, 2, 3

_oracle = None
_ref = None


class OracleHist(C.Structure):
    _fields_ = [("symbolCount", C.c_uint16 * 256), ("cumul", C.c_uint16 * 256)]


class OracleBlock(C.Structure):
    _fields_ = [("inOffset", C.c_uint64), ("outOffset", C.c_uint64), ("size", C.c_uint64), ("kind", C.c_uint64)]


def oracle() -> C.CDLL:
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
        lib = C.CDLL(ORACLE_SO)
        vp, sz, u32 = C.c_void_p, C.c_size_t, C.c_uint32
        lib.hsro_decode.restype = sz
        lib.hsro_decode.argtypes = [u32, u32, u32, vp, sz, vp, sz]
        lib.hsro_encode_raw.restype = sz
        lib.hsro_encode_raw.argtypes = [u32, u32, vp, sz, vp, sz, C.POINTER(OracleHist)]
        lib.hsro_capacity.restype = sz
        lib.hsro_capacity.argtypes = [u32, u32, sz]
        lib.hsro_make_hist.restype = None
        lib.hsro_make_hist.argtypes = [C.POINTER(OracleHist), vp, sz, u32]
        lib.hsro_normalize_hist.restype = None
        lib.hsro_normalize_hist.argtypes = [C.POINTER(OracleHist), vp, sz, u32]
        lib.hsro_observe_hist.restype = None
        lib.hsro_observe_hist.argtypes = [vp, vp, sz]
        lib.hsro_idx2idx.restype = u32
        lib.hsro_idx2idx.argtypes = [u32]
        lib.hsro_mt_walk.restype = sz
        lib.hsro_mt_walk.argtypes = [u32, vp, sz, C.POINTER(OracleBlock), sz]
        _oracle = lib
    return _oracle


_synth = None


def synth_zipf(n: int, s: float = 1.0, seed: int = 42, segment_bytes: int = 0) -> np.ndarray:
    """Deterministic Zipf(s) bytes from the generator built as a library of its own (oracle/Makefile): the same
    source as the product's hsr_synth_zipf, without mapping the product library (bench.py --impl reference)."""
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
        _synth = C.CDLL(SYNTH_SO)
        _synth.hsr_synth_zipf.restype = C.c_int
        _synth.hsr_synth_zipf.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_uint64, C.c_size_t]
    out = np.empty(n, np.uint8)
    if _synth.hsr_synth_zipf(out.ctypes.data, n, float(s), int(seed), int(segment_bytes)) != 0:
        raise RuntimeError("hsr_synth_zipf failed")
    return out


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libhsrans_ref.so is not built (needs /root/reference; `make -C oracle ref`)")
        lib = C.CDLL(REF_SO)
        vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
        lib.hsref_capacity.restype = sz
        lib.hsref_capacity.argtypes = [i, i, sz]
        lib.hsref_encode.restype = sz
        lib.hsref_encode.argtypes = [i, i, i, vp, sz, vp, sz]
        lib.hsref_decode.restype = sz
        lib.hsref_decode.argtypes = [i, i, i, i, vp, sz, vp, sz]
        lib.hsref_pool_create.restype = i
        lib.hsref_pool_create.argtypes = [i]
        lib.hsref_pool_threads.restype = i
        lib.hsref_pool_destroy.restype = None
        lib.hsref_set_max_simd.restype = None
        lib.hsref_set_max_simd.argtypes = [i]
        lib.hsref_has_avx2.restype = i
        lib.hsref_has_avx512.restype = i
        lib.hsref_cpu_name.restype = C.c_char_p
        lib.hsref_make_hist.restype = None
        lib.hsref_make_hist.argtypes = [vp, sz, i, vp, vp]
        lib.hsref_normalize_hist.restype = None
        lib.hsref_normalize_hist.argtypes = [vp, sz, i, vp, vp]
        lib.hsref_observe_hist.restype = None
        lib.hsref_observe_hist.argtypes = [vp, sz, vp]
        _ref = lib
    return _ref


def _u8(a) -> np.ndarray:
    a = np.frombuffer(a, dtype=np.uint8) if not isinstance(a, np.ndarray) else a
    return np.ascontiguousarray(a, dtype=np.uint8)


def oracle_capacity(family: int, states: int, n: int) -> int:
    return oracle().hsro_capacity(family, states, n)


def oracle_make_hist(data, bits: int):
    d = _u8(data)
    h = OracleHist()
    oracle().hsro_make_hist(C.byref(h), d.ctypes.data, d.size, bits)
    return np.array(h.symbolCount, dtype=np.uint16), np.array(h.cumul, dtype=np.uint16)


def oracle_normalize_hist(hist_u32: np.ndarray, data_bytes: int, bits: int):
    hist = np.ascontiguousarray(hist_u32, dtype=np.uint32)
    h = OracleHist()
    oracle().hsro_normalize_hist(C.byref(h), hist.ctypes.data, data_bytes, bits)
    return np.array(h.symbolCount, dtype=np.uint16), np.array(h.cumul, dtype=np.uint16)


def oracle_encode_raw(states: int, bits: int, data) -> np.ndarray:
    d = _u8(data)
    h = OracleHist()
    oracle().hsro_make_hist(C.byref(h), d.ctypes.data, d.size, bits)
    cap = oracle_capacity(RAW, states, d.size)
    out = np.zeros(cap, np.uint8)
    n = oracle().hsro_encode_raw(states, bits, d.ctypes.data, d.size, out.ctypes.data, cap, C.byref(h))
    if n == 0:
        raise RuntimeError("oracle raw encode failed")
    return out[:n].copy()


def oracle_decode(family: int, states: int, bits: int, stream, out_capacity: int):
    s = _u8(stream)
    out = np.full(max(out_capacity, 1), 0xCC, np.uint8)
    n = oracle().hsro_decode(family, states, bits, s.ctypes.data, s.size, out.ctypes.data, out_capacity)
    return n, out


def oracle_mt_walk(states: int, stream):
    s = _u8(stream)
    cnt = oracle().hsro_mt_walk(states, s.ctypes.data, s.size, None, 0)
    if cnt == C.c_size_t(-1).value:
        return None
    arr = (OracleBlock * max(cnt, 1))()
    oracle().hsro_mt_walk(states, s.ctypes.data, s.size, arr, cnt)
    return [(b.inOffset, b.outOffset, b.size, b.kind) for b in arr[:cnt]]


class RefEncoderOverflow(RuntimeError):
    """The reference encoder wrote outside the capacity its own *_capacity() promises (see ref_encode)."""


def ref_encode(family: int, states: int, bits: int, data) -> np.ndarray:
    """Runs the reference's encoder. Quirk found with the randomised campaign (scripts/gpu_soak.py): on barely
    compressible input at few probability bits the encoders need more than `*_capacity(n)` bytes — the raw one then
    memmoves header + words past the end of the caller's buffer (src/rANS32x32_16w.cpp:152-156; its capacity,
    :10-13, allows only n + 688 bytes) and can even walk below the buffer start. The buffer handed over here is
    therefore padded on both sides with guard bytes; a touched guard raises RefEncoderOverflow instead of
    corrupting the heap, and such a stream is not used."""
    d = _u8(data)
    cap = ref().hsref_capacity(family, states, d.size)
    front, back = max(4096, d.size // 4), 8192
    phys = np.full(front + cap + back, 0xAB, np.uint8)
    n = ref().hsref_encode(family, states, bits, d.ctypes.data, d.size, phys.ctypes.data + front, cap)
    if not (np.all(phys[:front] == 0xAB) and np.all(phys[front + cap:] == 0xAB)) or n > cap:
        raise RefEncoderOverflow(f"reference encoder overran its capacity (family {family}, N {states}, bits {bits}, n {d.size})")
    if n == 0:
        raise RuntimeError(f"reference encoder failed (family {family}, N {states}, bits {bits}, n {d.size})")
    return phys[front: front + n].copy()


def ref_decode(family: int, states: int, bits: int, stream, out_capacity: int, impl: int = IMPL_SCALAR):
    s = _u8(stream)
    padded = np.zeros(s.size + 128, np.uint8)  # the AVX decoders over-read up to 64 B (src/rANS32x32_16w.cpp:1226)
    padded[: s.size] = s
    out = np.full(max(out_capacity, 1) + 64, 0xCC, np.uint8)
    n = ref().hsref_decode(family, states, bits, impl, padded.ctypes.data, s.size, out.ctypes.data, out_capacity)
    return n, out


def ref_make_hist(data, bits: int):
    d = _u8(data)
    cnt, cum = np.zeros(256, np.uint16), np.zeros(256, np.uint16)
    ref().hsref_make_hist(d.ctypes.data, d.size, bits, cnt.ctypes.data, cum.ctypes.data)
    return cnt, cum


def ref_normalize_hist(hist_u32, data_bytes: int, bits: int):
    hist = np.ascontiguousarray(hist_u32, dtype=np.uint32)
    cnt, cum = np.zeros(256, np.uint16), np.zeros(256, np.uint16)
    ref().hsref_normalize_hist(hist.ctypes.data, data_bytes, bits, cnt.ctypes.data, cum.ctypes.data)
    return cnt, cum


def encode(family: int, states: int, bits: int, data) -> np.ndarray:
    """Stream producer: the reference's own encoder when built; the oracle's raw twin otherwise."""
    if have_ref():
        return ref_encode(family, states, bits, data)
    if family == RAW:
        return oracle_encode_raw(states, bits, data)
    raise RuntimeError("block_/mt_/32blk streams need oracle/_ref (the reference encoders) or the golden fixtures")
