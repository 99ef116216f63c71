"""ctypes binding of libhsrans_b200.so (include/hsrans_b200.h). Plain pointers and sizes, no torch types."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

HSR_RAW, HSR_BLOCK, HSR_MT, HSR_RAW32BLK = 0, 1, 2, 3
FAMILY_NAMES = {HSR_RAW: "rANS32x{N}_16w", HSR_BLOCK: "block_rANS32x{N}_16w", HSR_MT: "mt_rANS32x{N}_16w",
                HSR_RAW32BLK: "rANS32x{N}_32blk_16w"}
HSR_OUT_SHARD_LOCAL = 1


class HsrError(RuntimeError):
    pass


class Block(C.Structure):
    """hsr_block_t — one unit of kernel work (an mt_ block, a fill, or a raw stream)."""
    _fields_ = [("inOffset", C.c_uint64), ("inEnd", C.c_uint64), ("outOffset", C.c_uint64), ("count", C.c_uint64),
                ("kind", C.c_uint32), ("symbol", C.c_uint32), ("tail", C.c_uint32), ("reserved", C.c_uint32)]


def lib_path() -> str:
    # HSRANS_B200_LIB lets the A/B scripts run the test-suite against a build variant (scripts/build_variant.sh)
    return os.environ.get("HSRANS_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhsrans_b200.so")


_lib = None

# name -> (restype, argtypes); this is also the list the CPU test checks against include/hsrans_b200.h
SIGNATURES = {
    "hsr_version": (C.c_int, []),
    "hsr_device_count": (C.c_int, []),
    "hsr_last_error": (C.c_char_p, []),
    "hsr_set_option": (C.c_int, [C.c_char_p, C.c_long]),
    "hsr_get_option": (C.c_long, [C.c_char_p]),
    "hsr_capacity": (C.c_size_t, [C.c_int, C.c_int, C.c_size_t]),
    "hsr_host_alloc": (C.c_void_p, [C.c_size_t]),
    "hsr_host_free": (None, [C.c_void_p]),
    "hsr_decode": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "hsr_decode_mt_multi": (C.c_size_t, [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_int), C.c_int]),
    "hsr_decode_mt_shard": (C.c_size_t, [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                         C.POINTER(C.c_size_t)]),
    "hsr_set_device": (C.c_int, [C.c_int]),
    "hsr_decode_batch": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "hsr_mt_index": (C.c_long, [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(Block), C.c_size_t]),
    "hsr_mt_partition": (C.c_int, [C.POINTER(Block), C.c_size_t, C.c_int, C.POINTER(C.c_size_t)]),
    "hsr_stream_upload": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "hsr_stream_from_device": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "hsr_stream_from_device_indexed": (C.c_void_p, [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "hsr_stream_upload_batch": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "hsr_stream_free": (None, [C.c_void_p]),
    "hsr_stream_decoded_length": (C.c_uint64, [C.c_void_p]),
    "hsr_stream_shard_out_offset": (C.c_uint64, [C.c_void_p]),
    "hsr_stream_shard_out_bytes": (C.c_uint64, [C.c_void_p]),
    "hsr_stream_shard_in_bytes": (C.c_uint64, [C.c_void_p]),
    "hsr_stream_units": (C.c_uint64, [C.c_void_p]),
    "hsr_stream_index_ms": (C.c_double, [C.c_void_p]),
    "hsr_stream_copy_index": (C.c_int, [C.c_void_p, C.POINTER(Block), C.c_size_t]),
    "hsr_stream_decode_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p]),
    "hsr_stream_status": (C.c_uint, [C.c_void_p]),
    "hsr_make_hist": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "hsr_observe_hist_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "hsr_normalize_hist_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hsr_make_hist_segments_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "hsr_encode_mt": (C.c_size_t, [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t]),
    "hsr_encode_mt_device": (C.c_size_t, [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "hsr_encode_mt_device_indexed": (C.c_size_t, [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                                   C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]),
    "hsr_encode_mt_index_bound": (C.c_size_t, [C.c_int, C.c_size_t, C.c_size_t]),
    "hsr_encode_mt_bound": (C.c_size_t, [C.c_int, C.c_size_t, C.c_size_t]),
    "hsr_encode_mt_policy": (C.c_size_t, [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t]),
    "hsr_encode_mt_policy_device": (C.c_size_t, [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "hsr_synth_zipf": (C.c_int, [C.c_void_p, C.c_size_t, C.c_double, C.c_uint64, C.c_size_t]),
}


def lib() -> C.CDLL:
    """Loads the CUDA shared object. Fails loudly if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise HsrError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(or `make -C hypersonic-rans_b200`); there is no CPU fallback")
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().hsr_last_error().decode("utf-8", "replace")


def version() -> int:
    return lib().hsr_version()


def device_count() -> int:
    return lib().hsr_device_count()


def set_option(key: str, value: int) -> None:
    if lib().hsr_set_option(key.encode(), int(value)) != 0:
        raise HsrError(f"unknown option or value: {key}={value}")


def get_option(key: str) -> int:
    return lib().hsr_get_option(key.encode())


def capacity(family: int, state_count: int, n: int) -> int:
    return lib().hsr_capacity(family, state_count, n)


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _as_u8(buf) -> np.ndarray:
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    if a.dtype != np.uint8 or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


class HostBuffer:
    """Page-locked host buffer exposed as a numpy uint8 array (hsr_host_alloc / hsr_host_free)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self.ptr = lib().hsr_host_alloc(self.nbytes)
        if not self.ptr:
            raise HsrError(f"hsr_host_alloc({nbytes}) failed: {last_error()}")
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(max(self.nbytes, 1),))[: self.nbytes]

    def free(self) -> None:
        if self.ptr:
            self.array = None
            lib().hsr_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def host_alloc(nbytes: int) -> HostBuffer:
    return HostBuffer(nbytes)


def decode(family: int, state_count: int, bits: int, data, out_capacity: int, out: Optional[np.ndarray] = None):
    """Host-pointer decode (the drop-in for the reference's decodeFunc). Returns (decoded_length, out_array)."""
    src = _as_u8(data)
    if out is None:
        out = np.full(max(out_capacity, 1), 0xCC, dtype=np.uint8)  # the reference harness poisons with 0xCC (main.cpp:860)
    n = lib().hsr_decode(family, state_count, bits, _ptr(src), src.size, _ptr(out), out_capacity)
    return n, out


def decode_mt_multi(state_count: int, bits: int, data, out_capacity: int, devices: Optional[Sequence[int]] = None,
                    device_count_: Optional[int] = None, out: Optional[np.ndarray] = None):
    src = _as_u8(data)
    if out is None:
        out = np.full(max(out_capacity, 1), 0xCC, dtype=np.uint8)
    if devices is not None:
        arr = (C.c_int * len(devices))(*devices)
        n = lib().hsr_decode_mt_multi(state_count, bits, _ptr(src), src.size, _ptr(out), out_capacity, arr, len(devices))
    else:
        n = lib().hsr_decode_mt_multi(state_count, bits, _ptr(src), src.size, _ptr(out), out_capacity, None,
                                      int(device_count_ or 1))
    return n, out


class BatchItem(C.Structure):
    """hsr_batch_item_t"""
    _fields_ = [("inOffset", C.c_uint64), ("inLength", C.c_uint64), ("outOffset", C.c_uint64), ("outCapacity", C.c_uint64)]


def decode_batch(family: int, state_count: int, bits: int, in_base, out_base: np.ndarray, items):
    """Many independent streams of one codec in one launch. `items` = [(inOffset, inLength, outOffset, outCapacity)].
    Returns (number decoded, per-stream decoded lengths)."""
    src = _as_u8(in_base)
    arr = (BatchItem * max(len(items), 1))(*[BatchItem(*map(int, it)) for it in items])
    lengths = np.zeros(max(len(items), 1), np.uint64)
    ok = lib().hsr_decode_batch(family, state_count, bits, _ptr(src), _ptr(out_base), arr, len(items), _ptr(lengths))
    return ok, lengths[: len(items)]


def mt_index(state_count: int, data) -> list:
    """Host walk of the mt_ header chain -> list of Block records (raises on a malformed chain)."""
    src = _as_u8(data)
    cnt = lib().hsr_mt_index(state_count, _ptr(src), src.size, None, 0)
    if cnt < 0:
        raise HsrError(f"hsr_mt_index: {last_error()}")
    arr = (Block * max(cnt, 1))()
    got = lib().hsr_mt_index(state_count, _ptr(src), src.size, arr, cnt)
    if got != cnt:
        raise HsrError(f"hsr_mt_index: {last_error()}")
    return [arr[i] for i in range(cnt)]


def mt_partition(blocks: Sequence[Block], parts: int) -> list:
    arr = (Block * max(len(blocks), 1))(*blocks)
    first = (C.c_size_t * (parts + 1))()
    if lib().hsr_mt_partition(arr, len(blocks), parts, first) != 0:
        raise HsrError("hsr_mt_partition failed")
    return list(first)


class PreparedStream:
    """hsr_stream_t — compressed bytes resident in HBM plus the block index; decode launches are asynchronous."""

    def __init__(self, handle: int):
        if not handle:
            raise HsrError(f"stream preparation failed: {last_error()}")
        self.handle = handle

    @classmethod
    def upload(cls, family: int, state_count: int, bits: int, data, shard: int = 0, shards: int = 1) -> "PreparedStream":
        src = _as_u8(data)
        return cls(lib().hsr_stream_upload(family, state_count, bits, _ptr(src), src.size, shard, shards))

    @classmethod
    def upload_batch(cls, family: int, state_count: int, bits: int, in_base, items) -> "PreparedStream":
        src = _as_u8(in_base)
        arr = (BatchItem * max(len(items), 1))(*[BatchItem(*map(int, it)) for it in items])
        return cls(lib().hsr_stream_upload_batch(family, state_count, bits, _ptr(src), arr, len(items)))

    @classmethod
    def from_device(cls, family: int, state_count: int, bits: int, device_ptr: int, length: int) -> "PreparedStream":
        return cls(lib().hsr_stream_from_device(family, state_count, bits, device_ptr, length))

    @classmethod
    def from_device_indexed(cls, state_count: int, bits: int, device_ptr: int, length: int, d_index: int, num_units: int) -> "PreparedStream":
        """mt_ stream + the block table its producer wrote (encode_mt_device_indexed): no chain walk."""
        return cls(lib().hsr_stream_from_device_indexed(state_count, bits, device_ptr, length, d_index, num_units))

    decoded_length = property(lambda self: lib().hsr_stream_decoded_length(self.handle))
    shard_out_offset = property(lambda self: lib().hsr_stream_shard_out_offset(self.handle))
    shard_out_bytes = property(lambda self: lib().hsr_stream_shard_out_bytes(self.handle))
    shard_in_bytes = property(lambda self: lib().hsr_stream_shard_in_bytes(self.handle))
    units = property(lambda self: lib().hsr_stream_units(self.handle))
    index_ms = property(lambda self: lib().hsr_stream_index_ms(self.handle))

    def index(self) -> list:
        n = self.units
        arr = (Block * max(n, 1))()
        got = lib().hsr_stream_copy_index(self.handle, arr, n)
        return [arr[i] for i in range(max(got, 0))]

    def decode_async(self, out_ptr: int, out_capacity: int, cuda_stream: int = 0, shard_local: bool = False) -> int:
        rc = lib().hsr_stream_decode_async(self.handle, out_ptr, out_capacity, HSR_OUT_SHARD_LOCAL if shard_local else 0,
                                           cuda_stream)
        if rc < 0:
            raise HsrError(f"hsr_stream_decode_async: {last_error()}")
        return rc

    def status(self) -> int:
        return lib().hsr_stream_status(self.handle)

    def free(self) -> None:
        if self.handle:
            lib().hsr_stream_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def make_hist(data, bits: int):
    """Device make_hist (src/hist.cpp:217-222). Returns (symbolCount[256], cumul[256]) as uint16 arrays."""
    src = _as_u8(data)
    cnt = np.zeros(256, np.uint16)
    cum = np.zeros(256, np.uint16)
    rc = lib().hsr_make_hist(_ptr(src), src.size, bits, _ptr(cnt), _ptr(cum))
    if rc != 0:
        raise HsrError(f"hsr_make_hist failed ({rc}): {last_error()}")
    return cnt, cum


def observe_hist_device(d_data: int, size: int, d_hist: int, cuda_stream: int = 0) -> int:
    return lib().hsr_observe_hist_device(d_data, size, d_hist, cuda_stream)


def normalize_hist_device(d_hist: int, data_bytes: int, bits: int, d_count: int, d_cumul: int, cuda_stream: int = 0) -> int:
    return lib().hsr_normalize_hist_device(d_hist, data_bytes, bits, d_count, d_cumul, cuda_stream)


def make_hist_segments_device(d_data: int, size: int, segment_bytes: int, bits: int, d_counts: int, cuda_stream: int = 0) -> int:
    return lib().hsr_make_hist_segments_device(d_data, size, segment_bytes, bits, d_counts, cuda_stream)


def encode_mt(state_count: int, bits: int, data, block_size: int = 0) -> np.ndarray:
    """Device mt_ encoder (hsr_encode_mt): returns the compressed stream as a uint8 array; raises on failure."""
    src = _as_u8(data)
    bound = lib().hsr_encode_mt_bound(state_count, src.size, block_size)
    if bound == 0:
        raise HsrError("hsr_encode_mt_bound: unsupported arguments")
    out = np.empty(bound, np.uint8)
    n = lib().hsr_encode_mt(state_count, bits, _ptr(src), src.size, _ptr(out), bound, block_size)
    if n == 0:
        raise HsrError(f"hsr_encode_mt failed: {last_error()}")
    return out[:n].copy()


def encode_mt_policy(state_count: int, bits: int, data, max_block_size: int = 0) -> np.ndarray:
    """Device mt_ encoder with the block-split policy decided on the device (hsr_encode_mt_policy)."""
    src = _as_u8(data)
    bound = lib().hsr_encode_mt_bound(state_count, src.size, 0)
    if bound == 0:
        raise HsrError("hsr_encode_mt_bound: unsupported arguments")
    out = np.empty(bound, np.uint8)
    n = lib().hsr_encode_mt_policy(state_count, bits, _ptr(src), src.size, _ptr(out), bound, max_block_size)
    if n == 0:
        raise HsrError(f"hsr_encode_mt_policy failed: {last_error()}")
    return out[:n].copy()


def encode_mt_policy_device(state_count: int, bits: int, d_in: int, length: int, d_out: int, out_capacity: int, max_block_size: int = 0,
                            cuda_stream: int = 0) -> int:
    return lib().hsr_encode_mt_policy_device(state_count, bits, d_in, length, d_out, out_capacity, max_block_size, cuda_stream)


def encode_mt_device(state_count: int, bits: int, d_in: int, length: int, d_out: int, out_capacity: int, block_size: int = 0,
                     cuda_stream: int = 0) -> int:
    return lib().hsr_encode_mt_device(state_count, bits, d_in, length, d_out, out_capacity, block_size, cuda_stream)


def encode_mt_device_indexed(state_count: int, bits: int, d_in: int, length: int, d_out: int, out_capacity: int, d_index: int,
                             index_capacity: int, block_size: int = 0, policy: bool = False, cuda_stream: int = 0):
    """Device encode that also writes the decoder's unit table to d_index. Returns (compressed_bytes, num_units)."""
    n_units = C.c_size_t(0)
    comp = lib().hsr_encode_mt_device_indexed(state_count, bits, d_in, length, d_out, out_capacity, block_size, 1 if policy else 0,
                                              d_index, index_capacity, C.byref(n_units), cuda_stream)
    return comp, n_units.value


def encode_mt_index_bound(state_count: int, length: int, block_size: int = 0) -> int:
    return lib().hsr_encode_mt_index_bound(state_count, length, block_size)


def encode_mt_bound(state_count: int, length: int, block_size: int = 0) -> int:
    return lib().hsr_encode_mt_bound(state_count, length, block_size)


def synth_zipf(n: int, s: float = 1.0, seed: int = 42, segment_bytes: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
    """Deterministic Zipf(s) bytes; segment_bytes=0 -> 'iid', 65536 -> 'pw64k' (SURVEY.md §8d)."""
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    if lib().hsr_synth_zipf(_ptr(out), n, float(s), int(seed), int(segment_bytes)) != 0:
        raise HsrError("hsr_synth_zipf failed")
    return out
