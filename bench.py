#!/usr/bin/env python
"""bench.py — decoded GB/s of the B200 rANS decoder on BASELINE.json's headline configuration.

Workload (config 4 of BASELINE.json, the one its metric "per GPU & whole box (1/2/4/8 B200)" is quoted on):
mt_rANS32x64_16w, 15 probability bits, ONE 1,000,000,000-byte synthetic Zipf(s=1) stream, `pw64k` shape (rank->byte
permutation re-drawn every 64 KiB so the reference's mt_ encoder emits ~15 k independent blocks per GB),
encoded by the reference's own encoder. At N > 1 that SAME stream is sharded over the N ranks by contiguous block range
(hsr_stream_upload(shard = rank, shards = N), the GPU analogue of the reference spreading one stream over its thread
pool, src/mt_rANS32x64_16w_decode.cpp:137-265): STRONG scaling, no data-path collective. The decoded shards are
assembled on rank 0 over NCCL once, untimed, and all n bytes are compared with the original before a number is printed.

A "step" is one pass of the decode path over the whole stream (all ranks together):
  value  kernel path, compressed shard + block index already resident in HBM (hsr_stream_decode_async), CUDA events on
         the launching stream around exactly K steps, max over ranks; value = n / that. Consecutive steps overlap the
         way the library launches them (programmatic dependent launch: the next decode's warps fill the SM slots the
         previous one frees while it drains); the strictly serialised figure is reported beside it.
  e2e    the drop-in host-pointer call with pinned HOST buffers — hsr_decode() at N = 1, hsr_decode_mt_multi() over the
         N devices from rank 0 at N > 1 (one host buffer in, one out, like the reference's thread-pool decoder): header
         walk, H2D, kernels, D2H inside the timed region, every step.
  roofline       algorithmic bytes (compressed in + decoded out) / mean kernel step duration vs the measured HBM peak.
  cpu_baseline   the reference's own decoders (oracle/_ref, compiled unmodified) on this box's host cores.

The input stream is produced by the reference's own, unmodified encoder (oracle/_ref) during set-up, as the north star
requires; that, the synthetic-byte generator (built as a library of its own under oracle/_build) and the CPU baseline
are the only places this file executes anything under oracle/. The timed GPU path never does, and `--impl reference`
never maps the product library.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
"""
from __future__ import annotations

import argparse
import faulthandler
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FAMILY_RAW, FAMILY_BLOCK, FAMILY_MT, FAMILY_RAW32BLK = 0, 1, 2, 3
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


faulthandler.enable()  # a native abort still leaves the Python stack on stderr


def _stage(msg):
    if os.environ.get("HSR_BENCH_TRACE"):
        print(f"[bench +{time.time() - _T0:7.2f}s] {msg}", file=sys.stderr, flush=True)


_T0 = time.time()


# The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints "NCCL version ..." from C, past
# sys.stdout), so the real stdout is kept aside for that line and file descriptor 1 is pointed at stderr for everything
# else this process or its libraries print.
_JSON_OUT = None


def _claim_stdout():
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def _emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1_000_000_000, help="decoded bytes of the stream (whole job)")
    ap.add_argument("--bits", type=int, default=15)
    ap.add_argument("--states", type=int, default=64)
    ap.add_argument("--shape", default="pw64k", choices=["pw64k", "iid"])
    ap.add_argument("--zipf", type=float, default=1.0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="skip e2e / index / CPU legs (profiling runs)")
    ap.add_argument("--table", type=int, default=0, help="hsr_set_option table: 0 auto, 1 bitmap-rank, 2 packed, 3 wide (bits >= 13)")
    ap.add_argument("--ctas-per-sm", type=int, default=0, help="cap resident one-warp CTAs per SM (occupancy experiments)")
    ap.add_argument("--no-overlap", action="store_true", help="hsr_set_option overlap=0: strictly serial launches everywhere")
    ap.add_argument("--headline-only", action="store_true", help="skip the other-configs leg (profiling runs: only the headline kernel launches)")
    ap.add_argument("--extra", action="store_true", help="also measure the device encoder and the histogram kernels")
    ap.add_argument("--weak", action="store_true", help="N > 1: also time N independent per-rank streams (round 1's weak-scaling figure)")
    return ap.parse_args()


def workload_name(a):
    return (f"mt_rANS32x{a.states}_16w {a.bits}-bit decode of ONE {a.size:,}-byte Zipf(s={a.zipf:g}) {a.shape} stream, "
            f"sharded by contiguous block range over the GPUs (BASELINE config 4)")


def make_data(a, seed=42, size=None, shape=None):
    import checkers as ck
    shape = shape or a.shape
    return ck.synth_zipf(size or a.size, a.zipf, seed=seed, segment_bytes=65536 if shape == "pw64k" else 0)


def ref_encode(a, data, family=FAMILY_MT, states=None, bits=None):
    import checkers as ck
    if not ck.have_ref():
        raise RuntimeError("oracle/_ref/libhsrans_ref.so is missing: the input streams must come from the reference's "
                           "own encoder (build it in the container with `make -C oracle ref`; it travels with gpurun)")
    return ck.ref_encode(family, states or a.states, bits or a.bits, data)


class ClockSampler:
    """Samples SM clocks and throttle reasons of one GPU with NVML while a timed region runs."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def _time_ref(lib, family, states, bits, impl, padded, comp, out, n, reps):
    best = None
    for _ in range(reps + 1):  # first run is the dry run (src/main.cpp:862-866)
        t0 = time.perf_counter()
        got = lib.hsref_decode(family, states, bits, impl, padded.ctypes.data, comp, out.ctypes.data, n)
        dt = time.perf_counter() - t0
        if got != n:
            raise RuntimeError("reference decoder failed on its own stream")
        best = dt if best is None else min(best, dt)
    return n / best / 1e9


def cpu_baseline(a, stream, n, reps=3):
    """The reference's own mt_ decoders on the host cores: thread pool (all cores) and single thread — once capped at
    AVX2 (the north star's "fastest AVX2 decoder": hsref_set_max_simd(1) clears the AVX-512 feature flags exactly like
    the reference's --max-simd avx2, src/main.cpp:482-510) and once with whatever the CPU offers."""
    import checkers as ck
    lib = ck.ref()
    cores = os.cpu_count() or 1
    padded = np.zeros(stream.size + 128, np.uint8)
    padded[: stream.size] = stream
    out = np.empty(n + 64, np.uint8)
    res = {}
    for label, level in (("avx2", 1), ("native", 0)):
        lib.hsref_set_max_simd(level)
        threads = lib.hsref_pool_create(0)  # hardware_concurrency() - 1 workers + the calling thread (src/main.cpp:167)
        res[label] = {"pool": _time_ref(lib, FAMILY_MT, a.states, a.bits, ck.IMPL_POOL, padded, stream.size, out, n, reps),
                      "single": _time_ref(lib, FAMILY_MT, a.states, a.bits, ck.IMPL_SCALAR, padded, stream.size, out, n, reps)}
        lib.hsref_pool_destroy()
    lib.hsref_set_max_simd(0)
    return {"value": round(res["avx2"]["pool"], 4), "unit": "GB/s", "cores": threads + 1, "kind": "reference",
            "sample": f"whole {n:,}-byte stream, best of {reps} after a dry run, mt_rANS32x{a.states}_16w_decode_mt_{a.bits} "
                      f"with {threads} pool threads + caller, SIMD capped at AVX2",
            "single_thread_avx2_gbs": round(res["avx2"]["single"], 4),
            "pool_native_simd_gbs": round(res["native"]["pool"], 4), "single_thread_native_simd_gbs": round(res["native"]["single"], 4),
            "native_simd": "avx512" if lib.hsref_has_avx512() else ("avx2" if lib.hsref_has_avx2() else "scalar"),
            "host_cores": cores, "cpu": lib.hsref_cpu_name().decode(errors="replace").strip()}


def run_reference(a, rank, world):
    """--impl reference: the reference's own CPU implementation (thread-pool mt_ decoder, all host threads), rank 0 only.
    Does not import the product package: inputs come from the stand-alone generator under oracle/_build."""
    if rank != 0:
        return
    import checkers as ck
    data = make_data(a)
    stream = ref_encode(a, data)
    n = data.size
    lib = ck.ref()
    threads = lib.hsref_pool_create(0)
    padded = np.zeros(stream.size + 128, np.uint8)
    padded[: stream.size] = stream
    out = np.empty(n + 64, np.uint8)

    def step():
        got = lib.hsref_decode(FAMILY_MT, a.states, a.bits, ck.IMPL_POOL, padded.ctypes.data, stream.size, out.ctypes.data, n)
        if got != n:
            raise RuntimeError("reference decoder failed")

    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = time.perf_counter() - t0
    assert np.array_equal(out[:n], data)
    gbs = n * a.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": "decoded_GBps", "value": round(gbs, 4), "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": round(dt / a.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "strong" if a.gpus > 1 else "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(a), "codec": f"mt_rANS32x{a.states}_16w", "bits": a.bits, "shape": a.shape,
                   "decoded_bytes": n, "compressed_bytes": int(stream.size),
                   "note": "reference CPU thread-pool decoder (mt_rANS32xNN_16w_decode_mt, native SIMD dispatch) on the whole stream, rank 0"},
        "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": threads + 1, "kind": "reference",
                         "sample": f"whole {n:,}-byte stream per step, mt_rANS32x{a.states}_16w_decode_mt_{a.bits}"},
        "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "loaded_libraries": sorted({os.path.basename(l.split()[-1]) for l in open("/proc/self/maps") if "libhs" in l}),
    }
    _emit(line)


def time_decode(torch, ps, out_ptr, cap, reps, shard_local=False):
    st = torch.cuda.current_stream().cuda_stream
    ps.decode_async(out_ptr, cap, st, shard_local)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ps.decode_async(out_ptr, cap, st, shard_local)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def other_configs(pkg, torch, a, peak, heavy):
    """BASELINE configs 1-3 (single-recurrence codecs: one stream = one warp, latency-bound by construction, SURVEY.md
    finding 1), the same codecs as a batch of many streams, and config 4's low-parallelism shape (iid). With `heavy`
    also the device-encoder and histogram measurements."""
    import checkers as ck
    lib = ck.ref()
    res = {}
    n = 100_000_000
    data = make_data(a, seed=42, size=n, shape="iid")
    out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    for label, fam, states, bits in (("rANS32x64_16w_12_raw", 0, 64, 12), ("block_rANS32x32_16w_10", 1, 32, 10),
                                     ("rANS32x32_16w_11_raw", 0, 32, 11), ("rANS32x16_16w_12_raw", 0, 16, 12),
                                     ("rANS32x32_32blk_16w_15_raw", 3, 32, 15)):
        stream = ck.ref_encode(fam, states, bits, data)
        ps = pkg.PreparedStream.upload(fam, states, bits, stream)
        ms = time_decode(torch, ps, out.data_ptr(), n, 1)
        ok = ps.status() == 0 and bool(np.array_equal(out[:n].cpu().numpy(), data))
        # CPU beside it, one thread: the fastest AVX2 decoder (raw codecs: the explicit xmmShfl2 AVX2 entry point; block_:
        # the dispatching decoder with the AVX-512 flags cleared), and what the dispatcher picks natively
        lib.hsref_set_max_simd(1)
        t0 = time.perf_counter()
        ck.ref_decode(fam, states, bits, stream, n, ck.IMPL_AVX2)
        cpu_avx2 = n / (time.perf_counter() - t0) / 1e9
        lib.hsref_set_max_simd(0)
        entry = {"gpu_decoded_GBps": round(n / ms / 1e6, 4), "gpu_ms": round(ms, 3), "bit_exact": ok,
                 "frac": round((n + stream.size) / ms / 1e6 / peak, 6), "cpu_avx2_1thread_GBps": round(cpu_avx2, 4),
                 "streams": 1, "warps": 1}
        if fam == FAMILY_BLOCK or (fam == FAMILY_RAW and states != 16 and not (states == 32 and bits > 12)):
            t0 = time.perf_counter()
            got, _ = ck.ref_decode(fam, states, bits, stream, n, ck.IMPL_AVX512)
            if got == n:
                entry["cpu_avx512_1thread_GBps"] = round(n / (time.perf_counter() - t0) / 1e9, 4)
        res[label] = entry
        ps.free()
    del out

    # the same two codecs with many independent streams in one launch (hsr_stream_upload_batch): one warp per stream
    k_streams, each = 2368, 400_000
    data = make_data(a, seed=43, size=k_streams * each, shape="iid")
    for label, fam, states, bits in (("rANS32x64_16w_12_raw", 0, 64, 12), ("block_rANS32x32_16w_10", 1, 32, 10)):
        parts, items, pos = [], [], 0
        for k in range(k_streams):
            stream = ck.ref_encode(fam, states, bits, data[k * each:(k + 1) * each])
            pad = (-pos) % 16
            parts.append(np.zeros(pad, np.uint8)); pos += pad
            items.append((pos, stream.size, k * each, each))
            parts.append(stream); pos += stream.size
        in_base = np.concatenate(parts)
        ps = pkg.PreparedStream.upload_batch(fam, states, bits, in_base, items)
        total = k_streams * each
        out2 = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
        ms = time_decode(torch, ps, out2.data_ptr(), total, 5)
        ok = ps.status() == 0 and bool(np.array_equal(out2[:total].cpu().numpy(), data))
        res[label + "_batch"] = {"gpu_decoded_GBps": round(total / ms / 1e6, 2), "gpu_ms": round(ms, 3), "bit_exact": ok,
                                 "streams": k_streams, "bytes_per_stream": each, "compressed_bytes": int(in_base.size),
                                 "algorithmic_GBps": round((total + in_base.size) / ms / 1e6, 1),
                                 "frac": round((total + in_base.size) / ms / 1e6 / peak, 4)}
        ps.free()
        del out2
        # the same batch end to end through the host-pointer call (hsr_decode_batch, pinned buffers): pieces of the input
        # go up, groups of streams decode and come back as a pipeline
        hin, hout = pkg.host_alloc(in_base.size), pkg.host_alloc(total)
        hin.array[:] = in_base
        pkg.decode_batch(fam, states, bits, hin.array, hout.array, items)   # warm-up: scratch, events
        hout.array[:] = 0xCC
        t0 = time.perf_counter()
        n_ok, _ = pkg.decode_batch(fam, states, bits, hin.array, hout.array, items)
        dt = time.perf_counter() - t0
        res[label + "_batch"]["e2e_host_pointers"] = {"decoded_GBps": round(total / dt / 1e9, 2), "ms": round(dt * 1e3, 3),
                                                      "bit_exact": bool(n_ok == k_streams and np.array_equal(hout.array, data))}
        hin.free(); hout.free()

    # BASELINE config 4's low-parallelism shape: the reference encoder merges stationary (iid) data into ~32 MiB blocks
    # (src/mt_rANS32x64_16w_encode.cpp:207-213), i.e. ~60 independent warps of work per GB whatever the GPU
    data = make_data(a, seed=42, shape="iid")
    stream = ck.ref_encode(FAMILY_MT, a.states, a.bits, data)
    ps = pkg.PreparedStream.upload(FAMILY_MT, a.states, a.bits, stream)
    out4 = torch.empty(a.size + 64, dtype=torch.uint8, device="cuda")
    ms = time_decode(torch, ps, out4.data_ptr(), a.size, 2)
    ok = ps.status() == 0 and bool(np.array_equal(out4[:a.size].cpu().numpy(), data))
    blocks = ps.index()
    longest = max(blocks, key=lambda b: b.count)
    # the longest block as a stream of its own (header + its bytes; a chain's last block runs to the end of the input)
    lone = np.zeros(16 + (longest.inEnd - (longest.inOffset - 16)), np.uint8)
    lone[:8] = np.frombuffer(np.uint64(longest.count).tobytes(), np.uint8)
    lone[8:16] = np.frombuffer(np.uint64(lone.size).tobytes(), np.uint8)
    lone[16:] = stream[longest.inOffset - 16: longest.inEnd]
    ps1 = pkg.PreparedStream.upload(FAMILY_MT, a.states, a.bits, lone)
    ms1 = time_decode(torch, ps1, out4.data_ptr(), longest.count, 2)
    ok1 = ps1.status() == 0 and bool(np.array_equal(out4[:longest.count].cpu().numpy(), data[longest.outOffset: longest.outOffset + longest.count]))
    ps1.free()
    res["mt_iid_reference_encoded"] = {
        "gpu_decoded_GBps": round(a.size / ms / 1e6, 2), "gpu_ms": round(ms, 3), "blocks": int(ps.units), "bit_exact": ok,
        "frac": round((a.size + stream.size) / ms / 1e6 / peak, 5), "decoded_bytes": a.size, "compressed_bytes": int(stream.size),
        "longest_block": {"decoded_bytes": int(longest.count), "alone_ms": round(ms1, 3), "bit_exact": ok1,
                          "share_of_stream_kernel_time": round(ms1 / ms, 3)},
        "note": "one warp per block: the stream offers ~60 warps of parallelism; its kernel time IS its longest block on one warp"}
    ps.free()
    del out4
    if not heavy:
        return res

    # device-side producer (hsr_encode_mt_device) and the histogram kernels, on the same 1 GB of bytes
    n = a.size
    for shape, seg in (("pw64k", 65536), ("iid", 0)):
        data = make_data(a, seed=42, shape=shape)
        d_in = torch.from_numpy(data).cuda()
        bound = pkg.encode_mt_bound(a.states, n)
        d_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        comp = pkg.encode_mt_device(a.states, a.bits, d_in.data_ptr(), n, d_out.data_ptr(), bound, 0, st)   # warm-up, sizes scratch
        times = []
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            comp = pkg.encode_mt_device(a.states, a.bits, d_in.data_ptr(), n, d_out.data_ptr(), bound, 0, st)
            torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
        ps = pkg.PreparedStream.from_device(2, a.states, a.bits, d_out.data_ptr(), comp)
        out3 = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
        dec_ms = time_decode(torch, ps, out3.data_ptr(), n, 5)
        ok = comp > 0 and ps.status() == 0 and bool(torch.equal(out3[:n], d_in))
        walk_ms = ps.index_ms
        ps.free()
        # the same stream wrapped with the encoder's own block table: no chain walk
        cap = pkg.encode_mt_index_bound(a.states, n)
        d_idx = torch.empty(cap * 48, dtype=torch.uint8, device="cuda")
        comp_i, n_units = pkg.encode_mt_device_indexed(a.states, a.bits, d_in.data_ptr(), n, d_out.data_ptr(), bound, d_idx.data_ptr(), cap, 0, False, st)
        psi = pkg.PreparedStream.from_device_indexed(a.states, a.bits, d_out.data_ptr(), comp_i, d_idx.data_ptr(), n_units)
        psi.decode_async(out3.data_ptr(), n, st)
        torch.cuda.synchronize()
        ok_i = comp_i == comp and psi.status() == 0 and bool(torch.equal(out3[:n], d_in))
        res[f"device_encoder_{shape}"] = {"encode_GBps": round(n / min(times) / 1e9, 2), "encode_ms": round(min(times) * 1e3, 3),
                                          "compressed_bytes": int(comp), "blocks": int(n_units), "round_trip_bit_exact": ok,
                                          "decode_GBps_of_this_stream": round(n / dec_ms / 1e6, 2),
                                          "index_ms": {"device_walk": round(walk_ms, 3), "from_encoder_table": round(psi.index_ms, 3)},
                                          "indexed_round_trip_bit_exact": ok_i}
        psi.free()
        # the same bytes through the device block-split policy (hsr_encode_mt_policy_device, 256 KiB max blocks)
        comp_p = pkg.encode_mt_policy_device(a.states, a.bits, d_in.data_ptr(), n, d_out.data_ptr(), bound, 0, st)
        times_p = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            comp_p = pkg.encode_mt_policy_device(a.states, a.bits, d_in.data_ptr(), n, d_out.data_ptr(), bound, 0, st)
            torch.cuda.synchronize(); times_p.append(time.perf_counter() - t0)
        ps = pkg.PreparedStream.from_device(2, a.states, a.bits, d_out.data_ptr(), comp_p)
        dec_p = time_decode(torch, ps, out3.data_ptr(), n, 5)
        ok_p = comp_p > 0 and ps.status() == 0 and bool(torch.equal(out3[:n], d_in))
        res[f"device_encoder_policy_{shape}"] = {"encode_GBps": round(n / min(times_p) / 1e9, 2), "encode_ms": round(min(times_p) * 1e3, 3),
                                                 "compressed_bytes": int(comp_p), "blocks": int(ps.units), "max_block_bytes": 262144,
                                                 "round_trip_bit_exact": ok_p, "decode_GBps_of_this_stream": round(n / dec_p / 1e6, 2)}
        ps.free()
        if shape == "pw64k":
            hist = torch.zeros(256, dtype=torch.int32, device="cuda")
            counts = torch.zeros(((n + 65535) // 65536, 256), dtype=torch.int16, device="cuda")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for label, fn in (("observe_hist", lambda: pkg.observe_hist_device(d_in.data_ptr(), n, hist.data_ptr(), st)),
                              ("segment_hists_64k", lambda: pkg.make_hist_segments_device(d_in.data_ptr(), n, 65536, a.bits, counts.data_ptr(), st))):
                fn(); torch.cuda.synchronize()
                e0.record()
                for _ in range(5):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                res[label] = {"GBps": round(n / ms / 1e6, 1), "ms": round(ms, 3), "frac": round(n / ms / 1e6 / peak, 4)}
        del d_in, d_out, out3
    return res


def main():
    a = parse_args()
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        run_reference(a, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    pkg = entry.load_package()
    if pkg.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (" + pkg.last_error() + ")")
    torch.cuda.set_device(local_rank)
    pkg.lib().hsr_set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")  # host-side barriers that leave the GPUs alone

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        if world > 1:
            dist.barrier(group=cpu_group)

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    max_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.MAX)  # noqa: E731
    min_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.MIN)  # noqa: E731

    pkg.set_option("table", a.table)
    pkg.set_option("warps", a.ctas_per_sm)
    pkg.set_option("overlap", 0 if a.no_overlap else 1)

    # ---------------------------------------------------------------- the ONE stream: same bytes on every rank, encoded once
    _stage('synth')
    data = make_data(a)
    n = data.size
    _stage('reference encode')
    t0 = time.time()
    if rank == 0:
        stream = ref_encode(a, data)
    enc_s = time.time() - t0
    if world > 1:
        size_t = torch.tensor([stream.size if rank == 0 else 0], dtype=torch.int64, device="cuda")
        dist.broadcast(size_t, 0)
        buf = torch.from_numpy(stream).cuda() if rank == 0 else torch.empty(int(size_t.item()), dtype=torch.uint8, device="cuda")
        dist.broadcast(buf, 0)
        if rank != 0:
            stream = buf.cpu().numpy()
        del buf
    comp = stream.size

    # ---------------------------------------------------------------- kernel path: this rank's shard + index resident in HBM
    _stage('upload')
    ps = pkg.PreparedStream.upload(FAMILY_MT, a.states, a.bits, stream, shard=rank, shards=world)
    _stage('first decode + check')
    units = int(ps.units)
    my_off, my_bytes, my_in = int(ps.shard_out_offset), int(ps.shard_out_bytes), int(ps.shard_in_bytes)
    out_dev = torch.empty(my_bytes + 256, dtype=torch.uint8, device="cuda")
    cur = torch.cuda.current_stream().cuda_stream
    launches_per_step = ps.decode_async(out_dev.data_ptr(), my_bytes, cur, True)
    torch.cuda.synchronize()
    if ps.status() != 0 or not np.array_equal(out_dev[:my_bytes].cpu().numpy(), data[my_off: my_off + my_bytes]):
        raise SystemExit(f"rank {rank}: GPU output differs from the original bytes — refusing to report a number")
    # all n bytes, assembled on rank 0 over NCCL (untimed; not part of the decode roofline)
    assemble = None
    gathered = None
    if world > 1:
        meta = [None] * world
        dist.all_gather_object(meta, (my_off, my_bytes, units, my_in), group=cpu_group)
        plans = [pkg.ShardPlan(r, world, 0, 0, meta[r][0], meta[r][1], 0, 0) for r in range(world)]
        full = pkg.assemble_on(0, out_dev[:my_bytes], plans, n)   # warm-up (NCCL channel set-up) + the byte check
        if rank == 0 and not np.array_equal(full.cpu().numpy(), data):
            raise SystemExit("assembled output differs from the original bytes — refusing to report a number")
        del full
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        full = pkg.assemble_on(0, out_dev[:my_bytes], plans, n)
        a1.record()
        barrier()
        if rank == 0:
            ms = a0.elapsed_time(a1)
            assemble = {"ms": round(ms, 3), "GBps_into_rank0": round((n - my_bytes) / ms / 1e6, 1), "all_bytes_equal_original": True,
                        "transport": "torch.distributed NCCL send/recv, untimed"}
        del full
        gathered = meta

    _stage('warmup + timed steps')
    for _ in range(a.warmup):
        ps.decode_async(out_dev.data_ptr(), my_bytes, cur, True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for k in range(a.steps):
            ps.decode_async(out_dev.data_ptr(), my_bytes, cur, True)
        e1.record()
        barrier()
        # The timed region lasts ~20 ms, about one NVML query: keep the same launches going (untimed) under the same
        # sampler so the clock record covers a stretch of this exact load, not a single reading
        t_probe = time.time()
        while not (a.kernel_only or a.headline_only) and time.time() - t_probe < 0.25:  # not under the profiler
            for _ in range(10):
                ps.decode_async(out_dev.data_ptr(), my_bytes, cur, True)
            torch.cuda.synchronize()
    my_ms = e0.elapsed_time(e1) / a.steps
    ms_per_step = max_over_ranks(my_ms)
    ms_fastest_rank = min_over_ranks(my_ms)
    value = n / (ms_per_step * 1e-3) / 1e9
    clocks = clk.summary()
    clocks["window"] = "timed steps" if (a.kernel_only or a.headline_only) else "timed steps + 0.25 s of the same launches (untimed)"

    _stage('serialised steps')
    # the same K steps strictly serialised (one launch at a time, an event after each): per-launch durations
    pkg.set_option("overlap", 0)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record()
    for k in range(a.steps):
        ps.decode_async(out_dev.data_ptr(), my_bytes, cur, True)
        ev[k + 1].record()
    barrier()
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(a.steps)]
    serial_ms = max_over_ranks(ev[0].elapsed_time(ev[a.steps]) / a.steps)
    pkg.set_option("overlap", 0 if a.no_overlap else 1)

    if a.kernel_only:
        if rank == 0:
            _emit({"kernel_only": True, "value": round(value, 3), "unit": "GB/s", "ms_per_step": round(ms_per_step, 4),
                              "serialized_GBps": round(n / serial_ms / 1e6, 3), "serialized_ms": round(serial_ms, 4),
                              "bits": a.bits, "states": a.states, "table": a.table, "ctas_per_sm": a.ctas_per_sm, "blocks": units, "compressed": comp,
                              "n_gpus": world, "traffic_GBps": round((comp + n) / (ms_per_step * 1e-3) / 1e9, 1), "clocks": clocks})
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- end to end through the drop-in host call
    _stage('e2e')
    e2e_steps = a.e2e_steps or min(a.steps, 10)
    lib = pkg.lib()
    e2e = None
    e2e_shards = None
    clk2_summary = None
    if rank == 0:
        hin, hout = pkg.host_alloc(comp), pkg.host_alloc(n)
        hin.array[:] = stream

        def e2e_step():
            if world == 1:
                got = lib.hsr_decode(FAMILY_MT, a.states, a.bits, hin.ptr, comp, hout.ptr, n)
            else:
                got = lib.hsr_decode_mt_multi(a.states, a.bits, hin.ptr, comp, hout.ptr, n, None, world)
            if got != n:
                raise SystemExit(f"end-to-end decode failed: {pkg.last_error()}")

        for _ in range(max(1, min(a.warmup, 3))):
            e2e_step()
        if not np.array_equal(hout.array, data):
            raise SystemExit("end-to-end output differs from the original bytes")
        hout.array[:] = 0xCC
    host_barrier()
    if rank == 0:
        with ClockSampler(local_rank) as clk2:
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            e2e_s = time.perf_counter() - t0
        clk2_summary = clk2.summary()
        if not np.array_equal(hout.array, data):
            raise SystemExit("end-to-end output differs from the original bytes")
        e2e = {"value": round(n * e2e_steps / e2e_s / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": comp, "d2h_bytes_per_step": n,
               "steps": e2e_steps, "ms_per_step": round(e2e_s / e2e_steps * 1e3, 3),
               "api": "hsr_decode (host pointers, pinned)" if world == 1 else
                      f"hsr_decode_mt_multi over {world} devices from rank 0 (host pointers, pinned; the other ranks wait on a host barrier)"}
        hin.free(); hout.free()
    host_barrier()
    if world > 1:
        # the same bytes the other way the north star words it: one process per GPU, every rank decoding ITS shard from
        # its own pinned copy of the stream into its own pinned output buffer, all ranks at once
        hin, hout = pkg.host_alloc(comp), pkg.host_alloc(n)
        hin.array[:] = stream
        import ctypes
        off = ctypes.c_size_t(0)

        def shard_step():
            got = lib.hsr_decode_mt_shard(a.states, a.bits, hin.ptr, comp, hout.ptr, n, rank, world, ctypes.byref(off))
            if got != my_bytes:
                raise SystemExit(f"rank {rank}: hsr_decode_mt_shard failed: {pkg.last_error()}")

        shard_step()
        if off.value != my_off or not np.array_equal(hout.array[my_off: my_off + my_bytes], data[my_off: my_off + my_bytes]):
            raise SystemExit(f"rank {rank}: shard end-to-end output differs from the original bytes")
        host_barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            shard_step()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e_shards = {"value": round(n * e2e_steps / dt / 1e9, 3), "unit": "GB/s", "ms_per_step": round(dt / e2e_steps * 1e3, 3),
                      "api": "hsr_decode_mt_shard, one process per GPU, all ranks at once (each rank walks the whole chain, copies only its shard)"}
        hin.free(); hout.free()

    # ---------------------------------------------------------------- optional: round 1's weak-scaling figure
    weak = None
    if world > 1 and a.weak:
        wdata = make_data(a, seed=42 + rank)
        wstream = ref_encode(a, wdata)
        wps = pkg.PreparedStream.upload(FAMILY_MT, a.states, a.bits, wstream)
        wout = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
        wps.decode_async(wout.data_ptr(), n, cur)
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(a.steps):
            wps.decode_async(wout.data_ptr(), n, cur)
        w1.record()
        barrier()
        wms = max_over_ranks(w0.elapsed_time(w1) / a.steps)
        weak = {"value": round(world * n / wms / 1e6, 2), "unit": "GB/s", "note": "N independent 1 GB streams, one per rank (weak scaling)"}
        wps.free()
        del wout

    # ---------------------------------------------------------------- block_ at N > 1: replicas only
    # A block_ (or raw) stream is one recurrence and does not shard; what spreads over GPUs is a BATCH of streams, every
    # rank decoding its own replica of the batch with no exchange at all. Reported so the N-GPU line carries the block_
    # variant too: aggregate = N x the slowest rank's rate.
    block_replicas = None
    if world > 1 and not a.headline_only and not a.kernel_only:
        import checkers as ck
        _stage('block_ batch replicas')
        k_streams, each = 2368, 400_000     # 16 streams per SM, the batch of `other_configs` at N = 1
        bdata = make_data(a, seed=900 + rank, size=k_streams * each, shape="iid")
        parts, items, pos = [], [], 0
        for k in range(k_streams):
            bs = ck.ref_encode(FAMILY_BLOCK, 32, 10, bdata[k * each:(k + 1) * each])
            pad = (-pos) % 16
            parts.append(np.zeros(pad, np.uint8)); pos += pad
            items.append((pos, bs.size, k * each, each))
            parts.append(bs); pos += bs.size
        bps = pkg.PreparedStream.upload_batch(FAMILY_BLOCK, 32, 10, np.concatenate(parts), items)
        btotal = k_streams * each
        bout = torch.empty(btotal + 64, dtype=torch.uint8, device="cuda")
        bps.decode_async(bout.data_ptr(), btotal, cur)
        torch.cuda.synchronize()
        b_ok = bps.status() == 0 and bool(np.array_equal(bout[:btotal].cpu().numpy(), bdata))
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(5):
            bps.decode_async(bout.data_ptr(), btotal, cur)
        b1.record()
        barrier()
        bms = max_over_ranks(b0.elapsed_time(b1) / 5)
        all_ok = min_over_ranks(1.0 if b_ok else 0.0) > 0.5
        block_replicas = {"value": round(world * btotal / bms / 1e6, 2), "unit": "GB/s", "codec": "block_rANS32x32_16w 10-bit",
                          "streams_per_gpu": k_streams, "bytes_per_stream": each, "ms_per_step": round(bms, 4), "bit_exact_on_every_rank": all_ok,
                          "note": "replicas only: every rank decodes its own batch of independent block_ streams, no exchange"}
        bps.free()
        del bout

    if rank == 0:
        # ---------------------------------------------------------------- index timings (reported apart, SURVEY §8d "I")
        _stage('index timings')
        index_host_ms = ps.index_ms
        dev_in = torch.from_numpy(stream).cuda()
        ds = pkg.PreparedStream.from_device(FAMILY_MT, a.states, a.bits, dev_in.data_ptr(), comp)
        index_device_ms = ds.index_ms
        ds.free()
        del dev_in

        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
        # roofline of the dominant kernel, per GPU: this rank's shard bytes over this rank's mean launch duration
        achieved = (my_in + my_bytes) / (my_ms * 1e-3) / 1e9
        achieved_serial = (my_in + my_bytes) / (float(np.mean(step_ms)) * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")  # written from the committed `ncu --set full` capture of this command
        if world == 1 and os.path.exists(tpath) and a.size == 1_000_000_000 and a.bits == 15 and a.states == 64 and a.shape == "pw64k":
            try:
                traffic = json.load(open(tpath)).get("bytes_per_launch")
            except Exception:
                traffic = None
        table_kind = 2 if (a.bits <= 11 and a.table != 1) or a.table == 2 else 1
        line = {
            "metric": "decoded_GBps", "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(a), "codec": f"mt_rANS32x{a.states}_16w", "bits": a.bits, "shape": a.shape,
                       "decoded_bytes": n, "compressed_bytes": comp,
                       "blocks_per_gpu": [m[2] for m in gathered] if gathered else [units],
                       "decoded_bytes_per_gpu": [m[1] for m in gathered] if gathered else [my_bytes],
                       "stream_producer": "reference mt_ encoder (oracle/_ref), unmodified; same stream on every rank",
                       "l2_policy": f"inputs larger than L2: {(comp + n) / world / 1e9:.2f} GB touched per GPU and step vs 126 MB L2",
                       "parallelism": f"1 stream, {world} contiguous block ranges (hsr_stream_upload shard/shards), no collective",
                       "step_overlap": "off (--no-overlap)" if a.no_overlap else
                                       "consecutive launches overlap (programmatic dependent launch); serialised figure in value_serialized",
                       "table": "auto (bitmap-rank for bits>=12, packed slot table below)"},
            "value_serialized": {"value": round(n / serial_ms / 1e6, 3), "ms_per_step": round(serial_ms, 4),
                                 "note": "same K steps, one launch at a time (overlap=0), max over ranks"},
            "slowest_vs_fastest_rank_ms": [round(ms_per_step, 4), round(ms_fastest_rank, 4)],
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "peak_source": peak_src, "kernel": f"units_n{a.states}_b{a.bits}_t{table_kind}",
                         "algorithmic_bytes_per_launch": my_in + my_bytes, "kernel_ms": round(my_ms, 4),
                         "frac_serialized": round(achieved_serial / peak, 4), "kernel_ms_serialized": round(float(np.mean(step_ms)), 4),
                         "kernel_ms_serialized_min": round(float(np.min(step_ms)), 4), "scope": "rank 0's GPU and shard"},
            "e2e": e2e,
            "gpu_launches": int(launches_per_step * a.steps),
            "clocks": clocks, "clocks_e2e": clk2_summary,
            "assemble_on_rank0": assemble,
            "index_ms": {"host_walk": round(index_host_ms, 3), "device_walk": round(index_device_ms, 3)},
            "setup_s": {"reference_encode": round(enc_s, 2)},
        }
        if e2e_shards:
            line["e2e_one_process_per_gpu"] = e2e_shards
        if weak:
            line["weak_scaling"] = weak
        if block_replicas:
            line["block_replicas"] = block_replicas
        _stage('cpu baseline')
        if not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a, stream, n)
        _stage('other configs')
        if not a.headline_only:
            line["other_configs"] = other_configs(pkg, torch, a, peak, a.extra)
        _emit(line)
    ps.free()
    if world > 1:
        host_barrier()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
