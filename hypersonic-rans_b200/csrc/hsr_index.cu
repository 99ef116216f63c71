// hsr_index.cu — parallel construction of the mt_ block index for streams that live in device memory.
//
// The reference finds the blocks of an mt_ stream by walking a linked list on one thread: each header stores the
// distance to the next (src/mt_rANS32x64_16w_decode.cpp:57-59,94). On the GPU that walk is one dependent DRAM
// round trip per block (~0.87 us, 13 ms per GB). Here the stream is cut into K segments and K warps work at once:
//
//   find   warp k scans ITS segment for the first position that looks like a coded block header: the 256 u16 counts
//          at +16+4N must sum to 2^bits (a sliding-window sum, two coalesced loads per 32 candidate positions) and
//          the size / skip fields must be plausible. found[0] is the true chain start (byte 16).
//   walk   warp k follows the chain from found[k] until it lands exactly on some found[j], j > k, and hands over.
//          Pass 0 only counts units and decoded bytes; a single thread then strings the hand-overs together from
//          warp 0 (whose start is genuine, so everything it reaches is genuine; false positives are never reached)
//          and prefix-sums the counts; pass 1 repeats the walk writing hsr_block_t records at their final places
//          with the serial walk's exact arithmetic (src/mt_rANS32x64_16w_decode.cpp:43-94).
//
// The signature is only a hint — correctness rests on the hand-over rule — so any surprise (hop limit, count
// mismatch, malformed field) makes the caller fall back to the serial walk, which also reports the real error.
#include <algorithm>
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/hsrans_b200.h"
#include "hsr_device.cuh"

namespace hsr {

constexpr uint64_t kNone = ~0ull;
constexpr uint64_t kFillUnitIdx = 4ull << 20; // must match hsr_api.cu: fills are cut into pieces of this size
constexpr uint32_t kMaxHops = 1u << 14;

struct IndexPlan {
  const uint8_t *in;
  uint64_t compLen, n;
  uint32_t N, bits, K;
  uint64_t segBytes; // even
};

struct WalkerState {
  uint64_t found;     // byte position of the first header candidate in the segment, or kNone
  uint64_t units;     // pass 0: units emitted by this walker
  uint64_t outBytes;  // pass 0: decoded bytes covered (unclamped size fields)
  uint64_t unitBase;  // resolve: index of this walker's first unit
  uint64_t outBase;   // resolve: decoded offset of this walker's first unit
  uint32_t handTo;    // walker that continues the chain, or 0xffffffff at the end of the stream
  uint32_t flags;     // bit0 reachable, bit1 error/overflow in pass 0
};

__device__ __forceinline__ uint64_t ld_u64_by_lanes(const uint8_t *p, uint32_t lane, uint32_t firstLane)
{
  // lanes firstLane..firstLane+3 each fetch one u16; every lane gets the assembled u64
  uint32_t h = 0;
  if (lane >= firstLane && lane < firstLane + 4) h = ldg_u16(p + 2 * (lane - firstLane));
  const uint64_t a = __shfl_sync(kFull, h, firstLane), b = __shfl_sync(kFull, h, firstLane + 1);
  const uint64_t c = __shfl_sync(kFull, h, firstLane + 2), d = __shfl_sync(kFull, h, firstLane + 3);
  return a | (b << 16) | (c << 32) | (d << 48);
}

// ---------------------------------------------------------------------------------------------- find

__global__ void __launch_bounds__(128) index_find_kernel(IndexPlan pl, WalkerState *ws)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= pl.K) return;
  const uint64_t headerBytes = 16 + 4ull * pl.N + 512;
  uint64_t found = kNone;
  if (k == 0) {
    found = 16;
  } else {
    const uint64_t begin = 16 + k * pl.segBytes;
    uint64_t end = begin + pl.segBytes; // candidates p in [begin, end)
    if (pl.compLen >= headerBytes && end > pl.compLen - headerBytes + 2) end = pl.compLen - headerBytes + 2;
    if (begin < end && pl.compLen >= headerBytes) {
      const uint8_t *in = pl.in;
      const uint64_t cOff = 16 + 4ull * pl.N; // counts relative to the header start
      // S0 = sum of the 256 counts of the candidate at `begin`
      uint32_t part = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) part += ldg_u16(in + begin + cOff + 2 * (lane * 8 + i));
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(kFull, part, d);
      uint32_t s0 = part;
      const uint32_t want = 1u << pl.bits;
      // The window sum is a loop-carried value but the loads are not: fetch kBatch steps' worth of words first so
      // their DRAM latency overlaps, then slide through them.
      constexpr int kBatch = 8;
      for (uint64_t base0 = begin; base0 < end && found == kNone; base0 += 64ull * kBatch) {
        uint32_t leave[kBatch], enter[kBatch];
#pragma unroll
        for (int t = 0; t < kBatch; t++) {
          // lane l tests position base + 2l; moving one position right drops word [q] and gains word [q + 512]
          const uint64_t q = base0 + 64ull * t + cOff + 2 * lane;
          leave[t] = q + 2 <= pl.compLen ? ldg_u16(in + q) : 0u;
          enter[t] = q + 514 <= pl.compLen ? ldg_u16(in + q + 512) : 0u;
        }
#pragma unroll
        for (int t = 0; t < kBatch; t++) {
          const uint64_t base = base0 + 64ull * t;
          const uint32_t delta = enter[t] - leave[t];
          uint32_t incl = delta;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(kFull, incl, d);
            if (lane >= (uint32_t)d) incl += up;
          }
          const uint32_t mine = s0 + incl - delta; // sum of the counts for my candidate
          const uint64_t p = base + 2 * lane;
          const bool hit = found == kNone && p < end && mine == want;
          uint32_t m = __ballot_sync(kFull, hit);
          while (m && found == kNone) { // verify candidates in order (rare)
            const uint32_t l = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t cp = base + 2 * l;
            const uint64_t v = ld_u64_by_lanes(in + cp, lane, 0);
            const uint64_t skip = ld_u64_by_lanes(in + cp + 8, lane, 4);
            const bool sizeOk = !(v >> 63) && v != 0 && v <= pl.n;
            const bool skipOk = skip < (pl.compLen - (cp + 16)) / 2 && cp + 16 + 2 * (skip + 1) + 2 >= cp + headerBytes;
            // the states of a real header all sit in [2^15, 2^31)
            const uint32_t st = ldg_u32_a2(in + cp + 16 + 4 * lane);
            const bool statesOk = __all_sync(kFull, st >= kConsumePoint16 && st < 0x80000000u);
            if (sizeOk && skipOk && statesOk) found = cp;
          }
          s0 += __shfl_sync(kFull, incl, 31);
        }
      }
    }
  }
  if (lane == 0) {
    WalkerState w{};
    w.found = found;
    w.handTo = 0xffffffffu;
    ws[k] = w;
  }
}

// ---------------------------------------------------------------------------------------------- walk

// pass 0: count; pass 1: write records. One warp per walker.
template <int PASS>
__global__ void __launch_bounds__(128) index_walk_kernel(IndexPlan pl, WalkerState *ws, hsr_block_t *blocks, uint64_t maxBlocks,
                                                         unsigned long long *result /* [0] units, [1] error */)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= pl.K) return;
  WalkerState w = ws[k];
  if (w.found == kNone) return;
  if (PASS == 1 && !(w.flags & 1u)) return; // not on the chain
  const uint8_t *in = pl.in;
  const uint64_t N = pl.N, n = pl.n;
  const uint64_t outLengthInStates = n - N + 1;
  uint64_t pos = w.found, units = 0, outBytes = 0, i = PASS == 1 ? w.outBase : 0;
  uint64_t lastCodedUnit = kNone;
  uint32_t j = k + 1, hops = 0, err = 0, handTo = 0xffffffffu;
  bool done = false;
  while (!done) {
    if (pos + 8 > pl.compLen) break; // end of the stream
    if (PASS == 1 && !(i < outLengthInStates)) break; // src/mt_rANS32x64_16w_decode.cpp:97
    if (++hops > kMaxHops) { err = 1; break; }
    uint32_t hw = 0; // lanes 0..7 fetch {size, skip} in one round trip
    if (lane < 8 && pos + 2 * lane + 2 <= pl.compLen) hw = ldg_u16(in + pos + 2 * lane);
    uint64_t v = 0, skip = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      v |= (uint64_t)__shfl_sync(kFull, hw, t) << (16 * t);
      skip |= (uint64_t)__shfl_sync(kFull, hw, t + 4) << (16 * t);
    }
    if (v >> 63) { // single-symbol run (:46-54)
      const uint64_t size = v & ((1ull << 54) - 1);
      if (PASS == 1 && size > n - i) { err = 3; break; }
      const uint64_t pieces = size ? (size + kFillUnitIdx - 1) / kFillUnitIdx : 0;
      if (PASS == 1) {
        for (uint64_t o = 0, c = 0; c < pieces; c++, o += kFillUnitIdx) {
          const uint64_t idx = w.unitBase + units + c;
          if (lane == 0 && idx < maxBlocks) {
            hsr_block_t b{};
            b.inOffset = pos; b.inEnd = pos + 8; b.outOffset = i + o; b.count = size - o < kFillUnitIdx ? size - o : kFillUnitIdx;
            b.kind = 1; b.symbol = (uint32_t)(v >> 54) & 0xffu;
            blocks[idx] = b;
          }
        }
      }
      units += pieces;
      outBytes += size;
      i += size;
      pos += 8;
    } else {
      if (pos + 16 + 4 * N + 512 > pl.compLen) { err = 4; break; }
      if (skip >= (pl.compLen - (pos + 16)) / 2) { err = 5; break; }
      uint64_t after = pos + 16 + 2 * (skip + 1); // :59
      uint64_t count = v;
      if (PASS == 1) {
        uint64_t end = i + v; // :77-82
        if (end > outLengthInStates) end = outLengthInStates;
        else if (end & (N - 1)) { err = 7; break; }
        const uint64_t rows = end > i ? (end - i + N - 1) / N : 0;
        count = rows * N;
        if (!(i + count < outLengthInStates)) after = pl.compLen; // last block: see hsr_mt_index
      }
      if (after < pos + 16 + 4 * N + 512 - (PASS == 0 ? 2 : 0) || after > pl.compLen) { err = 6; break; }
      if (PASS == 1) {
        const uint64_t idx = w.unitBase + units;
        if (lane == 0 && idx < maxBlocks) {
          hsr_block_t b{};
          b.inOffset = pos + 16; b.inEnd = after; b.outOffset = i; b.count = count; b.kind = 0;
          blocks[idx] = b;
        }
        lastCodedUnit = w.unitBase + units;
      }
      units += 1;
      outBytes += v;
      i += count;
      pos = after;
    }
    while (j < pl.K && ws[j].found < pos) j++;  // kNone sorts last, so exhausted segments are skipped as well
    if (j < pl.K && ws[j].found == pos) { handTo = j; done = true; }
  }
  if (PASS == 0) {
    if (lane == 0) {
      ws[k].units = units;
      ws[k].outBytes = outBytes;
      ws[k].handTo = handTo;
      ws[k].flags = err ? 2u : 0u;
    }
  } else {
    if (lane == 0) {
      if (err || units != w.units) atomicExch(result + 1, 100ull + err);
      if (handTo == 0xffffffffu) { // the walker that finished the chain: ragged tail (:99-130)
        if (i < n) {
          if (lastCodedUnit == kNone || lastCodedUnit + 1 != w.unitBase + units) atomicExch(result + 1, 108ull);
          else if (lastCodedUnit < maxBlocks) {
            blocks[lastCodedUnit].tail = (uint32_t)(n - i);
            blocks[lastCodedUnit].count += n - i;
          }
        }
      }
    }
  }
}

// Strings the hand-overs together, starting from walker 0. The chain of walkers is itself a linked list, so it is
// copied into shared memory first: the serial part then costs ~30 cycles per walker instead of a DRAM round trip.
constexpr uint32_t kMaxWalkers = 3064;

__global__ void __launch_bounds__(1024) index_resolve_kernel(IndexPlan pl, WalkerState *ws, unsigned long long *result)
{
  __shared__ uint64_t sOut[kMaxWalkers];   // in: decoded bytes of the walker; out: its decoded base offset
  __shared__ uint32_t sUnits[kMaxWalkers]; // in: units of the walker;        out: its first unit index (fits: < 2^32 units)
  __shared__ uint32_t sHand[kMaxWalkers];  // in: next walker;                out: 1 if on the chain
  for (uint32_t k = threadIdx.x; k < pl.K; k += blockDim.x) {
    const WalkerState w = ws[k];
    sOut[k] = w.outBytes;
    sUnits[k] = (uint32_t)w.units;
    sHand[k] = (w.found == kNone || (w.flags & 2u) || w.units > 0xffffffffull) ? 0xfffffffeu : w.handTo;
  }
  __syncthreads();
  __shared__ uint32_t sFail;
  if (threadIdx.x == 0) {
    uint64_t unitBase = 0, outBase = 0;
    uint32_t k = 0, steps = 0, fail = 0;
    for (;;) {
      const uint32_t hand = sHand[k];
      if (hand == 0xfffffffeu || ++steps > pl.K || unitBase > 0xffffffffull) { fail = 200; break; }
      const uint64_t out = sOut[k];
      const uint32_t units = sUnits[k];
      sOut[k] = outBase;
      sUnits[k] = (uint32_t)unitBase;
      sHand[k] = 0xfffffffdu; // on the chain
      unitBase += units;
      outBase += out;
      if (hand == 0xffffffffu) break;
      if (hand <= k) { fail = 201; break; }
      k = hand;
    }
    sFail = fail;
    if (fail) result[1] = fail;
    else result[0] = unitBase;
  }
  __syncthreads();
  if (sFail) return;
  for (uint32_t k = threadIdx.x; k < pl.K; k += blockDim.x) {
    if (sHand[k] == 0xfffffffdu) {
      ws[k].flags |= 1u;
      ws[k].unitBase = sUnits[k];
      ws[k].outBase = sOut[k];
    }
  }
}

} // namespace hsr

using namespace hsr;

// Builds the index of the mt_ stream at dIn (device memory). Returns false when the caller must fall back to the
// serial walk (nothing is reported as an error here). *ms receives the GPU time of the four kernels.
bool hsr_parallel_mt_index(const uint8_t *dIn, uint64_t compLen, uint64_t n, int N, int bits, std::vector<hsr_block_t> *out, float *ms)
{
  const uint64_t headerBytes = 16 + 4ull * N + 512;
  if (compLen < 16 + headerBytes || n < (uint64_t)N) return false;
  IndexPlan pl{};
  pl.in = dIn; pl.compLen = compLen; pl.n = n; pl.N = (uint32_t)N; pl.bits = (uint32_t)bits;
  // about four 64 KiB blocks per segment, at most kMaxWalkers walkers
  uint64_t K = std::min<uint64_t>(kMaxWalkers, std::max<uint64_t>(1, compLen / (192 * 1024)));
  pl.segBytes = ((compLen - 16 + K - 1) / K + 63) & ~63ull;
  K = (compLen - 16 + pl.segBytes - 1) / pl.segBytes;
  pl.K = (uint32_t)K;

  WalkerState *dWs = nullptr;
  unsigned long long *dRes = nullptr;
  hsr_block_t *dBlocks = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool ok = false;
  uint64_t cap = std::max<uint64_t>(1024, n / 16384 + 64);
  do {
    if (cudaMalloc(&dWs, K * sizeof(WalkerState)) != cudaSuccess || cudaMalloc(&dRes, 16) != cudaSuccess ||
        cudaMalloc(&dBlocks, cap * sizeof(hsr_block_t)) != cudaSuccess)
      break;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemset(dRes, 0, 16);
    const unsigned grid = (unsigned)((K + 3) / 4);
    cudaEventRecord(e0);
    index_find_kernel<<<grid, 128>>>(pl, dWs);
    index_walk_kernel<0><<<grid, 128>>>(pl, dWs, nullptr, 0, dRes);
    index_resolve_kernel<<<1, 1024>>>(pl, dWs, dRes);
    unsigned long long res[2] = {0, 0};
    if (cudaMemcpy(res, dRes, 16, cudaMemcpyDeviceToHost) != cudaSuccess || res[1] || res[0] == 0) break;
    if (res[0] > cap) {
      cudaFree(dBlocks); dBlocks = nullptr;
      cap = res[0];
      if (cudaMalloc(&dBlocks, cap * sizeof(hsr_block_t)) != cudaSuccess) break;
    }
    index_walk_kernel<1><<<grid, 128>>>(pl, dWs, dBlocks, cap, dRes);
    cudaEventRecord(e1);
    unsigned long long res2[2] = {0, 0};
    if (cudaMemcpy(res2, dRes, 16, cudaMemcpyDeviceToHost) != cudaSuccess || res2[1]) break;
    out->resize((size_t)res[0]);
    if (cudaMemcpy(out->data(), dBlocks, out->size() * sizeof(hsr_block_t), cudaMemcpyDeviceToHost) != cudaSuccess) break;
    // the chain must cover the decoded length exactly
    const hsr_block_t &last = out->back();
    if (last.outOffset + last.count != n) break;
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    if (ms) *ms += t;
    ok = true;
  } while (false);
  (void)cudaGetLastError();
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(dWs); cudaFree(dRes); cudaFree(dBlocks);
  if (!ok) out->clear();
  return ok;
}
