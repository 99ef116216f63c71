// hsr_encode.cu — device-side producer of mt_rANS32xN_16w streams.
//
// Keeps the reference's stream format byte for byte (src/mt_rANS32x64_16w_encode.cpp:266-298,363-379: per block
// {u64 symbol count | u64 skip | u32 states[N] | u16 counts[256] | u16 words[]}) and its per-block histogram
// (observe_hist + normalize_hist over exactly the block's bytes, :207-210), but replaces the block policy constant:
// blocks have a FIXED size (default 64 KiB = the reference's MinBlockSize, :34-48) and every block is encoded from
// fresh states, so blocks are independent work for the whole GPU. Any reference decoder decodes the result
// (mt_rANS32xNN_16w_decode_<b> and the thread-pool variant); for inputs of at most one block the stream is
// byte-identical to the reference encoder's output.
//
// Pipeline (all on one CUDA stream):
//   enc_hist_kernel      one CTA per block: shared-memory-atomic byte counts + the reference's normalize_hist
//   enc_block_kernel     one warp per block, one rANS state per lane (two for N = 64): walks the block BACKWARDS
//                        (src/block_codec64.h:55-100), renormalisation words handed out by ballot/popc in reverse,
//                        written downward into a per-block scratch slot; x / freq by exact reciprocal multiply
//   enc_scan_kernel      exclusive scan of the block sizes -> stream offsets, total compressed length
//   enc_assemble_kernel  one CTA per block: writes the block header and moves its words to their final place
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/hsrans_b200.h"
#include "hsr_device.cuh"
#include "hsr_hist_device.cuh"

namespace hsr {

struct EncPlan {
  uint64_t n;          // input bytes
  uint32_t blockSize;  // symbols per block (multiple of N)
  uint32_t numBlocks;
  uint64_t lastStart;  // first symbol of the last block (it also owns the < N ragged symbols at the end)
  uint64_t slotBytes;  // scratch bytes per block: worst case one word per symbol
};

struct EncBlockMeta {
  uint32_t wordBytes;
  uint32_t pad;
  uint32_t states[64];
};

__device__ __forceinline__ uint64_t block_begin(const EncPlan &pl, uint32_t k) { return (uint64_t)k * pl.blockSize; }
__device__ __forceinline__ uint64_t block_end(const EncPlan &pl, uint32_t k) { return k + 1 == pl.numBlocks ? pl.n : (uint64_t)(k + 1) * pl.blockSize; }

// ---------------------------------------------------------------------------------------------- per-block histograms

constexpr int kSegWarps = 4; // warps per CTA in the per-block histogram kernels

__global__ void __launch_bounds__(kSegWarps * 32) enc_hist_kernel(const uint8_t *data, EncPlan pl, int bits, uint16_t *counts)
{
  __shared__ uint32_t sHist[kSegWarps][256];
  __shared__ uint16_t sCapped[kSegWarps][256];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t k = blockIdx.x * kSegWarps + warp; k < pl.numBlocks; k += gridDim.x * kSegWarps) {
    const uint64_t begin = block_begin(pl, k), end = block_end(pl, k);
    warp_observe(data, begin, end, sHist[warp], lane);
    warp_normalize(sHist[warp], end - begin, bits, sCapped[warp], sHist[warp], counts + (uint64_t)k * 256, lane); // :209-210
  }
}

// ---------------------------------------------------------------------------------------------- block encoder

// per-symbol encoder entry: x' = x + bias + mulhi(x, rcp) >> shift * cmpl   ==   ((x / f) << b) + cumul + x % f
// (exact for x < 2^31; the construction of F. Giesen's rans_byte.h RansEncSymbolInit, restated)
struct EncSym {
  uint32_t rcp;       // ceil(2^(shift + 31) / freq), or 0xffffffff for freq == 1
  uint32_t bias;      // cumul (+ 2^b - 1 for freq == 1)
  uint32_t cmplShift; // (2^b - freq) | shift << 16
  uint32_t xmax;      // emit a word while x >= ((2^15 >> b) << 16) * freq   (src/rANS32x32_16w.cpp:41,66)
};

template <int BITS, int N>
__global__ void __launch_bounds__(32, 32) enc_block_kernel(const uint8_t *__restrict__ in, EncPlan pl, const uint16_t *counts, uint8_t *scratch,
                                                           EncBlockMeta *meta, uint32_t *counter)
{
  __shared__ __align__(16) EncSym sSym[256];
  const uint32_t lane = lane_id();
  const uint32_t ltMask = lanemask_lt();
  const uint32_t lanePos = idx2idx_lane(lane);
  constexpr uint32_t kEmit = (kConsumePoint16 >> BITS) << 16;

  for (;;) {
    uint32_t k = 0;
    if (lane == 0)
      k = atomicAdd(counter, 1u);
    k = __shfl_sync(kFull, k, 0);
    if (k >= pl.numBlocks)
      break;

    // table: lane owns symbols 8l .. 8l+7
    {
      const uint16_t *c = counts + (uint64_t)k * 256 + lane * 8;
      uint32_t f[8], sum = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) { f[i] = c[i]; sum += f[i]; }
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(kFull, incl, d);
        if (lane >= (uint32_t)d) incl += up;
      }
      uint32_t cumul = incl - sum;
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; i++) {
        EncSym e;
        const uint32_t freq = f[i];
        if (freq < 2) {
          e.rcp = 0xffffffffu;
          e.bias = cumul + (1u << BITS) - 1u;
          e.cmplShift = ((1u << BITS) - freq);
        } else {
          uint32_t shift = 0;
          while (freq > (1u << shift)) shift++;
          e.rcp = (uint32_t)((((uint64_t)1 << (shift + 31)) + freq - 1) / freq);
          e.bias = cumul;
          e.cmplShift = ((1u << BITS) - freq) | ((shift - 1) << 16);
        }
        e.xmax = kEmit * freq;
        sSym[lane * 8 + i] = e;
        cumul += freq;
      }
      __syncwarp();
    }

    const uint64_t begin = block_begin(pl, k), end = block_end(pl, k);
    const uint64_t len = end - begin;
    const uint32_t rows = (uint32_t)(len / N);
    const uint32_t tailLen = (uint32_t)(len % N);
    const uint8_t *src = in + begin;

    uint8_t *slotEnd = scratch + (uint64_t)(k + 1) * pl.slotBytes; // words grow downward from here
    uint8_t *wp = slotEnd;
    uint32_t x0 = kConsumePoint16, x1 = kConsumePoint16; // fresh states (:222-223)

    // one half-row, lanes = states 32h .. 32h+31; emission order is state N-1 .. 0, i.e. ascending addresses = ascending state
    auto step = [&](uint32_t &x, uint32_t sym, bool active) {
      const EncSym e = sSym[sym];
      const bool emit = active && x >= e.xmax;
      const uint32_t m = __ballot_sync(kFull, emit);
      const uint32_t total = __popc(m);
      uint8_t *base = wp - 2u * total;
      if (emit) {
        *reinterpret_cast<uint16_t *>(base + 2u * __popc(m & ltMask)) = (uint16_t)x;
        x >>= 16;
      }
      wp = base;
      if (active) {
        const uint32_t q = __umulhi(x, e.rcp) >> (e.cmplShift >> 16);
        x = x + e.bias + q * (e.cmplShift & 0xffffu);
      }
    };

    if (tailLen) { // ragged last row first (:227-252)
      const uint8_t *row = src + (uint64_t)rows * N;
      if constexpr (N == 64) {
        const bool a1 = lanePos + 32u < tailLen;
        step(x1, a1 ? row[lanePos + 32u] : 0u, a1);
      }
      const bool a0 = lanePos < tailLen;
      step(x0, a0 ? row[lanePos] : 0u, a0);
    }
    // rows are independent of the states: keep the next kAhead rows' symbols in flight while the current one encodes
    constexpr int kAhead = 4;
    uint32_t q0[kAhead], q1[kAhead];
#pragma unroll
    for (int a = 0; a < kAhead; a++) {
      const int64_t r = (int64_t)rows - 1 - a;
      const uint8_t *row = src + (uint64_t)(r < 0 ? 0 : r) * N;
      q0[a] = __ldg(row + lanePos);
      q1[a] = N == 64 ? __ldg(row + lanePos + 32u) : 0u;
    }
    for (int64_t r = (int64_t)rows - 1; r >= 0; r -= kAhead) {
#pragma unroll
      for (int a = 0; a < kAhead; a++) {
        if (r - a < 0) break;
        const uint32_t s0 = q0[a], s1 = q1[a];
        const int64_t rn = r - a - kAhead; // refill this slot with the row kAhead further down
        const uint8_t *row = src + (uint64_t)(rn < 0 ? 0 : rn) * N;
        q0[a] = __ldg(row + lanePos);
        if constexpr (N == 64) q1[a] = __ldg(row + lanePos + 32u);
        if constexpr (N == 64) step(x1, s1, true);
        step(x0, s0, true);
      }
    }

    EncBlockMeta *mt = meta + k;
    mt->states[lane] = x0;
    if constexpr (N == 64)
      mt->states[lane + 32] = x1;
    if (lane == 0)
      mt->wordBytes = (uint32_t)(slotEnd - wp);
  }
}

// ---------------------------------------------------------------------------------------------- offsets

// offsets[k] = stream offset of block k's u64 size field; offsets[numBlocks] = total compressed length
__global__ void __launch_bounds__(1024) enc_scan_kernel(const EncBlockMeta *meta, uint32_t numBlocks, uint32_t headerBytes, uint64_t *offsets)
{
  __shared__ uint64_t sWarp[32];
  __shared__ uint64_t sCarry;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) sCarry = 16; // u64 n, u64 compressed length
  __syncthreads();
  for (uint32_t base = 0; base < numBlocks; base += 1024) {
    const uint32_t k = base + tid;
    const uint64_t v = k < numBlocks ? (uint64_t)headerBytes + meta[k].wordBytes : 0;
    uint64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint64_t up = __shfl_up_sync(kFull, incl, d);
      if (lane >= (uint32_t)d) incl += up;
    }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint64_t w = sWarp[lane], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint64_t up = __shfl_up_sync(kFull, wi, d);
        if (lane >= (uint32_t)d) wi += up;
      }
      sWarp[lane] = wi - w; // exclusive prefix of the warps
    }
    __syncthreads();
    const uint64_t carry = sCarry;
    if (k < numBlocks) offsets[k] = carry + sWarp[warp] + incl - v;
    __syncthreads();
    if (tid == 1023) sCarry = carry + sWarp[warp] + incl;
    __syncthreads();
  }
  if (tid == 0) offsets[numBlocks] = sCarry;
}

// ---------------------------------------------------------------------------------------------- assembly

__device__ __forceinline__ void st_u16(uint8_t *p, uint32_t v) { *reinterpret_cast<uint16_t *>(p) = (uint16_t)v; }

template <int N>
__global__ void __launch_bounds__(256) enc_assemble_kernel(EncPlan pl, const uint16_t *counts, const uint8_t *scratch, const EncBlockMeta *meta,
                                                           const uint64_t *offsets, uint8_t *out)
{
  const uint32_t tid = threadIdx.x;
  constexpr uint32_t kHeader = 16 + 4 * N + 512;
  for (uint32_t k = blockIdx.x; k < pl.numBlocks; k += gridDim.x) {
    const uint64_t off = offsets[k];
    const uint32_t wordBytes = meta[k].wordBytes;
    uint8_t *dst = out + off; // only 2-byte aligned from here on (src/mt_rANS32x64_16w_decode.cpp:43,57,64)
    if (k == 0 && tid < 8) { // stream header (:363-379)
      const uint64_t v = tid < 4 ? pl.n : offsets[pl.numBlocks];
      st_u16(out + 2 * tid, (uint32_t)(v >> (16 * (tid & 3))));
    }
    if (tid < 4) { // u64 symbol count of the block (:296-297)
      const uint64_t size = block_end(pl, k) - block_begin(pl, k);
      st_u16(dst + 2 * tid, (uint32_t)(size >> (16 * tid)));
    } else if (tid < 8) { // u64 skip: u16 units from the end of this field to the next header, minus one (:280-283)
      uint64_t skip = ((uint64_t)kHeader - 16 + wordBytes) / 2 - 1;
      if (k + 1 == pl.numBlocks) skip -= 1; // the reference measures its first-written block from its last word slot (:163,279)
      st_u16(dst + 2 * tid, (uint32_t)(skip >> (16 * (tid - 4))));
    }
    for (uint32_t i = tid; i < 2 * N; i += blockDim.x) // states as u16 halves
      st_u16(dst + 16 + 2 * i, meta[k].states[i >> 1] >> (16 * (i & 1)));
    for (uint32_t i = tid; i < 256; i += blockDim.x)
      st_u16(dst + 16 + 4 * N + 2 * i, counts[(uint64_t)k * 256 + i]);
    const uint8_t *src = scratch + (uint64_t)(k + 1) * pl.slotBytes - wordBytes;
    uint8_t *wdst = dst + kHeader;
    for (uint32_t i = tid; i < wordBytes / 2; i += blockDim.x)
      st_u16(wdst + 2 * i, *reinterpret_cast<const uint16_t *>(src + 2 * i));
  }
}

} // namespace hsr

using namespace hsr;

// ------------------------------------------------------------------------------------------------ host side

namespace {

template <int N>
void launch_encode_n(int bits, const uint8_t *dIn, const EncPlan &pl, const uint16_t *dCounts, uint8_t *dScratch, EncBlockMeta *dMeta,
                     uint32_t *dCounter, unsigned grid, cudaStream_t st)
{
  switch (bits) {
  case 10: enc_block_kernel<10, N><<<grid, 32, 0, st>>>(dIn, pl, dCounts, dScratch, dMeta, dCounter); break;
  case 11: enc_block_kernel<11, N><<<grid, 32, 0, st>>>(dIn, pl, dCounts, dScratch, dMeta, dCounter); break;
  case 12: enc_block_kernel<12, N><<<grid, 32, 0, st>>>(dIn, pl, dCounts, dScratch, dMeta, dCounter); break;
  case 13: enc_block_kernel<13, N><<<grid, 32, 0, st>>>(dIn, pl, dCounts, dScratch, dMeta, dCounter); break;
  case 14: enc_block_kernel<14, N><<<grid, 32, 0, st>>>(dIn, pl, dCounts, dScratch, dMeta, dCounter); break;
  default: enc_block_kernel<15, N><<<grid, 32, 0, st>>>(dIn, pl, dCounts, dScratch, dMeta, dCounter); break;
  }
}

bool make_plan(int N, uint64_t n, size_t blockSize, EncPlan *pl)
{
  if (blockSize == 0) blockSize = 65536;
  if (blockSize % (size_t)N || blockSize > (1u << 25) || n < (uint64_t)N) return false; // max block size of the reference, :47-48
  pl->n = n;
  pl->blockSize = (uint32_t)blockSize;
  uint64_t blocks = (n + blockSize - 1) / blockSize;
  // the last block must hold at least one full row: a shorter remainder rides along as the previous block's ragged tail
  if (blocks > 1 && n - (blocks - 1) * blockSize < (uint64_t)N) blocks -= 1;
  if (blocks > 0x7fffffffull) return false;
  pl->numBlocks = (uint32_t)blocks;
  pl->lastStart = (blocks - 1) * blockSize;
  const uint64_t lastLen = n - pl->lastStart;
  const uint64_t maxLen = std::max<uint64_t>(blockSize, lastLen);
  pl->slotBytes = (2 * maxLen + 2 * (uint64_t)N + 15) & ~15ull;
  return true;
}

// Scratch reused across calls (grow-only, one set per device): histograms, per-block word slots, block metadata.
struct EncScratch {
  int device = -1;
  uint16_t *dCounts = nullptr;
  uint8_t *dScratch = nullptr;
  EncBlockMeta *dMeta = nullptr;
  uint64_t *dOffsets = nullptr;
  uint32_t *dCounter = nullptr;
  size_t blocksCap = 0, scratchCap = 0;
  std::mutex mu;

  bool ensure(size_t blocks, size_t scratchBytes)
  {
    if (blocks > blocksCap) {
      cudaFree(dCounts); cudaFree(dMeta); cudaFree(dOffsets);
      dCounts = nullptr; dMeta = nullptr; dOffsets = nullptr; blocksCap = 0;
      const size_t want = blocks + blocks / 8 + 16;
      if (cudaMalloc(&dCounts, want * 512) != cudaSuccess || cudaMalloc(&dMeta, want * sizeof(EncBlockMeta)) != cudaSuccess ||
          cudaMalloc(&dOffsets, (want + 1) * 8) != cudaSuccess)
        return false;
      blocksCap = want;
    }
    if (scratchBytes > scratchCap) {
      cudaFree(dScratch);
      dScratch = nullptr; scratchCap = 0;
      if (cudaMalloc(&dScratch, scratchBytes + scratchBytes / 8) != cudaSuccess) return false;
      scratchCap = scratchBytes + scratchBytes / 8;
    }
    if (!dCounter && cudaMalloc(&dCounter, 16) != cudaSuccess) return false;
    return true;
  }
};

EncScratch *scratch_for_device(int device)
{
  static std::mutex mu;
  static std::vector<EncScratch *> all;
  std::lock_guard<std::mutex> lock(mu);
  for (EncScratch *s : all)
    if (s->device == device) return s;
  EncScratch *s = new EncScratch;
  s->device = device;
  all.push_back(s);
  return s;
}

} // namespace

extern "C" size_t hsr_encode_mt_bound(int N, size_t length, size_t blockSize)
{
  EncPlan pl;
  if (!(N == 32 || N == 64) || !make_plan(N, length, blockSize, &pl)) return 0;
  // one word per symbol is the hard worst case of a 16-bit-word rANS; headers on top
  return 16 + (size_t)pl.numBlocks * (16 + 4 * (size_t)N + 512) + 2 * length + 64;
}

// Device-pointer encode. dOut must hold hsr_encode_mt_bound() bytes (or at least the actual stream, checked after
// the scan); returns the compressed length, 0 on error. Synchronises the stream once to learn that length.
extern "C" size_t hsr_encode_mt_device(int N, int bits, const void *dInV, size_t length, void *dOutV, size_t outCapacity, size_t blockSize,
                                       void *cudaStream)
{
  if (!(N == 32 || N == 64) || bits < 10 || bits > 15 || !dInV || !dOutV) return 0;
  EncPlan pl;
  if (!make_plan(N, length, blockSize, &pl)) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(cudaStream);
  const uint8_t *dIn = static_cast<const uint8_t *>(dInV);
  uint8_t *dOut = static_cast<uint8_t *>(dOutV);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  EncScratch &sc = *scratch_for_device(dev);
  std::lock_guard<std::mutex> lock(sc.mu);
  if (!sc.ensure(pl.numBlocks, (size_t)pl.numBlocks * pl.slotBytes)) {
    (void)cudaGetLastError();
    return 0;
  }
  cudaMemsetAsync(sc.dCounter, 0, 16, st);
  const unsigned gridH = (unsigned)std::min<uint64_t>(pl.numBlocks, (uint64_t)sms * 8);
  const unsigned gridS = (unsigned)std::min<uint64_t>((pl.numBlocks + kSegWarps - 1) / kSegWarps, (uint64_t)sms * 16);
  enc_hist_kernel<<<gridS, kSegWarps * 32, 0, st>>>(dIn, pl, bits, sc.dCounts);
  const unsigned gridE = (unsigned)std::min<uint64_t>(pl.numBlocks, (uint64_t)sms * 32);
  if (N == 32) launch_encode_n<32>(bits, dIn, pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dCounter, gridE, st);
  else launch_encode_n<64>(bits, dIn, pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dCounter, gridE, st);
  enc_scan_kernel<<<1, 1024, 0, st>>>(sc.dMeta, pl.numBlocks, 16 + 4 * (uint32_t)N + 512, sc.dOffsets);
  uint64_t total = 0;
  if (cudaMemcpyAsync(&total, sc.dOffsets + pl.numBlocks, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  if (total > outCapacity) return 0;
  if (N == 32) enc_assemble_kernel<32><<<gridH, 256, 0, st>>>(pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dOffsets, dOut);
  else enc_assemble_kernel<64><<<gridH, 256, 0, st>>>(pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dOffsets, dOut);
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return 0;
  return (size_t)total;
}

// Host-pointer encode with the reference's encoder signature (src/mt_rANS32x64_16w.h:9-14) plus the block size.
extern "C" size_t hsr_encode_mt(int N, int bits, const uint8_t *pInData, size_t length, uint8_t *pOutData, size_t outCapacity, size_t blockSize)
{
  if (!pInData || !pOutData) return 0;
  const size_t bound = hsr_encode_mt_bound(N, length, blockSize);
  if (bound == 0) return 0;
  uint8_t *dIn = nullptr, *dOut = nullptr;
  size_t result = 0;
  if (cudaMalloc(&dIn, length + 16) == cudaSuccess && cudaMalloc(&dOut, bound) == cudaSuccess &&
      cudaMemcpy(dIn, pInData, length, cudaMemcpyHostToDevice) == cudaSuccess) {
    const size_t total = hsr_encode_mt_device(N, bits, dIn, length, dOut, bound, blockSize, nullptr);
    if (total && total <= outCapacity && cudaMemcpy(pOutData, dOut, total, cudaMemcpyDeviceToHost) == cudaSuccess)
      result = total;
  }
  (void)cudaGetLastError();
  cudaFree(dIn);
  cudaFree(dOut);
  return result;
}
