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
//   seg_count_kernel     one CTA per block: byte counts in conflict-free shared-memory counter columns (hsr_hist.cu)
//   seg_normalize_kernel the reference's normalize_hist, one LANE per block (hsr_hist_device.cuh)
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
  uint32_t blockSize;  // symbols per block (multiple of N); fixed-size mode
  uint32_t numBlocks;
  uint64_t lastStart;  // first symbol of the last block (it also owns the < N ragged symbols at the end)
  uint64_t slotBytes;  // scratch bytes per block: worst case one word per symbol (fixed-size mode)
  // policy mode (hsr_encode_mt_policy*): variable blocks from the device block-split pass
  const uint64_t *starts; // numBlocks + 1 symbol offsets, or nullptr in fixed-size mode
  const uint32_t *kinds;  // per block: bit 0 = single-symbol run, bits 8..15 = its symbol
  uint32_t slotPad;       // scratch slack per block in policy mode
};

struct EncBlockMeta {
  uint32_t wordBytes;
  uint32_t pad;
  uint32_t states[64];
};

__device__ __forceinline__ uint64_t block_begin(const EncPlan &pl, uint32_t k) { return pl.starts ? pl.starts[k] : (uint64_t)k * pl.blockSize; }
__device__ __forceinline__ uint64_t block_end(const EncPlan &pl, uint32_t k)
{
  if (pl.starts) return pl.starts[k + 1];
  return k + 1 == pl.numBlocks ? pl.n : (uint64_t)(k + 1) * pl.blockSize;
}
__device__ __forceinline__ bool block_is_run(const EncPlan &pl, uint32_t k) { return pl.kinds && (pl.kinds[k] & 1u); }
// end of block k's scratch slot (its words grow downward from there)
__device__ __forceinline__ uint64_t slot_end(const EncPlan &pl, uint32_t k)
{
  return pl.starts ? 2 * pl.starts[k + 1] + (uint64_t)(k + 1) * pl.slotPad : (uint64_t)(k + 1) * pl.slotBytes;
}

// ---------------------------------------------------------------------------------------------- per-block histograms

// per-block histograms: raw counts by one CTA per block, normalisation by one lane per block (hsr_hist.cu)
bool launch_range_histograms(const uint8_t *dData, const SegPlan &pl, int bits, uint32_t *dCounts32, uint16_t *dCounts, cudaStream_t st);
bool launch_range_counts(const uint8_t *dData, const SegPlan &pl, uint32_t *dCounts32, cudaStream_t st);

static SegPlan seg_plan(const EncPlan &pl) { return SegPlan{pl.starts, pl.kinds, pl.blockSize, pl.n, pl.numBlocks}; }

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
    if (block_is_run(pl, k)) { // single-symbol run: 8 header bytes, no words, no states
      if (lane == 0) meta[k].wordBytes = 0;
      continue;
    }

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

    uint8_t *slotEnd = scratch + slot_end(pl, k); // words grow downward from here
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
__global__ void __launch_bounds__(1024) enc_scan_kernel(const EncBlockMeta *meta, uint32_t numBlocks, uint32_t headerBytes, const uint32_t *kinds,
                                                        uint64_t *offsets)
{
  __shared__ uint64_t sWarp[32];
  __shared__ uint64_t sCarry;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) sCarry = 16; // u64 n, u64 compressed length
  __syncthreads();
  for (uint32_t base = 0; base < numBlocks; base += 1024) {
    const uint32_t k = base + tid;
    uint64_t v = 0;
    if (k < numBlocks) v = (kinds && (kinds[k] & 1u)) ? 8 : (uint64_t)headerBytes + meta[k].wordBytes;
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

// Moves `bytes` (even) of 16-bit words from src to dst, both only 2-byte aligned and misaligned against each other in
// general (the header in front of the words is 16 + 4N + 512 bytes, the scratch slot ends wherever the block's words
// began). dst is brought to a 16-byte boundary with single words; from there every thread writes one aligned 16-byte
// vector assembled from five consecutive 4-byte-aligned source words by funnel shifts (the source is off by 0 or 2
// bytes against a 4-byte grid) — 6 memory instructions per 16 bytes instead of 16. The fifth word of the last vector
// may lie up to 4 bytes past the block's words: inside the scratch allocation (its slots are contiguous and it carries
// slack at the end), read but never stored.
__device__ __forceinline__ void cta_move_words(uint8_t *dst, const uint8_t *src, uint32_t bytes, uint32_t tid, uint32_t threads)
{
  uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
  if (head > bytes) head = bytes;
  for (uint32_t i = tid; i < head / 2; i += threads)
    st_u16(dst + 2 * i, *reinterpret_cast<const uint16_t *>(src + 2 * i));
  dst += head; src += head; bytes -= head;
  const uint32_t vecs = bytes / 16;
  const uint32_t off = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);      // 0 or 2
  const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src - off);
  const uint32_t sh = off * 8u;
  uint4 *d16 = reinterpret_cast<uint4 *>(dst);
  for (uint32_t v = tid; v < vecs; v += threads) {
    const uint32_t *p = s4 + 4 * v;
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3);
    uint4 o;
    if (off) {
      const uint32_t w4 = __ldg(p + 4);
      o = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
    } else
      o = make_uint4(w0, w1, w2, w3);
    d16[v] = o;
  }
  const uint32_t done = vecs * 16;
  for (uint32_t i = tid; i < (bytes - done) / 2; i += threads)
    st_u16(dst + done + 2 * i, *reinterpret_cast<const uint16_t *>(src + done + 2 * i));
}

template <int N>
__global__ void __launch_bounds__(256, 8) enc_assemble_kernel(EncPlan pl, const uint16_t *counts, const uint8_t *scratch, const EncBlockMeta *meta,
                                                           const uint64_t *offsets, uint8_t *out, hsr_block_t *index)
{
  const uint32_t tid = threadIdx.x;
  constexpr uint32_t kHeader = 16 + 4 * N + 512;
  for (uint32_t k = blockIdx.x; k < pl.numBlocks; k += gridDim.x) {
    const uint64_t off = offsets[k];
    const uint32_t wordBytes = meta[k].wordBytes;
    if (index && tid == 0) { // the decoder's unit record of this block: what hsr_mt_index would find by walking the chain
      const uint64_t begin = block_begin(pl, k), end = block_end(pl, k);
      hsr_block_t b{};
      b.outOffset = begin;
      b.count = end - begin;
      if (block_is_run(pl, k)) {
        b.inOffset = off; b.inEnd = off + 8; b.kind = 1; b.symbol = (pl.kinds[k] >> 8) & 0xffu;
      } else {
        b.inOffset = off + 16; b.inEnd = offsets[k + 1]; b.kind = 0;
        b.tail = (uint32_t)((end - begin) % (uint64_t)N); // non-zero only for the block that reaches a ragged end
      }
      index[k] = b;
    }
    uint8_t *dst = out + off; // only 2-byte aligned from here on (src/mt_rANS32x64_16w_decode.cpp:43,57,64)
    if (k == 0 && tid < 8) { // stream header (:363-379)
      const uint64_t v = tid < 4 ? pl.n : offsets[pl.numBlocks];
      st_u16(out + 2 * tid, (uint32_t)(v >> (16 * (tid & 3))));
    }
    if (block_is_run(pl, k)) { // :300-305: size | 1 << 63 | symbol << 54, nothing else
      if (tid < 4) {
        const uint64_t v = (block_end(pl, k) - block_begin(pl, k)) | (1ull << 63) | ((uint64_t)((pl.kinds[k] >> 8) & 0xffu) << 54);
        st_u16(dst + 2 * tid, (uint32_t)(v >> (16 * tid)));
      }
      continue;
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
    cta_move_words(dst + kHeader, scratch + slot_end(pl, k) - wordBytes, wordBytes, tid, blockDim.x);
  }
}


// ---------------------------------------------------------------------------------------------- block-split policy

// the reference grows a block over further MinBlockSize (64 KiB) segments while coding a segment
// with the block's histogram costs less than giving it its own histogram plus half a header (_CanExtendHist,
// src/mt_rANS32x64_16w_encode.cpp:61-136, the log2f cost sums of :101-135 and the threshold of :98-99), and turns
// stretches of one repeated byte into 8-byte run blocks (:171-187, :300-305). The same decisions are taken here on the
// device from per-segment byte counts: one warp per chunk of `segsPerChunk` segments walks its segments FORWARD (the
// reference walks backward from the end of the file; any split decodes) and marks where blocks start.
// Differences, deliberately: blocks never span a chunk (max block size = chunk size, so the split itself is parallel
// and the result keeps >= n / chunk independent blocks for the GPU decoder), the costs use the raw segment counts
// instead of a second, normalised histogram per candidate, and runs are found at segment granularity.
constexpr uint32_t kSegBytes = 65536; // MinBlockSize for every bit width (:34-45)

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
  return v;
}

// flags[t]: 0 = segment t continues the open block, 1 = starts a coded block, 2 | symbol << 8 = starts a run block
__global__ void __launch_bounds__(32) enc_policy_kernel(const uint32_t *segCounts, uint64_t n, uint32_t numSegs, uint32_t segsPerChunk, int bits,
                                                        uint32_t headerBytes, uint32_t *flags)
{
  const uint32_t lane = threadIdx.x & 31u;
  const float replacePoint = (float)((((uint32_t)1 << bits) * (bits == 15 ? 50u : 500u)) >> 12); // HistReplaceMul, :17-29,98-99
  for (uint32_t chunk = blockIdx.x; (uint64_t)chunk * segsPerChunk < numSegs; chunk += gridDim.x) {
    const uint32_t t0 = chunk * segsPerChunk, t1 = min(numSegs, t0 + segsPerChunk);
    float oldLog[8];   // log2 of the open block's first-segment counts (absent symbols count 1, IsSafeHist :193-203)
    float oldLogTotal = 0.f;
    int open = 0;      // 0 none, 1 coded, 2 run
    uint32_t runSym = 0;
    for (uint32_t t = t0; t < t1; t++) {
      const uint64_t begin = (uint64_t)t * kSegBytes, end = t + 1 == numSegs ? n : begin + kSegBytes;
      const uint32_t len = (uint32_t)(end - begin);
      uint32_t c[8], maxC = 0, zeros = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        c[i] = segCounts[(uint64_t)t * 256 + lane + 32 * i];
        maxC = max(maxC, c[i]);
        zeros += c[i] == 0;
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        maxC = max(maxC, __shfl_xor_sync(kFull, maxC, d));
        zeros += __shfl_xor_sync(kFull, zeros, d);
      }
      const bool isRun = maxC == len;
      uint32_t sym = 0;
      if (isRun) {
        uint32_t mine = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (c[i] == len) mine = lane + 32 * i;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mine = min(mine, __shfl_xor_sync(kFull, mine, d));
        sym = mine;
      }
      uint32_t flag;
      if (isRun) {
        flag = (open == 2 && sym == runSym) ? 0u : (2u | (sym << 8));
        open = 2;
        runSym = sym;
      } else {
        bool extend = false;
        if (open == 1) { // :101-135
          const float logLen = __log2f((float)len);
          float before = 0.f, after = 0.f;
#pragma unroll
          for (int i = 0; i < 8; i++)
            if (c[i]) {
              before += (float)(c[i] - 1u) * (oldLogTotal - oldLog[i]);
              after += (float)c[i] * (logLen - __log2f((float)c[i]));
            }
          before = warp_sum(before);
          after = warp_sum(after) + (float)headerBytes * 0.5f;
          extend = (before - after) < replacePoint;
        }
        flag = extend ? 0u : 1u;
        if (!extend) {
#pragma unroll
          for (int i = 0; i < 8; i++) oldLog[i] = __log2f((float)max(c[i], 1u));
          oldLogTotal = __log2f((float)(len + zeros));
        }
        open = 1;
      }
      if (lane == 0) flags[t] = flag;
    }
  }
}

// flags -> block table: starts[b] (symbol offsets, starts[numBlocks] = n), kinds[b]; result[0] = numBlocks
__global__ void __launch_bounds__(1024) enc_blocks_kernel(const uint32_t *flags, uint32_t numSegs, uint64_t n, uint64_t *starts, uint32_t *kinds,
                                                          uint32_t *result)
{
  __shared__ uint32_t sWarp[32];
  __shared__ uint32_t sBase, sTotal;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) sBase = 0;
  __syncthreads();
  for (uint32_t base = 0; base < numSegs; base += 1024) {
    const uint32_t t = base + tid;
    const uint32_t f = t < numSegs ? flags[t] : 0u;
    const uint32_t m = __ballot_sync(kFull, f != 0u);
    if (lane == 0) sWarp[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
      const uint32_t v = sWarp[lane];
      uint32_t incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(kFull, incl, d);
        if (lane >= (uint32_t)d) incl += up;
      }
      sWarp[lane] = incl - v;
      if (lane == 31) sTotal = incl;
    }
    __syncthreads();
    if (f) {
      const uint32_t b = sBase + sWarp[warp] + __popc(m & lanemask_lt());
      starts[b] = (uint64_t)t * kSegBytes;
      kinds[b] = ((f & 2u) ? 1u : 0u) | (f & 0xff00u);
    }
    __syncthreads();
    if (tid == 0) sBase += sTotal;
    __syncthreads();
  }
  if (tid == 0) {
    starts[sBase] = n;
    result[0] = sBase;
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
  pl->starts = nullptr; pl->kinds = nullptr; pl->slotPad = 0;
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
  uint32_t *dCounts32 = nullptr; // raw per-block byte counts between the count and normalise kernels
  uint8_t *dScratch = nullptr;
  EncBlockMeta *dMeta = nullptr;
  uint64_t *dOffsets = nullptr;
  uint32_t *dCounter = nullptr;
  size_t blocksCap = 0, scratchCap = 0;
  // block-split policy pass
  uint32_t *dSegCounts = nullptr, *dFlags = nullptr, *dKinds = nullptr, *dResult = nullptr;
  uint64_t *dStarts = nullptr;
  size_t segsCap = 0;
  std::mutex mu;

  bool ensure_policy(size_t segs)
  {
    if (segs > segsCap) {
      cudaFree(dSegCounts); cudaFree(dFlags); cudaFree(dKinds); cudaFree(dStarts);
      dSegCounts = dFlags = dKinds = nullptr; dStarts = nullptr; segsCap = 0;
      const size_t want = segs + segs / 8 + 16;
      if (cudaMalloc(&dSegCounts, want * 1024) != cudaSuccess || cudaMalloc(&dFlags, want * 4) != cudaSuccess ||
          cudaMalloc(&dKinds, want * 4) != cudaSuccess || cudaMalloc(&dStarts, (want + 1) * 8) != cudaSuccess)
        return false;
      segsCap = want;
    }
    if (!dResult && cudaMalloc(&dResult, 16) != cudaSuccess) return false;
    return true;
  }

  bool ensure(size_t blocks, size_t scratchBytes)
  {
    if (blocks > blocksCap) {
      cudaFree(dCounts); cudaFree(dCounts32); cudaFree(dMeta); cudaFree(dOffsets);
      dCounts = nullptr; dCounts32 = nullptr; dMeta = nullptr; dOffsets = nullptr; blocksCap = 0;
      const size_t want = blocks + blocks / 8 + 16;
      if (cudaMalloc(&dCounts, want * 512) != cudaSuccess || cudaMalloc(&dCounts32, want * 1024) != cudaSuccess || cudaMalloc(&dMeta, want * sizeof(EncBlockMeta)) != cudaSuccess ||
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
static size_t encode_fixed_device(int N, int bits, const void *dInV, size_t length, void *dOutV, size_t outCapacity, size_t blockSize,
                                  hsr_block_t *dIndex, size_t indexCapacity, size_t *numUnits, void *cudaStream)
{
  if (!(N == 32 || N == 64) || bits < 10 || bits > 15 || !dInV || !dOutV) return 0;
  EncPlan pl;
  if (!make_plan(N, length, blockSize, &pl)) return 0;
  if (dIndex && indexCapacity < pl.numBlocks) return 0;
  if (numUnits) *numUnits = pl.numBlocks;
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
  if (!launch_range_histograms(dIn, seg_plan(pl), bits, sc.dCounts32, sc.dCounts, st)) return 0;
  const unsigned gridE = (unsigned)std::min<uint64_t>(pl.numBlocks, (uint64_t)sms * 32);
  if (N == 32) launch_encode_n<32>(bits, dIn, pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dCounter, gridE, st);
  else launch_encode_n<64>(bits, dIn, pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dCounter, gridE, st);
  enc_scan_kernel<<<1, 1024, 0, st>>>(sc.dMeta, pl.numBlocks, 16 + 4 * (uint32_t)N + 512, nullptr, sc.dOffsets);
  uint64_t total = 0;
  if (cudaMemcpyAsync(&total, sc.dOffsets + pl.numBlocks, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  if (total > outCapacity) return 0;
  if (N == 32) enc_assemble_kernel<32><<<gridH, 256, 0, st>>>(pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dOffsets, dOut, dIndex);
  else enc_assemble_kernel<64><<<gridH, 256, 0, st>>>(pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dOffsets, dOut, dIndex);
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return 0;
  return (size_t)total;
}

extern "C" size_t hsr_encode_mt_device(int N, int bits, const void *dIn, size_t length, void *dOut, size_t outCapacity, size_t blockSize,
                                       void *cudaStream)
{
  return encode_fixed_device(N, bits, dIn, length, dOut, outCapacity, blockSize, nullptr, 0, nullptr, cudaStream);
}

// Block-split policy on the device. Same stream format; blocks are whole numbers of 64 KiB
// segments chosen by the reference's cost model (see enc_policy_kernel), at most maxBlockSize bytes each (0 = 256 KiB;
// a multiple of 65536, at most 2^25 = the reference's MaxBlockSize), and stretches of one repeated byte become
// 8-byte run blocks. Every block is still encoded from fresh states, so the result stays independent GPU work.
static size_t encode_policy_device(int N, int bits, const void *dInV, size_t length, void *dOutV, size_t outCapacity, size_t maxBlockSize,
                                   hsr_block_t *dIndex, size_t indexCapacity, size_t *numUnits, void *cudaStream)
{
  if (!(N == 32 || N == 64) || bits < 10 || bits > 15 || !dInV || !dOutV || length < (size_t)N) return 0;
  if (maxBlockSize == 0) maxBlockSize = 4 * (size_t)kSegBytes;
  if (maxBlockSize % kSegBytes || maxBlockSize > (1u << 25)) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(cudaStream);
  const uint8_t *dIn = static_cast<const uint8_t *>(dInV);
  uint8_t *dOut = static_cast<uint8_t *>(dOutV);
  uint64_t segs = (length + kSegBytes - 1) / kSegBytes;
  if (segs > 1 && length - (segs - 1) * kSegBytes < (uint64_t)N) segs -= 1; // a shorter remainder rides along with the last segment
  if (segs > 0x7fffffffull) return 0;
  const uint32_t numSegs = (uint32_t)segs, segsPerChunk = (uint32_t)(maxBlockSize / kSegBytes);
  const uint32_t slotPad = (2u * (uint32_t)N + 31u) & ~15u;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  EncScratch &sc = *scratch_for_device(dev);
  std::lock_guard<std::mutex> lock(sc.mu);
  if (!sc.ensure(numSegs, 2 * length + (size_t)numSegs * slotPad + 64) || !sc.ensure_policy(numSegs)) {
    (void)cudaGetLastError();
    return 0;
  }
  const uint32_t headerBytes = 16 + 4 * (uint32_t)N + 512;
  cudaMemsetAsync(sc.dCounter, 0, 16, st);
  if (!launch_range_counts(dIn, SegPlan{nullptr, nullptr, (uint64_t)kSegBytes, (uint64_t)length, numSegs}, sc.dSegCounts, st)) return 0;
  const uint32_t chunks = (numSegs + segsPerChunk - 1) / segsPerChunk;
  enc_policy_kernel<<<std::min<uint32_t>(chunks, (uint32_t)sms * 32u), 32, 0, st>>>(sc.dSegCounts, length, numSegs, segsPerChunk, bits, headerBytes, sc.dFlags);
  enc_blocks_kernel<<<1, 1024, 0, st>>>(sc.dFlags, numSegs, length, sc.dStarts, sc.dKinds, sc.dResult);
  uint32_t numBlocks = 0;
  if (cudaMemcpyAsync(&numBlocks, sc.dResult, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess ||
      numBlocks == 0 || numBlocks > numSegs) {
    (void)cudaGetLastError();
    return 0;
  }
  if (dIndex && indexCapacity < numBlocks) return 0;
  if (numUnits) *numUnits = numBlocks;
  EncPlan pl{};
  pl.n = length; pl.blockSize = 0; pl.numBlocks = numBlocks; pl.lastStart = 0; pl.slotBytes = 0;
  pl.starts = sc.dStarts; pl.kinds = sc.dKinds; pl.slotPad = slotPad;
  const unsigned gridH = (unsigned)std::min<uint64_t>(numBlocks, (uint64_t)sms * 8);
  if (!launch_range_histograms(dIn, seg_plan(pl), bits, sc.dCounts32, sc.dCounts, st)) return 0;
  const unsigned gridE = (unsigned)std::min<uint64_t>(numBlocks, (uint64_t)sms * 32);
  if (N == 32) launch_encode_n<32>(bits, dIn, pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dCounter, gridE, st);
  else launch_encode_n<64>(bits, dIn, pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dCounter, gridE, st);
  enc_scan_kernel<<<1, 1024, 0, st>>>(sc.dMeta, numBlocks, headerBytes, sc.dKinds, sc.dOffsets);
  uint64_t total = 0;
  if (cudaMemcpyAsync(&total, sc.dOffsets + numBlocks, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  if (total > outCapacity) return 0;
  if (N == 32) enc_assemble_kernel<32><<<gridH, 256, 0, st>>>(pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dOffsets, dOut, dIndex);
  else enc_assemble_kernel<64><<<gridH, 256, 0, st>>>(pl, sc.dCounts, sc.dScratch, sc.dMeta, sc.dOffsets, dOut, dIndex);
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return 0;
  return (size_t)total;
}

extern "C" size_t hsr_encode_mt_policy_device(int N, int bits, const void *dIn, size_t length, void *dOut, size_t outCapacity,
                                              size_t maxBlockSize, void *cudaStream)
{
  return encode_policy_device(N, bits, dIn, length, dOut, outCapacity, maxBlockSize, nullptr, 0, nullptr, cudaStream);
}

// The encoder knows where it put every block: with dIndex (device memory, room for hsr_encode_mt_index_bound()
// records) it also writes the decoder's unit table, one record per block in chain order, so that
// hsr_stream_from_device_indexed() can wrap the stream without walking the header chain again.
extern "C" size_t hsr_encode_mt_device_indexed(int N, int bits, const void *dIn, size_t length, void *dOut, size_t outCapacity, size_t blockSize,
                                               int policy, hsr_block_t *dIndex, size_t indexCapacity, size_t *numUnits, void *cudaStream)
{
  if (!dIndex || !numUnits) return 0;
  return policy ? encode_policy_device(N, bits, dIn, length, dOut, outCapacity, blockSize, dIndex, indexCapacity, numUnits, cudaStream)
                : encode_fixed_device(N, bits, dIn, length, dOut, outCapacity, blockSize, dIndex, indexCapacity, numUnits, cudaStream);
}

extern "C" size_t hsr_encode_mt_index_bound(int N, size_t length, size_t blockSize)
{
  EncPlan pl;
  if (!(N == 32 || N == 64) || !make_plan(N, length, blockSize ? blockSize : 65536, &pl)) return 0;
  return (size_t)pl.numBlocks + 1; // policy mode merges 64 KiB segments, so the fixed 64 KiB count bounds it too
}

// Host-pointer form with the reference's encoder signature plus the maximum block size.
extern "C" size_t hsr_encode_mt_policy(int N, int bits, const uint8_t *pInData, size_t length, uint8_t *pOutData, size_t outCapacity,
                                       size_t maxBlockSize)
{
  if (!pInData || !pOutData) return 0;
  const size_t bound = hsr_encode_mt_bound(N, length, 0); // fixed 64 KiB blocks are the worst case for header bytes
  if (bound == 0) return 0;
  uint8_t *dIn = nullptr, *dOut = nullptr;
  size_t result = 0;
  if (cudaMalloc(&dIn, length + 16) == cudaSuccess && cudaMalloc(&dOut, bound) == cudaSuccess &&
      cudaMemcpy(dIn, pInData, length, cudaMemcpyHostToDevice) == cudaSuccess) {
    const size_t total = hsr_encode_mt_policy_device(N, bits, dIn, length, dOut, bound, maxBlockSize, nullptr);
    if (total && total <= outCapacity && cudaMemcpy(pOutData, dOut, total, cudaMemcpyDeviceToHost) == cudaSuccess)
      result = total;
  }
  (void)cudaGetLastError();
  cudaFree(dIn);
  cudaFree(dOut);
  return result;
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
