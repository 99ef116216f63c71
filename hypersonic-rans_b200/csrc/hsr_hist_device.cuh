// hsr_hist_device.cuh — device pieces of the histogram path shared by hsr_hist.cu and hsr_encode.cu:
// conflict-free shared-memory byte counting (src/hist.cpp:8-14) and the reference's normalize_hist (:16-215) restated
// for one CTA (one histogram) and for one lane per histogram (many). See hsr_hist.cu for the rationale.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace hsr {

constexpr int kHistThreads = 256; // normalize_kernel (one histogram, one CTA)

// ---- conflict-free byte counting (observe_hist, src/hist.cpp:8-14). Counters live in lane-private COLUMNS: word
// (bin, lane) sits at bin * 32 + lane, so lane l only ever touches bank l — a warp's 32 increments never share a bank,
// whatever the bytes are, and every shared-memory atomic is one conflict-free request (one histogram per warp pays one
// pass per distinct word of the busiest bank: 2.6-3.0 TB/s on Zipf(1)/uniform bytes; this layout 3.6 TB/s on any input,
// bound by the ~2.3 cycles a conflict-free ATOMS costs the SM's shared-memory pipe; profiles/r2/ubench_hist.jsonl).
// The two warps of a 64-thread CTA share one 32 KB plane through the u16 halves of each word (warp w adds 1 << 16 w);
// a thread counts at most kCntEpochVecs * 16 + 2 < 65536 bytes between two flushes, so a half never overflows.
constexpr int kCntThreads = 64;
constexpr int kCntInflight = 4;       // 16-byte loads requested per lane before the first is counted
constexpr int kCntEpochVecs = 4032;   // 64,512 bytes per thread and epoch
constexpr int kCntPlaneBytes = 256 * 32 * 4;

__device__ __forceinline__ void cnt_red(uint32_t addr, uint32_t inc) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(inc) : "memory"); }

__device__ __forceinline__ void cnt_count16(uint32_t base, uint32_t inc, const uint4 &q)
{
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; j++) {
#pragma unroll
    for (int b = 0; b < 4; b++)
      cnt_red(__byte_perm(w[j], 0, 0x4440 + b) * 128u + base, inc);
  }
}

// Counts the bytes p[0 .. len) that fall to CTA `rank` of the `ctas` CTAs sharing this range (16-byte vectors dealt
// round-robin over all their threads; the unaligned head and tail go to rank 0). All kCntThreads threads call.
// total[t] receives this CTA's count of bin threadIdx.x + 64 t. sPlane: kCntPlaneBytes of shared memory.
__device__ inline void cta_count(const uint8_t *p, uint64_t len, uint64_t rank, uint64_t ctas, uint32_t *sPlane, uint32_t total[4])
{
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sPlane) + lane * 4u;
  const uint32_t inc = 1u << (16u * warp);
  uint64_t head = (16u - (reinterpret_cast<uintptr_t>(p) & 15u)) & 15u; // bytes before the first 16-byte boundary
  if (head > len) head = len;
  const uint4 *v = reinterpret_cast<const uint4 *>(p + head);
  const uint64_t vecs = (len - head) / 16;
  const uint64_t stride = ctas * kCntThreads;
  total[0] = total[1] = total[2] = total[3] = 0;
  uint64_t i = rank * kCntThreads + tid;
  bool first = true;
  do {
    for (uint32_t k = tid * 4u; k < 256u * 32u; k += kCntThreads * 4u)
      *reinterpret_cast<uint4 *>(sPlane + k) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (first && rank == 0) { // fewer than 32 bytes in all
      if (tid < head) cnt_red((uint32_t)p[tid] * 128u + base, inc);
      const uint64_t done = head + vecs * 16;
      if (tid < len - done) cnt_red((uint32_t)p[done + tid] * 128u + base, inc);
    }
    first = false;
    for (uint32_t e = 0; e < (uint32_t)kCntEpochVecs && i < vecs; e += kCntInflight, i += stride * kCntInflight) {
      uint4 q[kCntInflight];
#pragma unroll
      for (int u = 0; u < kCntInflight; u++)
        if (i + u * stride < vecs) q[u] = __ldg(v + i + u * stride);
#pragma unroll
      for (int u = 0; u < kCntInflight; u++)
        if (i + u * stride < vecs) cnt_count16(base, inc, q[u]);
    }
    __syncthreads();
    // bin b: 32 lane words, read with a rotation that keeps the CTA's threads out of each other's banks
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const uint32_t b = tid + 64u * t;
      uint32_t sum = 0;
#pragma unroll 8
      for (uint32_t j = 0; j < 32u; j++) {
        const uint32_t w = sPlane[b * 32u + ((j + tid) & 31u)];
        sum += (w & 0xffffu) + (w >> 16);
      }
      total[t] += sum;
    }
  } while (__syncthreads_or(i < vecs)); // another epoch only beyond 64,512 bytes per thread
}

// ---- ranges of a buffer that each get their own histogram (the block_/mt_ per-block histograms): either fixed-size
// pieces of `blockSize` bytes or explicit starts[num + 1]; kinds[k] & 1 marks a single-symbol run (no histogram).
struct SegPlan {
  const uint64_t *starts;
  const uint32_t *kinds;
  uint64_t blockSize, n;
  uint32_t num;
};
__device__ __forceinline__ uint64_t seg_begin(const SegPlan &pl, uint32_t k) { return pl.starts ? pl.starts[k] : (uint64_t)k * pl.blockSize; }
__device__ __forceinline__ uint64_t seg_end(const SegPlan &pl, uint32_t k)
{
  if (pl.starts) return pl.starts[k + 1];
  return k + 1 == pl.num ? pl.n : (uint64_t)(k + 1) * pl.blockSize;
}
__device__ __forceinline__ bool seg_is_run(const SegPlan &pl, uint32_t k) { return pl.kinds && (pl.kinds[k] & 1u); }

// one 64-thread CTA at a time per range: raw u32 counts[k][256]
__device__ inline void cta_count_ranges(const uint8_t *data, const SegPlan &pl, uint32_t *counts32, uint32_t *sPlane)
{
  for (uint32_t k = blockIdx.x; k < pl.num; k += gridDim.x) {
    if (seg_is_run(pl, k)) continue;
    const uint64_t begin = seg_begin(pl, k), end = seg_end(pl, k);
    uint32_t total[4];
    cta_count(data + begin, end - begin, 0, 1, sPlane, total);
#pragma unroll
    for (int t = 0; t < 4; t++) counts32[(uint64_t)k * 256 + threadIdx.x + 64 * t] = total[t];
  }
}

// src/hist.cpp:112-129
__device__ inline void heapify(uint8_t *idx, const uint16_t *val, int n, int i)
{
  for (;;) {
    const int left = 2 * i + 1, right = 2 * i + 2;
    int largest = i;
    if (left < n && val[idx[left]] > val[idx[largest]]) largest = left;
    if (right < n && val[idx[right]] > val[idx[largest]]) largest = right;
    if (largest == i) return;
    const uint8_t t = idx[i]; idx[i] = idx[largest]; idx[largest] = t;
    i = largest;
  }
}

// normalises sHist[256] (u32 counts) to sum 2^bits; all threads call, thread 0 does the sequential part
__device__ inline void cta_normalize(const uint32_t *sHist, uint64_t dataBytes, int bits, uint16_t *sCapped, uint8_t *sIdx,
                              uint16_t *outCount, uint16_t *outCumul)
{
  const uint32_t total = 1u << bits;
  const float mul = __fdiv_rn((float)total, __ull2float_rn(dataBytes)); // :60
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const float scaled = __fadd_rn(__fmul_rn(__uint2float_rn(sHist[i]), mul), 0.5f); // :64, MUL then ADD
    uint16_t c = (uint16_t)__float2uint_rz(scaled);
    if (c == 0 && sHist[i]) c = 1; // :66-67
    sCapped[i] = c;
    sIdx[i] = (uint8_t)i;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t sum = 0;
    for (int i = 0; i < 256; i++) sum += sCapped[i];
    if (sum != total) { // :103
      for (int i = 256 / 2 - 1; i >= 0; i--) heapify(sIdx, sCapped, 256, i); // :133-134
      for (int i = 255; i >= 0; i--) {                                        // :136-140
        const uint8_t t = sIdx[0]; sIdx[0] = sIdx[i]; sIdx[i] = t;
        heapify(sIdx, sCapped, i, 0);
      }
      int minTwo = 0; // :145-154
      for (int i = 0; i < 256; i++)
        if (sCapped[sIdx[i]] >= 2) { minTwo = i; break; }
      bool ready = false;
      while (!ready && sum > total) { // :156-176
        for (int i = minTwo; i < 256; i++) {
          sCapped[sIdx[i]]--; sum--;
          if (sum == total) { ready = true; break; }
        }
        if (ready) break;
        for (int i = minTwo; i < 256; i++)
          if (sCapped[sIdx[i]] >= 2) { minTwo = i; break; }
      }
      while (!ready && sum < total) { // :178-198
        for (int i = 255; i >= minTwo; i--) {
          sCapped[sIdx[i]]++; sum++;
          if (sum == total) { ready = true; break; }
        }
        if (ready) break;
        for (int i = minTwo; i < 256; i++)
          if (sCapped[sIdx[i]] >= 2) { minTwo = i; break; }
      }
    }
    uint32_t counter = 0; // :201-209
    for (int i = 0; i < 256; i++) {
      if (outCumul) outCumul[i] = (uint16_t)counter;
      outCount[i] = sCapped[i];
      counter += sCapped[i];
    }
  }
  __syncthreads();
}


// ---- normalize_hist (src/hist.cpp:16-215) with ONE LANE per histogram: the order-dependent part (index heap-sort,
// steal / charity loops) is serial per histogram, so a warp runs 32 of them side by side instead of parking 31 lanes
// behind lane 0. Every lane executes the same literal sequence on its own histogram (same comparisons, same tie
// order as the reference); the SIMT hardware serialises only where their control flow differs (sift depths, loop
// trip counts). Scratch is lane-interleaved shared memory — element i of lane l at index i * 32 + l — so the u32 heap
// never has two lanes on one bank. heap: 256 * 32 u32, capped: 256 * 32 u16 (48 KB per warp).
constexpr int kLaneNormSmemBytes = 256 * 32 * 4 + 256 * 32 * 2;

__device__ __forceinline__ void lane_heapify(uint32_t *heap, uint32_t lane, int n, int i)
{
  uint32_t cur = heap[i * 32 + lane];
  for (;;) {
    const int left = 2 * i + 1, right = 2 * i + 2;
    int largest = i;
    uint32_t big = cur;
    if (left < n) {
      const uint32_t l = heap[left * 32 + lane];
      if ((l >> 8) > (big >> 8)) { largest = left; big = l; }
    }
    if (right < n) {
      const uint32_t r = heap[right * 32 + lane];
      if ((r >> 8) > (big >> 8)) { largest = right; big = r; }
    }
    if (largest == i) break;
    heap[i * 32 + lane] = big;
    heap[largest * 32 + lane] = cur;
    i = largest;
  }
}

// counts32: this lane's 256 raw counts (global memory); writes 256 u16 normalised counts to outCount
__device__ inline void lane_normalize(const uint32_t *counts32, uint64_t dataBytes, int bits, uint32_t *heap, uint16_t *capped, uint16_t *outCount,
                                      uint32_t lane)
{
  const uint32_t total = 1u << bits;
  const float mul = __fdiv_rn((float)total, __ull2float_rn(dataBytes)); // :60
  uint32_t sum = 0;
#pragma unroll 8
  for (int i = 0; i < 256; i++) {
    const uint32_t cnt = counts32[i];
    const float scaled = __fadd_rn(__fmul_rn(__uint2float_rn(cnt), mul), 0.5f); // :64, MUL then ADD, never an FMA
    uint32_t c = __float2uint_rz(scaled) & 0xffffu;
    if (c == 0 && cnt) c = 1; // :66-67
    capped[i * 32 + lane] = (uint16_t)c;
    heap[i * 32 + lane] = (c << 8) | (uint32_t)i;
    sum += c;
  }
  if (sum != total) { // :103
    for (int i = 256 / 2 - 1; i >= 0; i--) lane_heapify(heap, lane, 256, i); // :133-134
    for (int i = 255; i >= 0; i--) {                                          // :136-140
      const uint32_t t = heap[lane]; heap[lane] = heap[i * 32 + lane]; heap[i * 32 + lane] = t;
      lane_heapify(heap, lane, i, 0);
    }
    // heap[] is now the ascending order; the packed values go stale once capped[] changes, so capped[] is used below
    int minTwo = 0; // :145-154
    for (int i = 0; i < 256; i++)
      if (capped[(heap[i * 32 + lane] & 0xffu) * 32 + lane] >= 2) { minTwo = i; break; }
    bool ready = false;
    while (!ready && sum > total) { // :156-176
      for (int i = minTwo; i < 256; i++) {
        capped[(heap[i * 32 + lane] & 0xffu) * 32 + lane]--; sum--;
        if (sum == total) { ready = true; break; }
      }
      if (ready) break;
      for (int i = minTwo; i < 256; i++)
        if (capped[(heap[i * 32 + lane] & 0xffu) * 32 + lane] >= 2) { minTwo = i; break; }
    }
    while (!ready && sum < total) { // :178-198
      for (int i = 255; i >= minTwo; i--) {
        capped[(heap[i * 32 + lane] & 0xffu) * 32 + lane]++; sum++;
        if (sum == total) { ready = true; break; }
      }
      if (ready) break;
      for (int i = minTwo; i < 256; i++)
        if (capped[(heap[i * 32 + lane] & 0xffu) * 32 + lane] >= 2) { minTwo = i; break; }
    }
  }
  for (int i = 0; i < 256; i++) outCount[i] = capped[i * 32 + lane];
}

// one lane per range of `pl`: counts32[k][256] -> outCounts[k][256]
__device__ inline void warp_normalize_ranges(const uint32_t *counts32, const SegPlan &pl, int bits, uint16_t *outCounts, uint8_t *smem)
{
  uint32_t *heap = reinterpret_cast<uint32_t *>(smem);
  uint16_t *capped = reinterpret_cast<uint16_t *>(smem + 256 * 32 * 4);
  const uint32_t lane = threadIdx.x & 31u;
  for (uint64_t k0 = (uint64_t)blockIdx.x * 32; k0 < pl.num; k0 += (uint64_t)gridDim.x * 32) {
    const uint64_t k = k0 + lane;
    if (k < pl.num && !seg_is_run(pl, (uint32_t)k))
      lane_normalize(counts32 + k * 256, seg_end(pl, (uint32_t)k) - seg_begin(pl, (uint32_t)k), bits, heap, capped, outCounts + k * 256, lane);
    __syncwarp();
  }
}

} // namespace hsr
