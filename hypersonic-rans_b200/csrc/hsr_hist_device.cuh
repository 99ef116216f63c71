// hsr_hist_device.cuh — device pieces of the histogram path shared by hsr_hist.cu and hsr_encode.cu:
// privatised shared-memory byte counting and the reference's normalize_hist restated for one CTA
// (src/hist.cpp:8-14 and :16-215). See hsr_hist.cu for the rationale.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace hsr {

constexpr int kHistThreads = 256;
constexpr int kHistWarps = kHistThreads / 32;

__device__ __forceinline__ void count4(uint32_t *h, uint32_t v)
{
  atomicAdd(h + (v & 0xffu), 1u);
  atomicAdd(h + ((v >> 8) & 0xffu), 1u);
  atomicAdd(h + ((v >> 16) & 0xffu), 1u);
  atomicAdd(h + (v >> 24), 1u);
}

// counts bytes [begin, end) of data into the CTA's privatised histograms, then reduces them into sOut[256]
__device__ inline void cta_observe(const uint8_t *data, uint64_t begin, uint64_t end, uint32_t (*sPriv)[256], uint32_t *sOut,
                            uint32_t ctaRank, uint32_t ctaCount)
{
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  for (int k = tid; k < kHistWarps * 256; k += kHistThreads)
    (&sPriv[0][0])[k] = 0;
  __syncthreads();
  uint32_t *mine = sPriv[warp];

  // unaligned head up to the first 16-byte boundary, handled by CTA 0
  const uint8_t *p = data + begin;
  uint64_t len = end - begin;
  uint64_t head = (16u - (reinterpret_cast<uintptr_t>(p) & 15u)) & 15u;
  if (head > len) head = len;
  if (ctaRank == 0 && tid < head)
    atomicAdd(mine + p[tid], 1u);
  const uint4 *v = reinterpret_cast<const uint4 *>(p + head);
  const uint64_t vecs = (len - head) / 16;
  for (uint64_t i = (uint64_t)ctaRank * kHistThreads + tid; i < vecs; i += (uint64_t)ctaCount * kHistThreads) {
    const uint4 q = __ldg(v + i);
    count4(mine, q.x); count4(mine, q.y); count4(mine, q.z); count4(mine, q.w);
  }
  const uint64_t done = head + vecs * 16;
  if (ctaRank == 0 && tid < len - done)
    atomicAdd(mine + p[done + tid], 1u);
  __syncthreads();
  for (int b = tid; b < 256; b += kHistThreads) {
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < kHistWarps; w++) s += sPriv[w][b];
    sOut[b] = s;
  }
  __syncthreads();
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


// ---- one WARP per segment: the order-dependent part of normalize_hist runs on one lane whatever the CTA size, so
// many short segments (per-block histograms) are best served by as many concurrent warps as the SM holds.

// counts bytes [begin, end) into the warp's private histogram h[256] (shared memory, zeroed here)
__device__ inline void warp_observe(const uint8_t *data, uint64_t begin, uint64_t end, uint32_t *h, uint32_t lane)
{
  for (int k = lane; k < 256; k += 32) h[k] = 0;
  __syncwarp();
  const uint8_t *p = data + begin;
  const uint64_t len = end - begin;
  uint64_t head = (16u - (reinterpret_cast<uintptr_t>(p) & 15u)) & 15u;
  if (head > len) head = len;
  if (lane < head) atomicAdd(h + p[lane], 1u);
  const uint4 *v = reinterpret_cast<const uint4 *>(p + head);
  const uint64_t vecs = (len - head) / 16;
  for (uint64_t i = lane; i < vecs; i += 32) {
    const uint4 q = __ldg(v + i);
    count4(h, q.x); count4(h, q.y); count4(h, q.z); count4(h, q.w);
  }
  const uint64_t done = head + vecs * 16;
  if (lane < len - done) atomicAdd(h + p[done + lane], 1u);
  __syncwarp();
}

// heapify of src/hist.cpp:112-129 on packed heap slots: slot = capped value << 8 | symbol index, so one shared-memory
// load per child replaces the reference's pVal[pIdx[child]] double indirection. Comparisons use the value only
// (strict >, exactly like the reference), so the resulting order — ties included — is the reference's.
__device__ inline void heapify_packed(uint32_t *heap, int n, int i)
{
  uint32_t cur = heap[i];
  for (;;) {
    const int left = 2 * i + 1, right = 2 * i + 2;
    int largest = i;
    uint32_t big = cur;
    if (left < n) {
      const uint32_t l = heap[left];
      if ((l >> 8) > (big >> 8)) { largest = left; big = l; }
    }
    if (right < n) {
      const uint32_t r = heap[right];
      if ((r >> 8) > (big >> 8)) { largest = right; big = r; }
    }
    if (largest == i) break;
    heap[i] = big;       // std::swap(pIdx[i], pIdx[largest]) ...
    heap[largest] = cur; // ... the moved element keeps sinking
    i = largest;
  }
}

// normalize_hist (src/hist.cpp:16-215) by one warp: lanes scale, lane 0 runs the sort and the steal/charity loops.
// `heap` is 256 u32 of scratch (it may alias the histogram `h`, which is dead once the scaled counts exist).
__device__ inline void warp_normalize(const uint32_t *h, uint64_t dataBytes, int bits, uint16_t *capped, uint32_t *heap, uint16_t *outCount,
                                      uint32_t lane)
{
  const uint32_t total = 1u << bits;
  const float mul = __fdiv_rn((float)total, __ull2float_rn(dataBytes));
  uint32_t part = 0;
  uint16_t mine[8];
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const int i = lane + 32 * t;
    const uint32_t cnt = h[i];
    const float scaled = __fadd_rn(__fmul_rn(__uint2float_rn(cnt), mul), 0.5f);
    uint16_t c = (uint16_t)__float2uint_rz(scaled);
    if (c == 0 && cnt) c = 1;
    mine[t] = c;
    part += c;
  }
  __syncwarp(); // every lane has read its counts: `heap` may now overwrite `h`
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const int i = lane + 32 * t;
    capped[i] = mine[t];
    heap[i] = ((uint32_t)mine[t] << 8) | (uint32_t)i;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  __syncwarp();
  if (lane == 0 && part != total) {
    uint32_t sum = part;
    for (int i = 256 / 2 - 1; i >= 0; i--) heapify_packed(heap, 256, i); // :133-134
    for (int i = 255; i >= 0; i--) {                                      // :136-140
      const uint32_t t = heap[0]; heap[0] = heap[i]; heap[i] = t;
      heapify_packed(heap, i, 0);
    }
    // heap[] is now the ascending order; the packed values are stale once capped[] changes, so use capped[] below
    int minTwo = 0;
    for (int i = 0; i < 256; i++)
      if (capped[heap[i] & 0xffu] >= 2) { minTwo = i; break; }
    bool ready = false;
    while (!ready && sum > total) {
      for (int i = minTwo; i < 256; i++) {
        capped[heap[i] & 0xffu]--; sum--;
        if (sum == total) { ready = true; break; }
      }
      if (ready) break;
      for (int i = minTwo; i < 256; i++)
        if (capped[heap[i] & 0xffu] >= 2) { minTwo = i; break; }
    }
    while (!ready && sum < total) {
      for (int i = 255; i >= minTwo; i--) {
        capped[heap[i] & 0xffu]++; sum++;
        if (sum == total) { ready = true; break; }
      }
      if (ready) break;
      for (int i = minTwo; i < 256; i++)
        if (capped[heap[i] & 0xffu] >= 2) { minTwo = i; break; }
    }
  }
  __syncwarp();
  for (int i = lane; i < 256; i += 32) outCount[i] = capped[i];
  __syncwarp();
}

} // namespace hsr
