// hsr_hist.cu — device histogram count + normalise, bit-exact with the reference's make_hist
// (src/hist.cpp:217-222 = observe_hist :8-14 + normalize_hist :16-215).
//
// observe: privatised shared-memory histograms (one copy per warp, u32 bins, shared-memory atomics), 16-byte
// coalesced loads, one global atomicAdd per bin per CTA at the end.
// normalise: 256 elements of order-dependent integer logic. The float scale is done with explicit round-to-
// nearest MUL and ADD (never an FMA — the reference build has no FMA target, src/hist.cpp:60-64); the index
// heap-sort and the steal/charity loops (:105-198) run literally, on one thread, because the tie order of that
// particular unstable sort decides which symbols receive the +-1.
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/hsrans_b200.h"

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
__device__ void cta_observe(const uint8_t *data, uint64_t begin, uint64_t end, uint32_t (*sPriv)[256], uint32_t *sOut,
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

__global__ void __launch_bounds__(kHistThreads) observe_kernel(const uint8_t *data, uint64_t size, uint32_t *hist)
{
  __shared__ uint32_t sPriv[kHistWarps][256];
  __shared__ uint32_t sOut[256];
  cta_observe(data, 0, size, sPriv, sOut, blockIdx.x, gridDim.x);
  for (int b = threadIdx.x; b < 256; b += kHistThreads)
    if (sOut[b]) atomicAdd(hist + b, sOut[b]);
}

// src/hist.cpp:112-129
__device__ void heapify(uint8_t *idx, const uint16_t *val, int n, int i)
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
__device__ void cta_normalize(const uint32_t *sHist, uint64_t dataBytes, int bits, uint16_t *sCapped, uint8_t *sIdx,
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

__global__ void __launch_bounds__(kHistThreads) normalize_kernel(const uint32_t *hist, uint64_t dataBytes, int bits,
                                                                 uint16_t *symbolCount, uint16_t *cumul)
{
  __shared__ uint32_t sHist[256];
  __shared__ uint16_t sCapped[256];
  __shared__ uint8_t sIdx[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sHist[i] = hist[i];
  __syncthreads();
  cta_normalize(sHist, dataBytes, bits, sCapped, sIdx, symbolCount, cumul);
}

// one CTA per segment: count the segment's bytes, normalise, write its 256 u16 counts
__global__ void __launch_bounds__(kHistThreads) segments_kernel(const uint8_t *data, uint64_t size, uint64_t segmentBytes,
                                                                int bits, uint16_t *symbolCounts)
{
  __shared__ uint32_t sPriv[kHistWarps][256];
  __shared__ uint32_t sOut[256];
  __shared__ uint16_t sCapped[256];
  __shared__ uint8_t sIdx[256];
  const uint64_t segments = (size + segmentBytes - 1) / segmentBytes;
  for (uint64_t seg = blockIdx.x; seg < segments; seg += gridDim.x) {
    const uint64_t begin = seg * segmentBytes;
    const uint64_t end = begin + segmentBytes < size ? begin + segmentBytes : size;
    cta_observe(data, begin, end, sPriv, sOut, 0, 1);
    cta_normalize(sOut, end - begin, bits, sCapped, sIdx, symbolCounts + seg * 256, nullptr);
  }
}

} // namespace hsr

using namespace hsr;

extern "C" int hsr_observe_hist_device(const void *dData, size_t size, uint32_t *dHist, void *cudaStream)
{
  cudaStream_t st = static_cast<cudaStream_t>(cudaStream);
  if (!dHist || (!dData && size)) return -1;
  if (cudaMemsetAsync(dHist, 0, 256 * sizeof(uint32_t), st) != cudaSuccess) return -2;
  if (size == 0) return 0;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint64_t per = (uint64_t)kHistThreads * 16 * 8;
  uint64_t grid = (size + per - 1) / per;
  if (grid > (uint64_t)sms * 8) grid = (uint64_t)sms * 8;
  observe_kernel<<<(unsigned)grid, kHistThreads, 0, st>>>(static_cast<const uint8_t *>(dData), size, dHist);
  return cudaGetLastError() == cudaSuccess ? 1 : -2;
}

extern "C" int hsr_normalize_hist_device(const uint32_t *dHist, size_t dataBytes, int bits, uint16_t *dSymbolCount,
                                         uint16_t *dCumul, void *cudaStream)
{
  if (!dHist || !dSymbolCount || bits < 1 || bits > 15 || dataBytes == 0) return -1;
  normalize_kernel<<<1, kHistThreads, 0, static_cast<cudaStream_t>(cudaStream)>>>(dHist, dataBytes, bits, dSymbolCount, dCumul);
  return cudaGetLastError() == cudaSuccess ? 1 : -2;
}

extern "C" int hsr_make_hist_segments_device(const void *dData, size_t size, size_t segmentBytes, int bits,
                                             uint16_t *dSymbolCounts, void *cudaStream)
{
  if (!dData || !dSymbolCounts || size == 0 || segmentBytes == 0 || bits < 1 || bits > 15) return -1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint64_t segments = (size + segmentBytes - 1) / segmentBytes;
  const unsigned grid = (unsigned)(segments < (uint64_t)sms * 8 ? segments : (uint64_t)sms * 8);
  segments_kernel<<<grid, kHistThreads, 0, static_cast<cudaStream_t>(cudaStream)>>>(static_cast<const uint8_t *>(dData), size, segmentBytes,
                                                                                bits, dSymbolCounts);
  return cudaGetLastError() == cudaSuccess ? 1 : -2;
}

extern "C" int hsr_make_hist(const uint8_t *pData, size_t size, int bits, uint16_t symbolCount[256], uint16_t cumul[256])
{
  if (!pData || size == 0 || !symbolCount || bits < 1 || bits > 15) return -1;
  uint8_t *dData = nullptr;
  uint32_t *dHist = nullptr;
  uint16_t *dOut = nullptr;
  int rc = -2;
  if (cudaMalloc(&dData, size) == cudaSuccess && cudaMalloc(&dHist, 1024) == cudaSuccess && cudaMalloc(&dOut, 1024) == cudaSuccess &&
      cudaMemcpy(dData, pData, size, cudaMemcpyHostToDevice) == cudaSuccess && hsr_observe_hist_device(dData, size, dHist, nullptr) > 0 &&
      hsr_normalize_hist_device(dHist, size, bits, dOut, dOut + 256, nullptr) > 0) {
    uint16_t host[512];
    if (cudaMemcpy(host, dOut, 1024, cudaMemcpyDeviceToHost) == cudaSuccess) {
      memcpy(symbolCount, host, 512);
      if (cumul) memcpy(cumul, host + 256, 512);
      rc = 0;
    }
  }
  (void)cudaGetLastError();
  cudaFree(dData); cudaFree(dHist); cudaFree(dOut);
  return rc;
}
