// hsr_hist.cu — device histogram count + normalise, bit-exact with the reference's make_hist
// (src/hist.cpp:217-222 = observe_hist :8-14 + normalize_hist :16-215).
//
// observe: lane-private counter columns in shared memory (conflict-free shared-memory atomics whatever the bytes are),
// four 16-byte coalesced loads in flight per lane, one global atomicAdd per bin per CTA at the end.
// normalise: 256 elements of order-dependent integer logic. The float scale is done with explicit round-to-
// nearest MUL and ADD (never an FMA — the reference build has no FMA target, src/hist.cpp:60-64); the index
// heap-sort and the steal/charity loops (:105-198) run literally, on one thread, because the tie order of that
// particular unstable sort decides which symbols receive the +-1.
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/hsrans_b200.h"
#include "hsr_hist_device.cuh"

namespace hsr {

// observe_hist (src/hist.cpp:8-14) at streaming speed. Counters live in lane-private COLUMNS: word (bin, lane) sits at
// bin * 32 + lane, so lane l only ever touches bank l — a warp's 32 increments never share a bank, whatever the bytes
// are, and every shared-memory atomic is one conflict-free request (round 1's one-histogram-per-warp layout pays one
// pass per distinct word of the busiest bank: 2.6-3.0 TB/s on Zipf(1)/uniform bytes; this layout 3.6 TB/s on any input,
// bound by the ~2.3 cycles a conflict-free ATOMS costs the SM's shared-memory pipe; profiles/r2/ubench_hist.jsonl).
// The two warps of a CTA share the 32 KB plane through the u16 halves of each word (warp w adds 1 << 16 w); a thread
// counts at most kObsEpochVecs * 16 < 65536 bytes between two flushes, so a half can never overflow into its neighbour.
constexpr int kObsThreads = 64;
constexpr int kObsInflight = 4;       // 16-byte loads requested per lane before the first is counted
constexpr int kObsEpochVecs = 4032;   // 64,512 bytes per thread and epoch
constexpr int kObsPlaneBytes = 256 * 32 * 4;

__device__ __forceinline__ void obs_count16(uint32_t base, uint32_t inc, const uint4 &q)
{
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; j++) {
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const uint32_t byte = __byte_perm(w[j], 0, 0x4440 + b);
      asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(byte * 128u + base), "r"(inc) : "memory");
    }
  }
}

__global__ void __launch_bounds__(kObsThreads) observe_kernel(const uint8_t *data, uint64_t size, uint32_t *hist)
{
  extern __shared__ __align__(16) uint32_t sPlane[]; // [256 bins][32 lanes], u16 halves per warp
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t sBase = (uint32_t)__cvta_generic_to_shared(sPlane);
  const uint32_t base = sBase + lane * 4u;
  const uint32_t inc = 1u << (16u * warp);

  uint64_t head = (16u - (reinterpret_cast<uintptr_t>(data) & 15u)) & 15u; // bytes before the first 16-byte boundary
  if (head > size) head = size;
  const uint4 *v = reinterpret_cast<const uint4 *>(data + head);
  const uint64_t vecs = (size - head) / 16;
  const uint64_t stride = (uint64_t)gridDim.x * kObsThreads;
  uint32_t total[4] = {0, 0, 0, 0}; // this thread's bins tid, tid + 64, tid + 128, tid + 192 over all epochs

  uint64_t i = (uint64_t)blockIdx.x * kObsThreads + tid;
  bool first = true;
  do {
    for (uint32_t k = tid; k < 256u * 32u; k += kObsThreads) sPlane[k] = 0;
    __syncthreads();
    if (first && blockIdx.x == 0) { // the unaligned head and tail bytes: CTA 0, once (fewer than 32 bytes in all)
      if (tid < head) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((uint32_t)data[tid] * 128u + base), "r"(inc) : "memory");
      const uint64_t done = head + vecs * 16;
      if (tid < size - done) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((uint32_t)data[done + tid] * 128u + base), "r"(inc) : "memory");
    }
    first = false;
    for (uint32_t e = 0; e < (uint32_t)kObsEpochVecs && i < vecs; e += kObsInflight, i += stride * kObsInflight) {
      uint4 q[kObsInflight];
#pragma unroll
      for (int u = 0; u < kObsInflight; u++)
        if (i + u * stride < vecs) q[u] = __ldg(v + i + u * stride);
#pragma unroll
      for (int u = 0; u < kObsInflight; u++)
        if (i + u * stride < vecs) obs_count16(base, inc, q[u]);
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
    __syncthreads();
  } while (__syncthreads_or(i < vecs)); // another epoch only for inputs beyond 64,512 bytes per thread
#pragma unroll
  for (int t = 0; t < 4; t++)
    if (total[t]) atomicAdd(hist + tid + 64 * t, total[t]);
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

// one WARP per segment: count the segment's bytes, normalise, write its 256 u16 counts
constexpr int kSegWarps = 4;
__global__ void __launch_bounds__(kSegWarps * 32) segments_kernel(const uint8_t *data, uint64_t size, uint64_t segmentBytes, int bits,
                                                                  uint16_t *symbolCounts)
{
  __shared__ uint32_t sHist[kSegWarps][256];
  __shared__ uint16_t sCapped[kSegWarps][256];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint64_t segments = (size + segmentBytes - 1) / segmentBytes;
  for (uint64_t seg = (uint64_t)blockIdx.x * kSegWarps + warp; seg < segments; seg += (uint64_t)gridDim.x * kSegWarps) {
    const uint64_t begin = seg * segmentBytes;
    const uint64_t end = begin + segmentBytes < size ? begin + segmentBytes : size;
    warp_observe(data, begin, end, sHist[warp], lane);
    warp_normalize(sHist[warp], end - begin, bits, sCapped[warp], sHist[warp], symbolCounts + seg * 256, lane);
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
  // persistent CTAs: six 33 KB CTAs (32 KB plane + the 1 KB every CTA reserves) fill an SM's shared memory
  static int configured[64] = {0};
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    if (cudaFuncSetAttribute(observe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kObsPlaneBytes) != cudaSuccess) return -2;
    configured[dev] = 1;
  }
  const uint64_t per = (uint64_t)kObsThreads * 16 * kObsInflight;
  uint64_t grid = (size + per - 1) / per;
  if (grid > (uint64_t)sms * 6) grid = (uint64_t)sms * 6;
  observe_kernel<<<(unsigned)grid, kObsThreads, kObsPlaneBytes, st>>>(static_cast<const uint8_t *>(dData), size, dHist);
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
  const uint64_t ctas = (segments + kSegWarps - 1) / kSegWarps;
  const unsigned grid = (unsigned)(ctas < (uint64_t)sms * 16 ? ctas : (uint64_t)sms * 16);
  segments_kernel<<<grid, kSegWarps * 32, 0, static_cast<cudaStream_t>(cudaStream)>>>(static_cast<const uint8_t *>(dData), size, segmentBytes,
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
