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
#include "hsr_hist_device.cuh"

namespace hsr {

__global__ void __launch_bounds__(kHistThreads) observe_kernel(const uint8_t *data, uint64_t size, uint32_t *hist)
{
  __shared__ uint32_t sPriv[kHistWarps][256];
  __shared__ uint32_t sOut[256];
  cta_observe(data, 0, size, sPriv, sOut, blockIdx.x, gridDim.x);
  for (int b = threadIdx.x; b < 256; b += kHistThreads)
    if (sOut[b]) atomicAdd(hist + b, sOut[b]);
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
