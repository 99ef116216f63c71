// hsr_hist.cu — device histogram count + normalise, bit-exact with the reference's make_hist
// (src/hist.cpp:217-222 = observe_hist :8-14 + normalize_hist :16-215).
//
// observe: lane-private counter columns in shared memory (conflict-free shared-memory atomics whatever the bytes are),
// four 16-byte coalesced loads in flight per lane, one global atomicAdd per bin per CTA at the end.
// normalise: 256 elements of order-dependent integer logic. The float scale is done with explicit round-to-
// nearest MUL and ADD (never an FMA — the reference build has no FMA target, src/hist.cpp:60-64); the index
// heap-sort and the steal/charity loops (:105-198) run literally, one thread per histogram, because the tie order of
// that particular unstable sort decides which symbols receive the +-1 (one histogram: thread 0 of a CTA; the
// per-block histograms of the block_/mt_ encoders: one LANE each, 32 histograms per warp).
#include <atomic>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/hsrans_b200.h"
#include "hsr_hist_device.cuh"

namespace hsr {

// observe_hist (src/hist.cpp:8-14) at streaming speed: see cta_count (hsr_hist_device.cuh)
__global__ void __launch_bounds__(kCntThreads) observe_kernel(const uint8_t *data, uint64_t size, uint32_t *hist)
{
  extern __shared__ __align__(16) uint32_t sPlane[];
  uint32_t total[4];
  cta_count(data, size, blockIdx.x, gridDim.x, sPlane, total);
#pragma unroll
  for (int t = 0; t < 4; t++)
    if (total[t]) atomicAdd(hist + threadIdx.x + 64 * t, total[t]);
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

// per-range histograms in two steps: raw counts by one 64-thread CTA per range (conflict-free columns), then the
// order-dependent normalisation with one LANE per range (hsr_hist_device.cuh). Round 1 gave a warp to each range and
// parked 31 lanes behind lane 0's heap-sort: 1.19 ms per GB of 64 KiB segments, most of it one-lane shared-memory
// requests that occupy the SM's data pipe like full ones.
__global__ void __launch_bounds__(kCntThreads) seg_count_kernel(const uint8_t *data, SegPlan pl, uint32_t *counts32)
{
  extern __shared__ __align__(16) uint32_t sPlane[];
  cta_count_ranges(data, pl, counts32, sPlane);
}

__global__ void __launch_bounds__(32) seg_normalize_kernel(const uint32_t *counts32, SegPlan pl, int bits, uint16_t *symbolCounts)
{
  extern __shared__ __align__(16) uint8_t sNorm[];
  warp_normalize_ranges(counts32, pl, bits, symbolCounts, sNorm);
}

// host-side launchers, shared with the device encoder (hsr_encode.cu)
static bool configure_range_kernels(int *smsOut)
{
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  *smsOut = sms;
  static std::atomic<bool> configured[64]; // zero-initialised; setting the attribute twice is harmless
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    if (cudaFuncSetAttribute(seg_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCntPlaneBytes) != cudaSuccess ||
        cudaFuncSetAttribute(seg_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLaneNormSmemBytes) != cudaSuccess)
      return false;
    configured[dev] = true;
  }
  return true;
}

bool launch_range_counts(const uint8_t *dData, const SegPlan &pl, uint32_t *dCounts32, cudaStream_t st)
{
  int sms;
  if (!configure_range_kernels(&sms)) return false;
  const unsigned gridC = (unsigned)(pl.num < (uint64_t)sms * 6 ? pl.num : (uint64_t)sms * 6);
  seg_count_kernel<<<gridC, kCntThreads, kCntPlaneBytes, st>>>(dData, pl, dCounts32);
  return cudaGetLastError() == cudaSuccess;
}

bool launch_range_histograms(const uint8_t *dData, const SegPlan &pl, int bits, uint32_t *dCounts32, uint16_t *dCounts, cudaStream_t st)
{
  int sms;
  if (!configure_range_kernels(&sms) || !launch_range_counts(dData, pl, dCounts32, st)) return false;
  const uint64_t warps = ((uint64_t)pl.num + 31) / 32;
  const unsigned gridN = (unsigned)(warps < (uint64_t)sms * 4 ? warps : (uint64_t)sms * 4);
  seg_normalize_kernel<<<gridN, 32, kLaneNormSmemBytes, st>>>(dCounts32, pl, bits, dCounts);
  return cudaGetLastError() == cudaSuccess;
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
  static std::atomic<bool> configured[64];
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    if (cudaFuncSetAttribute(observe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCntPlaneBytes) != cudaSuccess) return -2;
    configured[dev] = true;
  }
  const uint64_t per = (uint64_t)kCntThreads * 16 * kCntInflight;
  uint64_t grid = (size + per - 1) / per;
  if (grid > (uint64_t)sms * 6) grid = (uint64_t)sms * 6;
  observe_kernel<<<(unsigned)grid, kCntThreads, kCntPlaneBytes, st>>>(static_cast<const uint8_t *>(dData), size, dHist);
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
  const uint64_t segments = (size + segmentBytes - 1) / segmentBytes;
  if (segments > 0x7fffffffull) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(cudaStream);
  uint32_t *dCounts32 = nullptr; // stream-ordered scratch: the raw counts between the two kernels
  if (cudaMallocAsync(&dCounts32, segments * 1024, st) != cudaSuccess) {
    (void)cudaGetLastError();
    return -2;
  }
  SegPlan pl{nullptr, nullptr, (uint64_t)segmentBytes, (uint64_t)size, (uint32_t)segments};
  const bool ok = launch_range_histograms(static_cast<const uint8_t *>(dData), pl, bits, dCounts32, dSymbolCounts, st);
  cudaFreeAsync(dCounts32, st);
  return ok ? 1 : -2;
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
