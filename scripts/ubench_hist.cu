// ubench_hist.cu — which byte-histogram organisation reaches the HBM roofline on B200? (design input for
// observe_kernel, src/hist.cpp:8-14; results in profiles/r2/ubench_hist.jsonl)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_hist.bin scripts/ubench_hist.cu hypersonic-rans_b200/csrc/hsr_synth.cpp
// Variants (all checked against a CPU count of the same bytes):
//   0  one u32 histogram per warp, shared-memory atomics on data-dependent banks (round 1's kernel)
//   1  lane-private columns, u32: word (bin, lane) at bin * 32 + lane — every lane owns bank `lane`, so a warp's 32
//      increments never share a bank whatever the bytes are. One warp per CTA, 32 KB.
//   2  lane-private columns, u16 halves: two warps share the words, warp w adds 1 << 16w. 32 KB per 2-warp CTA.
//   3  as 2 with four 16-byte loads in flight per lane
//   4  as 2 but plain LDS.U16 / add / STS.U16 on the thread's own half-word instead of an atomic
//   5  as 3, four warps per CTA on u8 quarter... not possible without overflow; instead: 4 warps, two u16 planes (64 KB)
//   6  match.any-aggregated increments into one histogram per warp
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

extern "C" int hsr_synth_zipf(uint8_t *out, size_t n, double s, uint64_t seed, size_t segmentBytes);

__device__ __forceinline__ void red_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// ---- variant 0
__global__ void __launch_bounds__(256) k0(const uint8_t *data, uint64_t size, uint32_t *hist)
{
  __shared__ uint32_t sPriv[8][256];
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  for (int k = tid; k < 8 * 256; k += 256) (&sPriv[0][0])[k] = 0;
  __syncthreads();
  uint32_t *mine = sPriv[warp];
  const uint4 *v = reinterpret_cast<const uint4 *>(data);
  const uint64_t vecs = size / 16;
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + tid; i < vecs; i += (uint64_t)gridDim.x * 256) {
    const uint4 q = __ldg(v + i);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      atomicAdd(mine + (w[j] & 0xffu), 1u); atomicAdd(mine + ((w[j] >> 8) & 0xffu), 1u);
      atomicAdd(mine + ((w[j] >> 16) & 0xffu), 1u); atomicAdd(mine + (w[j] >> 24), 1u);
    }
  }
  __syncthreads();
  for (int b = tid; b < 256; b += 256) {
    uint32_t s = 0;
    for (int w = 0; w < 8; w++) s += sPriv[w][b];
    if (s) atomicAdd(hist + b, s);
  }
}

// ---- lane-private columns. WARPS warps share one plane of 256 x 32 words; with WARPS == 2 the counters are u16
// halves, with WARPS == 1 full u32. INFLIGHT 16-byte loads per lane are requested before the first is counted.
template <int WARPS, int PLANES, int INFLIGHT, bool RMW>
__global__ void __launch_bounds__(WARPS *PLANES * 32) kcol(const uint8_t *data, uint64_t size, uint32_t *hist)
{
  extern __shared__ __align__(16) uint32_t s[];
  constexpr int T = WARPS * PLANES * 32;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t plane = warp / WARPS, half = warp % WARPS;
  for (int k = tid; k < PLANES * 8192; k += T) s[k] = 0;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(s) + plane * 32768u + lane * 4u + (RMW ? half * 2u : 0u);
  const uint32_t inc = WARPS == 2 ? (1u << (16 * half)) : 1u;
  const uint4 *v = reinterpret_cast<const uint4 *>(data);
  const uint64_t vecs = size / 16;
  const uint64_t stride = (uint64_t)gridDim.x * T;
  // u16 halves: at most 4032 vectors (64,512 bytes) per thread between flushes; this benchmark's sizes stay below that
  for (uint64_t i = (uint64_t)blockIdx.x * T + tid; i < vecs; i += stride * INFLIGHT) {
    uint4 q[INFLIGHT];
#pragma unroll
    for (int u = 0; u < INFLIGHT; u++)
      if (i + u * stride < vecs) q[u] = __ldg(v + i + u * stride);
#pragma unroll
    for (int u = 0; u < INFLIGHT; u++) {
      if (i + u * stride >= vecs) break;
      const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const uint32_t byte = __byte_perm(w[j], 0, 0x4440 + b);
          const uint32_t a = byte * 128u + base;
          if (RMW) {
            uint32_t c;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(c) : "r"(a) : "memory");
            c += 1;
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(c) : "memory");
          } else
            red_add(a, inc);
        }
      }
    }
  }
  __syncthreads();
  // bin b of plane p: 32 lane words, read with a rotation so the CTA's threads stay out of each other's banks
  for (int b = tid; b < 256; b += T) {
    uint32_t sum = 0;
    for (int p = 0; p < PLANES; p++)
      for (int j = 0; j < 32; j++) {
        const uint32_t w = s[p * 8192 + b * 32 + ((j + tid) & 31)];
        sum += WARPS == 2 ? (w & 0xffffu) + (w >> 16) : w;
      }
    if (sum) atomicAdd(hist + b, sum);
  }
}

// ---- variant 6: match.any aggregation
__global__ void __launch_bounds__(256) k6(const uint8_t *data, uint64_t size, uint32_t *hist)
{
  __shared__ uint32_t sPriv[8][256];
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int k = tid; k < 8 * 256; k += 256) (&sPriv[0][0])[k] = 0;
  __syncthreads();
  uint32_t *mine = sPriv[warp];
  const uint4 *v = reinterpret_cast<const uint4 *>(data);
  const uint64_t vecs = size / 16;
  const uint64_t vecsWarp = vecs & ~31ull; // full warps only (benchmark sizes are multiples of 512)
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + tid; i < vecsWarp; i += (uint64_t)gridDim.x * 256) {
    const uint4 q = __ldg(v + i);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const uint32_t byte = (w[j] >> (8 * b)) & 0xffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, byte);
        if ((peers & ((1u << lane) - 1)) == 0) mine[byte] += __popc(peers); // leader: lanes own distinct bins here
        __syncwarp();
      }
  }
  __syncthreads();
  for (int b = tid; b < 256; b += 256) {
    uint32_t s = 0;
    for (int w = 0; w < 8; w++) s += sPriv[w][b];
    if (s) atomicAdd(hist + b, s);
  }
}

struct Variant { const char *name; void (*launch)(const uint8_t *, uint64_t, uint32_t *, int sms); };

template <int WARPS, int PLANES, int INFLIGHT, bool RMW>
static void launch_col(const uint8_t *d, uint64_t n, uint32_t *h, int sms)
{
  const int smem = PLANES * 32768;
  cudaFuncSetAttribute(kcol<WARPS, PLANES, INFLIGHT, RMW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int perSm = (228 * 1024) / (smem + 1024);
  kcol<WARPS, PLANES, INFLIGHT, RMW><<<sms * perSm, WARPS * PLANES * 32, smem>>>(d, n, h);
}
static void launch0(const uint8_t *d, uint64_t n, uint32_t *h, int sms) { k0<<<sms * 8, 256>>>(d, n, h); }
static void launch6(const uint8_t *d, uint64_t n, uint32_t *h, int sms) { k6<<<sms * 8, 256>>>(d, n, h); }

int main(int argc, char **argv)
{
  const size_t n = argc > 1 ? (size_t)atoll(argv[1]) : 1000000000ull; // multiple of 512
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const Variant variants[] = {
      {"0 one u32 histogram per warp, atomics on data-dependent banks (round 1)", launch0},
      {"1 lane-private u32 columns, 1 warp per 32 KB CTA", launch_col<1, 1, 1, false>},
      {"1b lane-private u32 columns, 1 warp per 32 KB CTA, 4 loads in flight", launch_col<1, 1, 4, false>},
      {"2 lane-private u16 halves, 2 warps per 32 KB CTA", launch_col<2, 1, 1, false>},
      {"3 lane-private u16 halves, 2 warps per 32 KB CTA, 4 loads in flight", launch_col<2, 1, 4, false>},
      {"3b lane-private u16 halves, 2 warps per 32 KB CTA, 2 loads in flight", launch_col<2, 1, 2, false>},
      {"4 lane-private u16 halves, plain LDS/add/STS, 4 loads in flight", launch_col<2, 1, 4, true>},
      {"5 lane-private u16 halves, 4 warps per 64 KB CTA, 4 loads in flight", launch_col<2, 2, 4, false>},
      {"6 match.any-aggregated, one histogram per warp", launch6},
  };
  struct Input { const char *name; double s; size_t seg; } inputs[] = {{"zipf1_pw64k", 1.0, 65536}, {"zipf1_iid", 1.0, 0}, {"zipf3_iid", 3.0, 0}, {"uniform", 0.0, 0}};
  std::vector<uint8_t> host(n);
  uint8_t *d; uint32_t *h;
  cudaMalloc(&d, n); cudaMalloc(&h, 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (const Input &in : inputs) {
    hsr_synth_zipf(host.data(), n, in.s, 42, in.seg);
    uint32_t want[256] = {0};
    for (size_t i = 0; i < n; i++) want[host[i]]++;
    cudaMemcpy(d, host.data(), n, cudaMemcpyHostToDevice);
    for (const Variant &v : variants) {
      uint32_t got[256];
      cudaMemset(h, 0, 1024);
      v.launch(d, n, h, sms);
      cudaError_t err = cudaDeviceSynchronize();
      cudaMemcpy(got, h, 1024, cudaMemcpyDeviceToHost);
      const bool ok = err == cudaSuccess && memcmp(got, want, 1024) == 0;
      float best = 1e30f, sum = 0;
      const int reps = 10;
      for (int r = 0; r < reps; r++) {
        cudaMemsetAsync(h, 0, 1024);
        cudaEventRecord(e0);
        v.launch(d, n, h, sms);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms); sum += ms;
      }
      printf("{\"input\": \"%s\", \"variant\": \"%s\", \"bytes\": %zu, \"ms_mean\": %.4f, \"ms_best\": %.4f, \"GBps_mean\": %.1f, \"exact\": %s, \"err\": \"%s\"}\n",
             in.name, v.name, n, sum / reps, best, n / (sum / reps) / 1e6, ok ? "true" : "false", cudaGetErrorString(err));
      fflush(stdout);
    }
  }
  return 0;
}
