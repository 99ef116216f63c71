// ubench_lsu.cu — which B200 pipes do LDS / SHFL / STG / LDGSTS share, and what does each cost? (design input for the
// decode loop; results in profiles/r1/ubench_lsu.jsonl)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_lsu.bin scripts/ubench_lsu.cu
// Persistent one-warp CTAs, 20 per SM like the decoder. Every op is an `asm volatile` with loop-invariant operands, so
// an iteration is nothing but the ops under test plus the loop counter. Reported: SM cycles per warp-iteration, i.e.
// the time the SM needs for ONE warp's 8 ops of each selected kind (20 warps keep every pipe busy).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 20000;
enum { LDS = 1, LDS2 = 2, LDS4 = 4, SHFL = 8, STG8 = 16, STG32 = 32, LDGSTS = 64, POPC = 128, STS8 = 256 };

#define REP8(x) x x x x x x x x

template <int MODE>
__global__ void __launch_bounds__(32, 20) k(uint32_t *out, uint8_t *gbuf, const uint8_t *gsrc)
{
  __shared__ __align__(16) uint32_t sm[2048];
  const uint32_t lane = threadIdx.x;
  for (int i = lane; i < 2048; i += 32) sm[i] = i;
  __syncwarp();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t a1 = sbase + lane * 4;                                  // conflict-free
  const uint32_t a2 = sbase + (lane & 15) * 4 + (lane >> 4) * 128;       // 2 words per bank
  const uint32_t a4 = sbase + (lane & 7) * 4 + (lane >> 3) * 128;        // 4 words per bank
  uint8_t *g = gbuf + (size_t)blockIdx.x * 4096 + lane;
  uint8_t *g4 = gbuf + (size_t)blockIdx.x * 4096 + lane * 4;
  const uint8_t *gs = gsrc + (size_t)blockIdx.x * 65536 + lane * 16;
  uint32_t v = lane, w = 0, s0 = lane * 3, s1 = lane * 5, s2 = lane * 7, s3 = lane * 11;
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
    if (MODE & LDS)  { REP8(asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(w) : "r"(a1));) }
    if (MODE & LDS2) { REP8(asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(w) : "r"(a2));) }
    if (MODE & LDS4) { REP8(asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(w) : "r"(a4));) }
    if (MODE & SHFL) {
      asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(s0) : "r"(lane ^ 5)); asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(s1) : "r"(lane ^ 5));
      asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(s2) : "r"(lane ^ 5)); asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(s3) : "r"(lane ^ 5));
      asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(s0) : "r"(lane ^ 9)); asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(s1) : "r"(lane ^ 9));
      asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(s2) : "r"(lane ^ 9)); asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(s3) : "r"(lane ^ 9));
    }
    if (MODE & STG8) {
      uint8_t *q = g + 256 * (it & 15);
      asm volatile("st.global.u8 [%0], %1;" ::"l"(q), "r"(v) : "memory"); asm volatile("st.global.u8 [%0+32], %1;" ::"l"(q), "r"(v) : "memory");
      asm volatile("st.global.u8 [%0+64], %1;" ::"l"(q), "r"(v) : "memory"); asm volatile("st.global.u8 [%0+96], %1;" ::"l"(q), "r"(v) : "memory");
      asm volatile("st.global.u8 [%0+128], %1;" ::"l"(q), "r"(v) : "memory"); asm volatile("st.global.u8 [%0+160], %1;" ::"l"(q), "r"(v) : "memory");
      asm volatile("st.global.u8 [%0+192], %1;" ::"l"(q), "r"(v) : "memory"); asm volatile("st.global.u8 [%0+224], %1;" ::"l"(q), "r"(v) : "memory");
    }
    if (MODE & STG32){ asm volatile("st.global.u32 [%0], %1;" ::"l"(g4 + 128 * (it & 15)), "r"(v) : "memory");
                       asm volatile("st.global.u32 [%0], %1;" ::"l"(g4 + 128 * (it & 15) + 2048), "r"(v) : "memory"); }
    if (MODE & STS8) { REP8(asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(a1 + 1024 + 0), "r"(v) : "memory");) }
    if (MODE & POPC) {
      asm volatile("popc.b32 %0, %0;" : "+r"(s0)); asm volatile("popc.b32 %0, %0;" : "+r"(s1)); asm volatile("popc.b32 %0, %0;" : "+r"(s2)); asm volatile("popc.b32 %0, %0;" : "+r"(s3));
      asm volatile("popc.b32 %0, %0;" : "+r"(s0)); asm volatile("popc.b32 %0, %0;" : "+r"(s1)); asm volatile("popc.b32 %0, %0;" : "+r"(s2)); asm volatile("popc.b32 %0, %0;" : "+r"(s3));
    }
    if (MODE & LDGSTS) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.commit_group;\n\tcp.async.wait_group 1;" ::"r"(sbase + 4096 + lane * 16 + (it & 1) * 512), "l"(gs + (it & 63) * 512) : "memory");
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  out[blockIdx.x * 32 + lane] = w + v + s0 + s1 + s2 + s3;
}

static int g_sms = 0;
static double g_ghz = 0;
static uint32_t *g_out; static uint8_t *g_buf, *g_src;

template <int MODE>
static void run(const char *name)
{
  const int grid = g_sms * 20;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 32>>>(g_out, g_buf, g_src);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<grid, 32>>>(g_out, g_buf, g_src);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * g_ghz * 1e9 / ((double)ITER * 20.0);
  printf("{\"kernel\": \"%s\", \"ms\": %.3f, \"sm_cycles_per_warp_iteration\": %.2f, \"err\": \"%s\"}\n", name, ms, cyc, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  g_sms = p.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  g_ghz = khz / 1e6;
  cudaMalloc(&g_out, (size_t)g_sms * 20 * 32 * 4);
  cudaMalloc(&g_buf, (size_t)g_sms * 20 * 4096 + 4096);
  cudaMalloc(&g_src, (size_t)g_sms * 20 * 65536);
  cudaMemset(g_src, 1, (size_t)g_sms * 20 * 65536);
  printf("{\"sms\": %d, \"clock_ghz\": %.3f, \"note\": \"8 ops of each kind per iteration (stg32: 2 x 128 B, ldgsts: one 512-byte segment)\"}\n", g_sms, g_ghz);
  run<LDS>("lds x8 (conflict-free)");
  run<LDS2>("lds x8 (2-way conflict)");
  run<LDS4>("lds x8 (4-way conflict)");
  run<SHFL>("shfl x8");
  run<LDS | SHFL>("lds x8 + shfl x8");
  run<POPC>("popc x8");
  run<LDS | POPC>("lds x8 + popc x8");
  run<SHFL | POPC>("shfl x8 + popc x8");
  run<STG8>("stg.u8 x8 (32 B each)");
  run<LDS | STG8>("lds x8 + stg.u8 x8");
  run<LDS2 | STG8>("lds2way x8 + stg.u8 x8");
  run<STG32>("stg.u32 x2 (128 B each)");
  run<LDS2 | STG32>("lds2way x8 + stg.u32 x2");
  run<STS8>("sts.u8 x8");
  run<LDS | STS8>("lds x8 + sts.u8 x8");
  run<LDGSTS>("ldgsts 512 B");
  run<LDS2 | LDGSTS>("lds2way x8 + ldgsts 512 B");
  run<LDS4 | LDGSTS>("lds4way x8 + ldgsts 512 B");
  run<LDS2 | SHFL | STG8 | LDGSTS>("lds2way x8 + shfl x8 + stg.u8 x8 + ldgsts");
  return 0;
}
