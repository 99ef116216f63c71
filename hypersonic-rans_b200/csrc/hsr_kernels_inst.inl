// hsr_kernels_inst.inl — instantiates every (bits, table) kernel for one state count HSR_N.
// Included by hsr_kernels_n32.cu and hsr_kernels_n64.cu so the two compile in parallel.
#include "hsr_kernels.cuh"

namespace hsr {

#define HSR_CAT2(a, b) a##b
#define HSR_CAT(a, b) HSR_CAT2(a, b)
// e.g. units_n64_b15_t1: mt_/raw units kernel, 64 states, 15 bits, table kind 1 (bitmap-rank)
#define HSR_NAME(prefix, BITS, TK) HSR_CAT(prefix, HSR_CAT(HSR_N, HSR_CAT(_b, HSR_CAT(BITS, HSR_CAT(_t, TK)))))

// 64 registers keep 32 one-warp CTAs per SM resident (the hardware limit); the decode loop needs ~50.
#ifndef HSR_UNITS_MIN_CTAS
#define HSR_UNITS_MIN_CTAS 32
#endif
#define HSR_DEFINE(BITS, TK)                                                                                          \
  __global__ void __launch_bounds__(32, HSR_UNITS_MIN_CTAS) HSR_NAME(units_n, BITS, TK)(DecodeParams p)                               \
  { units_kernel_body<BITS, HSR_N, TK>(p); }                                                                          \
  __global__ void __launch_bounds__(32, 16) HSR_NAME(block_n, BITS, TK)(BlockStreamParams p)                           \
  { block_kernel_body<BITS, HSR_N, TK>(p); }

#define HSR_ENTRY(BITS, TK)                                                                                           \
  { (const void *)HSR_NAME(units_n, BITS, TK), (const void *)HSR_NAME(block_n, BITS, TK), WarpLayout<BITS, HSR_N, TK>::kBytes,         \
    WarpLayout<BITS, HSR_N, TK>::kDynamic ? 1 : 0 }
#define HSR_NONE { nullptr, nullptr, 0, 0 }

HSR_DEFINE(10, 1) HSR_DEFINE(11, 1) HSR_DEFINE(12, 1) HSR_DEFINE(13, 1) HSR_DEFINE(14, 1) HSR_DEFINE(15, 1)
HSR_DEFINE(10, 2) HSR_DEFINE(11, 2) HSR_DEFINE(12, 2)
HSR_DEFINE(13, 3) HSR_DEFINE(14, 3) HSR_DEFINE(15, 3)

extern const KernelEntry HSR_CAT(kKernels, HSR_N)[6][3] = {
  { HSR_ENTRY(10, 1), HSR_ENTRY(10, 2), HSR_NONE }, { HSR_ENTRY(11, 1), HSR_ENTRY(11, 2), HSR_NONE }, { HSR_ENTRY(12, 1), HSR_ENTRY(12, 2), HSR_NONE },
  { HSR_ENTRY(13, 1), HSR_NONE, HSR_ENTRY(13, 3) }, { HSR_ENTRY(14, 1), HSR_NONE, HSR_ENTRY(14, 3) }, { HSR_ENTRY(15, 1), HSR_NONE, HSR_ENTRY(15, 3) },
};

} // namespace hsr
