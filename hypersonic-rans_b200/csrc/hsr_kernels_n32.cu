// rANS32x32_16w family kernels (32 interleaved states: one per lane)
#define HSR_N 32
#include "hsr_kernels_inst.inl"
