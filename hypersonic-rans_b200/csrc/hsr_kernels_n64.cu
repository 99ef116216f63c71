// rANS32x64_16w family kernels (64 interleaved states: two per lane)
#define HSR_N 64
#include "hsr_kernels_inst.inl"
