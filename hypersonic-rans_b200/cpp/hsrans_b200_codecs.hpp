// hsrans_b200_codecs.hpp — C++ host-side mirror of the reference's decoder entry points, over the C-ABI.
//
// The reference exports one free function per (codec, implementation, probability bits) with the signature
//     size_t f(const uint8_t *pInData, const size_t inLength, uint8_t *pOutData, const size_t outCapacity);
// (codec_info_t::decodeFunc, src/main.cpp:149) and lists them in `_Codecs[].decoders[]` (src/main.cpp:172-236).
// This header provides functions with the SAME type and the reference's naming scheme, prefixed `cuda_`, so a
// maintainer can append them to those tables:
//     cuda_rANS32x32_16w_decode_<b>          replaces rANS32x32_16w_decode_scalar_<b>   (src/rANS32x32_16w.h:9 ...)
//     cuda_rANS32x64_16w_decode_<b>          replaces rANS32x64_16w_decode_scalar_<b>   (src/rANS32x64_16w.h:9 ...)
//     cuda_block_rANS32x32_16w_decode_<b>    replaces block_rANS32x32_16w_decode_<b>    (src/block_rANS32x32_16w.h:15-20)
//     cuda_block_rANS32x64_16w_decode_<b>    replaces block_rANS32x64_16w_decode_<b>    (src/block_rANS32x64_16w.h:15-20)
//     cuda_mt_rANS32x32_16w_decode_<b>       replaces mt_rANS32x32_16w_decode_<b> and _decode_mt_<b> (src/mt_rANS32x32_16w.h:16-28)
//     cuda_mt_rANS32x64_16w_decode_<b>       replaces mt_rANS32x64_16w_decode_<b> and _decode_mt_<b> (src/mt_rANS32x64_16w.h:16-28)
//     cuda_rANS32x16_16w_decode_<b>          replaces rANS32x16_16w_decode_scalar_<b>   (src/rANS32x16_16w.h:9 ...)
//     cuda_rANS32x32_32blk_16w_decode_<b>    replaces rANS32x32_32blk_16w_decode_scalar_<b> (src/rans32x32_32blk_16w.h:9 ...)
// for <b> in 10..15, plus the `*_capacity` twins. Same return convention: decoded size, or 0 on any error.
// The mt_ thread-pool entry points have twins with the pool signature as well,
//     cuda_mt_rANS32xNN_16w_decode_mt_<b>(pInData, inLength, pOutData, outCapacity, thread_pool *)
// (src/mt_rANS32x64_16w.h:23-28), so that `decode_with_thread_pool_wrapper<>` (src/main.cpp:163-170) instantiates on
// them and the "decode (multi threaded)" row can be SWAPPED, not only appended to. The pool argument is ignored: the
// GPU is the pool (hsr_decode_mt_multi spreads one stream over several GPUs the way decode_mt spreads it over
// threads, src/mt_rANS32x64_16w_decode.cpp:137-265).
#ifndef HSRANS_B200_CODECS_HPP
#define HSRANS_B200_CODECS_HPP

#include <stddef.h>
#include <stdint.h>

#include "../../include/hsrans_b200.h"

#define HSR_B200_DECODER(name, family, states, bits)                                                                     \
  inline size_t name##_##bits(const uint8_t *pInData, const size_t inLength, uint8_t *pOutData, const size_t outCapacity) \
  {                                                                                                                      \
    return hsr_decode(family, states, bits, pInData, inLength, pOutData, outCapacity);                                   \
  }

#define HSR_B200_DECODERS(name, family, states)                                                                          \
  HSR_B200_DECODER(name, family, states, 15)                                                                             \
  HSR_B200_DECODER(name, family, states, 14)                                                                             \
  HSR_B200_DECODER(name, family, states, 13)                                                                             \
  HSR_B200_DECODER(name, family, states, 12)                                                                             \
  HSR_B200_DECODER(name, family, states, 11)                                                                             \
  HSR_B200_DECODER(name, family, states, 10)

HSR_B200_DECODERS(cuda_rANS32x32_16w_decode, HSR_RAW, 32)
HSR_B200_DECODERS(cuda_rANS32x64_16w_decode, HSR_RAW, 64)
HSR_B200_DECODERS(cuda_block_rANS32x32_16w_decode, HSR_BLOCK, 32)
HSR_B200_DECODERS(cuda_block_rANS32x64_16w_decode, HSR_BLOCK, 64)
HSR_B200_DECODERS(cuda_mt_rANS32x32_16w_decode, HSR_MT, 32)
HSR_B200_DECODERS(cuda_mt_rANS32x64_16w_decode, HSR_MT, 64)
HSR_B200_DECODERS(cuda_rANS32x16_16w_decode, HSR_RAW, 16)
HSR_B200_DECODERS(cuda_rANS32x32_32blk_16w_decode, HSR_RAW32BLK, 32)

struct thread_pool; // src/thread_pool.h:7 (opaque here as there)

#define HSR_B200_POOL_DECODER(name, states, bits)                                                                        \
  inline size_t name##_##bits(const uint8_t *pInData, const size_t inLength, uint8_t *pOutData, const size_t outCapacity, \
                              thread_pool * /* the GPU is the pool */)                                                   \
  {                                                                                                                      \
    return hsr_decode(HSR_MT, states, bits, pInData, inLength, pOutData, outCapacity);                                   \
  }
#define HSR_B200_POOL_DECODERS(name, states)                                                                             \
  HSR_B200_POOL_DECODER(name, states, 15) HSR_B200_POOL_DECODER(name, states, 14) HSR_B200_POOL_DECODER(name, states, 13) \
  HSR_B200_POOL_DECODER(name, states, 12) HSR_B200_POOL_DECODER(name, states, 11) HSR_B200_POOL_DECODER(name, states, 10)
HSR_B200_POOL_DECODERS(cuda_mt_rANS32x32_16w_decode_mt, 32)
HSR_B200_POOL_DECODERS(cuda_mt_rANS32x64_16w_decode_mt, 64)
#undef HSR_B200_POOL_DECODERS
#undef HSR_B200_POOL_DECODER

#undef HSR_B200_DECODERS
#undef HSR_B200_DECODER

// src/rANS32x32_16w.cpp:10-13, src/block_rANS32x32_16w_encode.cpp:47-54, src/mt_rANS32x64_16w_encode.cpp:50-57
inline size_t cuda_rANS32x32_16w_capacity(const size_t inputSize) { return hsr_capacity(HSR_RAW, 32, inputSize); }
inline size_t cuda_rANS32x64_16w_capacity(const size_t inputSize) { return hsr_capacity(HSR_RAW, 64, inputSize); }
inline size_t cuda_block_rANS32x32_16w_capacity(const size_t inputSize) { return hsr_capacity(HSR_BLOCK, 32, inputSize); }
inline size_t cuda_block_rANS32x64_16w_capacity(const size_t inputSize) { return hsr_capacity(HSR_BLOCK, 64, inputSize); }
inline size_t cuda_mt_rANS32x32_16w_capacity(const size_t inputSize) { return hsr_capacity(HSR_MT, 32, inputSize); }
inline size_t cuda_mt_rANS32x64_16w_capacity(const size_t inputSize) { return hsr_capacity(HSR_MT, 64, inputSize); }
// src/rANS32x16_16w.cpp:10-13, src/rans32x32_32blk_16w.cpp:10-13
inline size_t cuda_rANS32x16_16w_capacity(const size_t inputSize) { return hsr_capacity(HSR_RAW, 16, inputSize); }
inline size_t cuda_rANS32x32_32blk_16w_capacity(const size_t inputSize) { return hsr_capacity(HSR_RAW32BLK, 32, inputSize); }

// Device producers with the reference's mt_ encoder signature (src/mt_rANS32x64_16w.h:9-14):
//     size_t mt_rANS32x64_16w_encode_<b>(const uint8_t *pInData, const size_t length, uint8_t *pOutData, const size_t outCapacity)
// Same stream format, fixed 64 KiB blocks (see hsr_encode_mt); decodable by every reference mt_ decoder.
#define HSR_B200_ENCODER(states, bits)                                                                                           \
  inline size_t cuda_mt_rANS32x##states##_16w_encode_##bits(const uint8_t *pInData, const size_t length, uint8_t *pOutData,     \
                                                            const size_t outCapacity)                                           \
  {                                                                                                                              \
    return hsr_encode_mt(states, bits, pInData, length, pOutData, outCapacity, 0);                                               \
  }
HSR_B200_ENCODER(32, 15) HSR_B200_ENCODER(32, 14) HSR_B200_ENCODER(32, 13) HSR_B200_ENCODER(32, 12) HSR_B200_ENCODER(32, 11) HSR_B200_ENCODER(32, 10)
HSR_B200_ENCODER(64, 15) HSR_B200_ENCODER(64, 14) HSR_B200_ENCODER(64, 13) HSR_B200_ENCODER(64, 12) HSR_B200_ENCODER(64, 11) HSR_B200_ENCODER(64, 10)
#undef HSR_B200_ENCODER

// hist.h:54-58 twins on the device (bit-exact): make_hist = observe_hist + normalize_hist.
struct cuda_hist_t { // same layout as hist_t, src/hist.h:6-10
  uint16_t symbolCount[256];
  uint16_t cumul[256];
};
inline bool cuda_make_hist(cuda_hist_t *pHist, const uint8_t *pData, const size_t size, const size_t totalSymbolCountBits)
{
  return hsr_make_hist(pData, size, (int)totalSymbolCountBits, pHist->symbolCount, pHist->cumul) == 0;
}

#endif // HSRANS_B200_CODECS_HPP
