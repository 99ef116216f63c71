// ref_shim.cpp — extern "C" doorway onto the UNMODIFIED reference, compiled where it lies.
//
// TEST INFRASTRUCTURE ONLY (see hsrans_oracle.h). This file contains no reference code: it includes the
// reference's public headers from /root/reference/src (passed with -I by oracle/Makefile) and forwards to the
// reference's own per-bits entry points, so Python (ctypes) and bench.py can use the reference as
//   (1) the stream producer (its scalar encoders: src/rANS32x32_16w.cpp:34, src/block_rANS32x32_16w_encode.cpp:137,
//       src/mt_rANS32x64_16w_encode.cpp:140),
//   (2) the byte-exact checker (its decoders), and
//   (3) the timed CPU baseline (fastest AVX2 decoders + the mt_ thread pool, src/main.cpp:163-214).
// The output of this build lives only in oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "hist.h"
#include "simd_platform.h"
#include "thread_pool.h"
#include "rANS32x16_16w.h"
#include "rANS32x32_16w.h"
#include "rANS32x64_16w.h"
#include "rans32x32_32blk_16w.h"
#include "block_rANS32x32_16w.h"
#include "block_rANS32x64_16w.h"
#include "mt_rANS32x32_16w.h"
#include "mt_rANS32x64_16w.h"

typedef size_t (*dec_fn)(const uint8_t *, const size_t, uint8_t *, const size_t);
typedef size_t (*dec_pool_fn)(const uint8_t *, const size_t, uint8_t *, const size_t, thread_pool *);
typedef size_t (*enc_fn)(const uint8_t *, const size_t, uint8_t *, const size_t);
typedef size_t (*enc_hist_fn)(const uint8_t *, const size_t, uint8_t *, const size_t, const hist_t *);

#define BITS6(prefix) { prefix##_10, prefix##_11, prefix##_12, prefix##_13, prefix##_14, prefix##_15 }
#define MIXED6(lo, hi) { lo##_10, lo##_11, lo##_12, hi##_13, hi##_14, hi##_15 }

static const enc_hist_fn raw_enc[2][6] = { BITS6(rANS32x32_16w_encode_scalar), BITS6(rANS32x64_16w_encode_scalar) };
static const dec_fn raw_dec_scalar[2][6] = { BITS6(rANS32x32_16w_decode_scalar), BITS6(rANS32x64_16w_decode_scalar) };
// the `candidateForFastest` AVX2 rows of src/main.cpp:202-214: xmmShfl2 varC for bits <= 12, xmmShfl2 varA above
static const dec_fn raw_dec_avx2[2][6] = {
  MIXED6(rANS32x32_xmmShfl2_16w_decode_avx2_varC, rANS32x32_xmmShfl2_16w_decode_avx2_varA),
  MIXED6(rANS32x64_xmmShfl2_16w_decode_avx2_varC, rANS32x64_xmmShfl2_16w_decode_avx2_varA) };
// (the 32-state codec has no AVX-512 decoder above 12 bits)
static const dec_fn raw_dec_avx512[2][6] = {
  { rANS32x32_xmmShfl2_16w_decode_avx512_varC_10, rANS32x32_xmmShfl2_16w_decode_avx512_varC_11, rANS32x32_xmmShfl2_16w_decode_avx512_varC_12, nullptr, nullptr, nullptr },
  MIXED6(rANS32x64_xmmShfl2_16w_decode_avx512_varC, rANS32x64_xmmShfl2_16w_decode_avx512_varA) };
// 16-state codec and the 32blk layout: `candidateForFastest` AVX2 rows of src/main.cpp:216-228 (xmmShfl varC /
// varC2 for bits <= 12, xmmShfl varA / varA2 above)
static const enc_hist_fn raw16_enc[6] = BITS6(rANS32x16_16w_encode_scalar);
static const dec_fn raw16_dec_scalar[6] = BITS6(rANS32x16_16w_decode_scalar);
static const dec_fn raw16_dec_avx2[6] = MIXED6(rANS32x16_xmmShfl_16w_decode_avx2_varC, rANS32x16_xmmShfl_16w_decode_avx2_varA);
static const enc_hist_fn blk32_enc[6] = BITS6(rANS32x32_32blk_16w_encode_scalar);
static const dec_fn blk32_dec_scalar[6] = BITS6(rANS32x32_32blk_16w_decode_scalar);
static const dec_fn blk32_dec_avx2[6] = MIXED6(rANS32x32_32blk_16w_decode_avx2_varC2, rANS32x32_32blk_16w_decode_avx2_varA2);
static const enc_fn block_enc[2][6] = { BITS6(block_rANS32x32_16w_encode), BITS6(block_rANS32x64_16w_encode) };
static const dec_fn block_dec[2][6] = { BITS6(block_rANS32x32_16w_decode), BITS6(block_rANS32x64_16w_decode) };
static const enc_fn mt_enc[2][6] = { BITS6(mt_rANS32x32_16w_encode), BITS6(mt_rANS32x64_16w_encode) };
static const dec_fn mt_dec[2][6] = { BITS6(mt_rANS32x32_16w_decode), BITS6(mt_rANS32x64_16w_decode) };
static const dec_pool_fn mt_dec_pool[2][6] = { BITS6(mt_rANS32x32_16w_decode_mt), BITS6(mt_rANS32x64_16w_decode_mt) };

static thread_pool *g_pool = nullptr;
static size_t g_pool_threads = 0;

static bool ok(int family, int N, int bits)
{
  if (bits < 10 || bits > 15) return false;
  if (family == 0) return N == 16 || N == 32 || N == 64;
  if (family == 3) return N == 32;
  return (family == 1 || family == 2) && (N == 32 || N == 64);
}

extern "C" {

enum { HSREF_RAW = 0, HSREF_BLOCK = 1, HSREF_MT = 2, HSREF_RAW32BLK = 3 };
// impl: 0 = scalar (raw) / CPU-dispatching decoder (block_, mt_ single thread)
//       1 = fastest AVX2 raw decoder / same dispatching decoder
//       2 = AVX-512 raw decoder (xmmShfl2) / same dispatching decoder
//       3 = mt_ only: thread-pool decoder (src/mt_rANS32x64_16w_decode.cpp:137-265)
enum { HSREF_IMPL_SCALAR = 0, HSREF_IMPL_AVX2 = 1, HSREF_IMPL_AVX512 = 2, HSREF_IMPL_POOL = 3 };

size_t hsref_capacity(int family, int N, size_t n)
{
  if (!ok(family, N, 10)) return 0;
  if (family == HSREF_RAW) return N == 16 ? rANS32x16_16w_capacity(n) : N == 32 ? rANS32x32_16w_capacity(n) : rANS32x64_16w_capacity(n);
  if (family == HSREF_RAW32BLK) return rANS32x32_32blk_16w_capacity(n);
  if (family == HSREF_BLOCK) return N == 32 ? block_rANS32x32_16w_capacity(n) : block_rANS32x64_16w_capacity(n);
  return N == 32 ? mt_rANS32x32_16w_capacity(n) : mt_rANS32x64_16w_capacity(n);
}

// raw: the harness builds the whole-buffer histogram first (src/main.cpp:746) and hands it to the encoder.
size_t hsref_encode(int family, int N, int bits, const uint8_t *in, size_t n, uint8_t *out, size_t cap)
{
  if (!ok(family, N, bits)) return 0;
  const int s = N == 64, b = bits - 10;
  if (family == HSREF_RAW || family == HSREF_RAW32BLK) {
    hist_t hist;
    make_hist(&hist, in, n, (size_t)bits);
    if (family == HSREF_RAW32BLK) return blk32_enc[b](in, n, out, cap, &hist);
    return N == 16 ? raw16_enc[b](in, n, out, cap, &hist) : raw_enc[s][b](in, n, out, cap, &hist);
  }
  return family == HSREF_BLOCK ? block_enc[s][b](in, n, out, cap) : mt_enc[s][b](in, n, out, cap);
}

size_t hsref_encode_raw_with_hist(int N, int bits, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                                  const uint16_t symbolCount[256])
{
  if (!ok(0, N, bits)) return 0;
  hist_t hist;
  memcpy(hist.symbolCount, symbolCount, sizeof(hist.symbolCount));
  if (!inplace_complete_hist(&hist, (size_t)bits)) return 0;
  return N == 16 ? raw16_enc[bits - 10](in, n, out, cap, &hist) : raw_enc[N == 64][bits - 10](in, n, out, cap, &hist);
}

int hsref_pool_threads(void) { return (int)g_pool_threads; }

void hsref_pool_destroy(void);

int hsref_pool_create(int threads) // threads <= 0: hardware_concurrency() - 1 like src/main.cpp:167
{
  hsref_pool_destroy();
  size_t t = threads > 0 ? (size_t)threads : thread_pool_max_threads();
  if (threads <= 0) t = t > 1 ? t - 1 : 1;
  g_pool = thread_pool_new(t);
  g_pool_threads = g_pool ? t : 0;
  return (int)g_pool_threads;
}

void hsref_pool_destroy(void)
{
  if (g_pool) thread_pool_destroy(&g_pool); // deletes the pool but leaves the caller's pointer as it was (src/thread_pool.cpp:108-114)
  g_pool = nullptr;
  g_pool_threads = 0;
}

size_t hsref_decode(int family, int N, int bits, int impl, const uint8_t *in, size_t inLen, uint8_t *out, size_t cap)
{
  if (!ok(family, N, bits)) return 0;
  const int s = N == 64, b = bits - 10;
  if (family == HSREF_RAW32BLK || (family == HSREF_RAW && N == 16)) { // no AVX-512 decoders exist for these two
    _DetectCPUFeatures();
    const dec_fn *scalar = family == HSREF_RAW32BLK ? blk32_dec_scalar : raw16_dec_scalar;
    const dec_fn *avx2 = family == HSREF_RAW32BLK ? blk32_dec_avx2 : raw16_dec_avx2;
    if (impl == HSREF_IMPL_AVX2) return avx2Supported ? avx2[b](in, inLen, out, cap) : 0;
    if (impl == HSREF_IMPL_AVX512) return 0;
    return scalar[b](in, inLen, out, cap);
  }
  if (family == HSREF_RAW) {
    _DetectCPUFeatures();
    if (impl == HSREF_IMPL_AVX2) return avx2Supported ? raw_dec_avx2[s][b](in, inLen, out, cap) : 0;
    if (impl == HSREF_IMPL_AVX512) return (avx512FSupported && avx512BWSupported && avx512DQSupported && raw_dec_avx512[s][b]) ? raw_dec_avx512[s][b](in, inLen, out, cap) : 0;
    return raw_dec_scalar[s][b](in, inLen, out, cap);
  }
  if (family == HSREF_BLOCK) return block_dec[s][b](in, inLen, out, cap);
  if (impl == HSREF_IMPL_POOL) {
    if (!g_pool) hsref_pool_create(0);
    return mt_dec_pool[s][b](in, inLen, out, cap, g_pool);
  }
  return mt_dec[s][b](in, inLen, out, cap);
}

// --max-simd semantics of src/main.cpp:463-618: level 0 = whatever cpuid says, 1 = cap at AVX2
// (clears the avx512* globals so the block_/mt_ dispatchers take their AVX2 sections), 2 = scalar only.
void hsref_set_max_simd(int level)
{
  _CpuFeaturesDetected = false;
  _DetectCPUFeatures();
  if (level >= 1) {
    avx512FSupported = avx512PFSupported = avx512ERSupported = avx512CDSupported = avx512BWSupported = false;
    avx512DQSupported = avx512VLSupported = avx512IFMASupported = avx512VBMISupported = avx512VNNISupported = false;
    avx512VBMI2Supported = avx512POPCNTDQSupported = avx512BITALGSupported = avx5124VNNIWSupported = avx5124FMAPSSupported = false;
  }
  if (level >= 2) {
    avx2Supported = avxSupported = fma3Supported = false;
    sse42Supported = sse41Supported = ssse3Supported = sse3Supported = false;
  }
}

int hsref_has_avx2(void) { _DetectCPUFeatures(); return avx2Supported; }
int hsref_has_avx512(void) { _DetectCPUFeatures(); return avx512FSupported && avx512BWSupported && avx512DQSupported; }
const char *hsref_cpu_name(void) { _DetectCPUFeatures(); return _CpuName; }

// src/hist.cpp:217-222 and :16-215, for pinning the oracle's and the device kernel's normalisation
void hsref_make_hist(const uint8_t *data, size_t size, int bits, uint16_t symbolCount[256], uint16_t cumul[256])
{
  hist_t h;
  make_hist(&h, data, size, (size_t)bits);
  memcpy(symbolCount, h.symbolCount, sizeof(h.symbolCount));
  memcpy(cumul, h.cumul, sizeof(h.cumul));
}

void hsref_normalize_hist(const uint32_t hist[256], size_t dataBytes, int bits, uint16_t symbolCount[256], uint16_t cumul[256])
{
  hist_t h;
  normalize_hist(&h, hist, dataBytes, (size_t)bits);
  memcpy(symbolCount, h.symbolCount, sizeof(h.symbolCount));
  memcpy(cumul, h.cumul, sizeof(h.cumul));
}

void hsref_observe_hist(const uint8_t *data, size_t size, uint32_t hist[256]) { observe_hist(hist, data, size); }

} // extern "C"
