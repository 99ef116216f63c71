/*
 * hsrans_b200.h — C-ABI of the B200-native (sm_100a) decoder for hypersonic-rANS's interleaved
 * 32-bit-state / 16-bit-word streams: rANS32x32_16w, rANS32x64_16w, block_rANS32x{32,64}_16w and
 * mt_rANS32x{32,64}_16w, probability bits 10..15, plus the histogram count/normalise step.
 *
 * Plain pointers and sizes only; no torch, no C++ types. Every entry point names the reference interface it
 * replaces (file:line relative to the reference repository). The shared object is
 * hypersonic-rans_b200/libhsrans_b200.so; C++ callers can use hypersonic-rans_b200/cpp/hsrans_b200_codecs.hpp,
 * which wraps these calls in functions that carry the reference's exact per-bits names and `decodeFunc` type.
 *
 * Conventions copied from the reference (src/rANS32x32_16w.cpp:164-191): decoders return the decoded byte
 * count on success and 0 on ANY error (short input, capacity too small, histogram not summing to 2^bits,
 * misaligned block end, CUDA failure, no GPU). There is no CPU fallback: without a usable CUDA device every
 * compute entry point returns 0 / a negative status and hsr_last_error() says why.
 */
#ifndef HSRANS_B200_H
#define HSRANS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSR_VERSION 200 /* 200: + hsr_encode_mt_device_indexed, hsr_stream_from_device_indexed, options "overlap" / "contexts";
                            overlapping launches, pooled host contexts. 110: + HSR_RAW32BLK, stateCount 16, policy encoder */

/* Stream framings (SURVEY.md §8a "Stream formats"). */
typedef enum hsr_family {
  HSR_RAW = 0,   /* rANS32xN_16w        — one histogram, one recurrence   (src/rANS32x32_16w.cpp:130-158)            */
  HSR_BLOCK = 1, /* block_rANS32xN_16w  — in-band histograms, carried states (src/block_rANS32x32_16w_encode.cpp:262-285) */
  HSR_MT = 2,    /* mt_rANS32xN_16w     — independent blocks with state snapshots (src/mt_rANS32x64_16w_encode.cpp:266-298) */
  HSR_RAW32BLK = 3 /* rANS32x32_32blk_16w — raw header + u32 blockSize[31], one private word sub-stream per state
                      (src/rans32x32_32blk_16w.cpp:147-176); stateCount must be 32 */
} hsr_family_t;
/* stateCount: 32 or 64 for every family; HSR_RAW also takes 16 = rANS32x16_16w (src/rANS32x16_16w.cpp:162-271). */

/* ------------------------------------------------------------------------------------------------ general */

int hsr_version(void);
/* Number of usable CUDA devices (0 if none / driver missing). */
int hsr_device_count(void);
/* Thread-local description of the last failure in this thread ("" if none). */
const char *hsr_last_error(void);
/* Tuning knobs, for benchmarking variants without rebuilding. Unknown keys return -1.
 *   "table"      0 = auto (packed slot table for bits <= 12, bitmap-rank table above; wide one-lookup tables for
 *                13..15 bits when a launch has at most one unit per SM), 1 = bitmap-rank, 2 = packed, 3 = wide
 *   "warps"      cap on resident one-warp CTAs per SM for the mt_ kernel (1..32, 0 = as many as fit; experiments)
 *   "chunk_mb"   host pipeline chunk size in MiB for hsr_decode on mt_ streams (0 = auto)
 *   "index"      block index of device-resident mt_ streams: 0 = segment-parallel with serial fallback,
 *                1 = serial walk only, 2 = segment-parallel only (fail instead of falling back; tests)
 *   "overlap"    1 (default) = consecutive decode launches on one CUDA stream overlap: the next launch's warps take
 *                over SM slots as the previous launch drains (programmatic dependent launch); 0 = strictly serial
 *   "contexts"   host-pointer decodes that may be in flight per device at once (1..16, default 4); each holds its
 *                own streams and device scratch, further callers wait
 *   "batch_group_mb"  hsr_decode_batch: traffic per pipeline group in MiB (0 = auto: enough bytes per group for its
 *                copies to outlast the one-warp decode time of its longest stream; tests)                           */
int hsr_set_option(const char *key, long value);
long hsr_get_option(const char *key);

/* Worst-case compressed size for n input bytes (the buffer size a harness must allocate).
 * Replaces rANS32x32_16w_capacity (src/rANS32x32_16w.cpp:10-13), rANS32x16_16w_capacity (src/rANS32x16_16w.cpp:10-13),
 * rANS32x32_32blk_16w_capacity (src/rans32x32_32blk_16w.cpp:10-13), block_rANS32x32_16w_capacity
 * (src/block_rANS32x32_16w_encode.cpp:47-54), mt_rANS32x64_16w_capacity (src/mt_rANS32x64_16w_encode.cpp:50-57). */
size_t hsr_capacity(int family, int stateCount, size_t inputSize);

/* Page-locked host buffers for the harness (the reference harness allocates 64-byte aligned buffers,
 * src/main.cpp:649-650). hsr_decode accepts any host pointer; pinned ones avoid a staging copy. */
void *hsr_host_alloc(size_t bytes);
void hsr_host_free(void *p);

/* One stream of a batch (hsr_decode_batch, hsr_stream_upload_batch): byte ranges relative to the batch's base pointers. */
typedef struct hsr_batch_item {
  uint64_t inOffset, inLength, outOffset, outCapacity;
} hsr_batch_item_t;

/* ------------------------------------------------------------------------------------------------ drop-in decode */

/* Host-pointer decode: H2D, index (mt_), kernels, D2H, synchronise. Replaces every function of type
 *   size_t f(const uint8_t *pInData, const size_t inLength, uint8_t *pOutData, const size_t outCapacity)
 * i.e. codec_info_t::decodeFunc (src/main.cpp:149): rANS32x32_16w_decode_scalar_<b> (src/rANS32x32_16w.cpp:161),
 * rANS32x64_16w_decode_scalar_<b> (src/rANS32x64_16w.cpp:168), rANS32x16_16w_decode_scalar_<b>
 * (src/rANS32x16_16w.cpp:162), rANS32x32_32blk_16w_decode_scalar_<b> (src/rans32x32_32blk_16w.cpp:183) and all their AVX variants,
 * block_rANS32xNN_16w_decode_<b> (src/block_rANS32x32_16w_decode.cpp:165-193),
 * mt_rANS32xNN_16w_decode_<b> / _decode_mt_<b> (src/mt_rANS32x64_16w_decode.cpp:301-361).
 * Uses the current CUDA device (hsr_set_device). Re-entrant like the reference's decoders (SURVEY.md 8b "Threading"):
 * calls from several host threads run concurrently, each on its own pooled context (option "contexts").
 * Limits: decoded length >= stateCount (below that the reference itself is undefined, src/rANS32x32_16w.cpp:206); a
 * raw stream or a single mt_ block may not exceed 4 GiB of compressed bytes (32-bit word cursor per warp); the
 * prepared-stream entry points (hsr_stream_*) reject decoded lengths above 1 TiB. */
size_t hsr_decode(int family, int stateCount, int bits, const uint8_t *pInData, size_t inLength, uint8_t *pOutData,
                  size_t outCapacity);
/* Same, mt_ only, sharding the block chain over `deviceCount` GPUs of this box from ONE process by contiguous
 * block ranges balanced on compressed bytes (the GPU analogue of the thread pool in
 * src/mt_rANS32x64_16w_decode.cpp:137-265). devices == NULL means 0..deviceCount-1. */
size_t hsr_decode_mt_multi(int stateCount, int bits, const uint8_t *pInData, size_t inLength, uint8_t *pOutData,
                           size_t outCapacity, const int *devices, int deviceCount);

/* One process per GPU: decodes only shard `shard` of `shards` of an mt_ chain — the same contiguous block ranges as
 * hsr_stream_upload(shard, shards) — on the current device, from the caller's host buffers, into the caller's
 * full-size output buffer at the shard's own offset (written to *pShardOffset). Only the shard's compressed bytes cross
 * PCIe. Returns the decoded bytes of the shard; 0 on error or when the shard owns no block (hsr_last_error() is ""). */
size_t hsr_decode_mt_shard(int stateCount, int bits, const uint8_t *pInData, size_t inLength, uint8_t *pOutData, size_t outCapacity,
                           int shard, int shards, size_t *pShardOffset);

int hsr_set_device(int device);

/* Many independent streams of ONE codec in one launch. A raw or block_ stream is a single 32/64-lane recurrence
 * (one warp of parallelism, SURVEY.md finding 1), so batching streams is what fills the GPU for those codecs — the
 * analogue of running the reference's decoder on many files from many threads. Stream i occupies
 * inBase[inOffset, inOffset + inLength) and decodes to outBase[outOffset, outOffset + n) with n <= outCapacity;
 * inOffset must be even (streams are sequences of 16-bit words; an odd one marks the stream malformed); output ranges
 * must not overlap. decodedLengths[i] receives n, or 0 if stream i is malformed (the others are still
 * decoded; the output range of a malformed stream is left untouched). Returns the number of streams decoded. Each
 * stream is checked like hsr_decode checks its input. The call is a pipeline: the input range goes to the device in
 * pieces, groups of streams (in input order) are launched as their bytes land, and their decoded bytes return while
 * later groups decode.
 *
 * Diagnostics: with the environment variable HSR_TRACE_PIPELINE set, hsr_decode (mt_) and hsr_decode_batch print the
 * device timestamps of every copy piece, launch and copy-back of the call to stderr. */

size_t hsr_decode_batch(int family, int stateCount, int bits, const uint8_t *inBase, uint8_t *outBase,
                        const hsr_batch_item_t *items, size_t count, uint64_t *decodedLengths);

/* ------------------------------------------------------------------------------------------------ mt_ index (host) */

/* One unit of work for the kernels. For mt_ it is one block of the chain; single-symbol runs are split into
 * fills of bounded size. */
typedef struct hsr_block {
  uint64_t inOffset;  /* coded: byte offset of the block's u32 states[N] (mt_) or of the u16 counts (raw) */
  uint64_t inEnd;     /* coded: byte offset one past the block's last word */
  uint64_t outOffset; /* first decoded byte of this unit */
  uint64_t count;     /* decoded bytes in this unit (coded: rows*N, plus `tail` extra lanes on the last one) */
  uint32_t kind;      /* 0 coded mt_ layout, 1 fill, 2 coded raw layout, 3 coded raw 32blk layout */
  uint32_t symbol;    /* fill value for kind 1 */
  uint32_t tail;      /* 0, or the number of leftover symbols (< N) decoded after the last full row */
  uint32_t reserved;
} hsr_block_t;

/* Walks the mt_ header chain on the host (src/mt_rANS32x64_16w_decode.cpp:40-66,94 — the serial walk the
 * reference does on the calling thread). Writes up to maxBlocks records; returns the number of units, or -1 on a
 * malformed chain. Call with blocks == NULL to count. */
long hsr_mt_index(int stateCount, const uint8_t *pInData, size_t inLength, hsr_block_t *blocks, size_t maxBlocks);

/* Contiguous partition of `count` units over `parts` shards, balanced on compressed bytes + decoded bytes.
 * firstUnit must hold parts+1 entries; shard p owns units [firstUnit[p], firstUnit[p+1]). */
int hsr_mt_partition(const hsr_block_t *blocks, size_t count, int parts, size_t *firstUnit);

/* ------------------------------------------------------------------------------------------------ device-resident API */

/* A stream prepared for repeated device-side decoding: the compressed bytes in HBM (byte-exact stream format,
 * 16-byte aligned base) plus the block index. This is what `value` in bench.py times. */
typedef struct hsr_stream hsr_stream_t;

/* Upload from host memory to the current device; mt_: also builds and uploads the index. For mt_ a shard can be
 * selected with (shard, shards): only that contiguous block range is uploaded and decoded (one process per GPU).
 * Returns NULL on error. */
hsr_stream_t *hsr_stream_upload(int family, int stateCount, int bits, const uint8_t *pInData, size_t inLength, int shard,
                                int shards);
/* Wrap compressed bytes that already live in device memory (16-byte aligned, readable up to inLength).
 * mt_: the block index is built on the device — K warps each find the first header of their stream segment and
 * walk the chain from there, handing over where they meet (hsr_index.cu); see hsr_stream_index_ms. */
hsr_stream_t *hsr_stream_from_device(int family, int stateCount, int bits, const void *dIn, size_t inLength);
/* The same for an mt_ stream whose producer already knows where every block lies (hsr_encode_mt_device_indexed):
 * dIndex holds numUnits hsr_block_t records in chain order (device memory). The table is copied back and checked
 * record by record against the stream bounds — nothing it claims can make a kernel read or write out of range —
 * and the header chain is not walked again. */
hsr_stream_t *hsr_stream_from_device_indexed(int stateCount, int bits, const void *dIn, size_t inLength, const hsr_block_t *dIndex,
                                             size_t numUnits);
/* A batch of independent streams of one codec (see hsr_decode_batch) made resident for repeated decoding. Offsets in
 * `items` are relative to inBase; the decode writes stream i at dOut + items[i].outOffset, so dOut must hold
 * hsr_stream_decoded_length() = max(outOffset + n) bytes. Fails if any stream of the batch is malformed. */
hsr_stream_t *hsr_stream_upload_batch(int family, int stateCount, int bits, const uint8_t *inBase, const hsr_batch_item_t *items,
                                      size_t count);
void hsr_stream_free(hsr_stream_t *s);

uint64_t hsr_stream_decoded_length(const hsr_stream_t *s); /* header n (whole stream) */
uint64_t hsr_stream_shard_out_offset(const hsr_stream_t *s); /* first decoded byte owned by this shard */
uint64_t hsr_stream_shard_out_bytes(const hsr_stream_t *s);  /* decoded bytes owned by this shard */
uint64_t hsr_stream_shard_in_bytes(const hsr_stream_t *s);   /* compressed bytes resident for this shard */
uint64_t hsr_stream_units(const hsr_stream_t *s);
double hsr_stream_index_ms(const hsr_stream_t *s);           /* time spent building the block index */
int hsr_stream_copy_index(const hsr_stream_t *s, hsr_block_t *blocks, size_t maxBlocks);

/* Launch the decode of the prepared stream into device memory on `cudaStream` (a cudaStream_t, may be NULL).
 * dOut is the base of the WHOLE decoded buffer (shards write at their own offsets) and must hold
 * hsr_stream_decoded_length bytes — or, with HSR_OUT_SHARD_LOCAL, only this shard's bytes.
 * Asynchronous; returns the number of kernels launched (0 for a shard that owns no block of the chain) or a negative
 * status. Any number of decodes of one hsr_stream_t may be in flight, on one CUDA stream or several: every launch
 * draws a private work counter from a per-device ring. Consecutive launches on one CUDA stream overlap (option
 * "overlap"): a decode never reads what another decode wrote, so the next one's warps fill the SM slots the previous
 * one frees while it drains; completion is still in stream order. The status bits are shared by all launches. */
#define HSR_OUT_SHARD_LOCAL 1u
int hsr_stream_decode_async(hsr_stream_t *s, void *dOut, size_t outCapacity, unsigned flags, void *cudaStream);
/* After synchronising: 0 if the last decode saw a well-formed stream, else a bit set of HSR_ERR_*. */
#define HSR_ERR_HIST 1u     /* a histogram does not sum to 2^bits (src/hist.cpp:308-324) */
#define HSR_ERR_OVERRUN 2u  /* the word cursor ran past the block / stream end */
#define HSR_ERR_ALIGN 4u    /* block end not a multiple of the state count (src/block_rANS32x32_16w_decode.cpp:82-83) */
#define HSR_ERR_BOUNDS 8u   /* a block would write past the decoded length */
#define HSR_ERR_INTERNAL 16u /* a staging barrier never completed (should not happen; reported instead of hanging) */
unsigned hsr_stream_status(hsr_stream_t *s);

/* ------------------------------------------------------------------------------------------------ histogram */

/* Replaces make_hist (src/hist.cpp:217-222) = observe_hist (:8-14) + normalize_hist (:16-215), bit-exact.
 * Host-pointer form: copies the data to the device, counts with a shared-memory-atomic kernel, normalises on
 * the device with the reference's exact float and heap-sort sequence, returns the 256 counts and cumuls. */
int hsr_make_hist(const uint8_t *pData, size_t size, int bits, uint16_t symbolCount[256], uint16_t cumul[256]);
/* Device-pointer forms. dHist: 256 x u32 byte counts. dSymbolCount/dCumul: 256 x u16 each. */
int hsr_observe_hist_device(const void *dData, size_t size, uint32_t *dHist, void *cudaStream);
int hsr_normalize_hist_device(const uint32_t *dHist, size_t dataBytes, int bits, uint16_t *dSymbolCount,
                              uint16_t *dCumul, void *cudaStream);
/* Segmented form for block_/mt_ style per-block histograms: segment k covers bytes [k*segmentBytes,
 * min((k+1)*segmentBytes, size)); writes 256 u16 counts per segment, each summing to 2^bits. */
int hsr_make_hist_segments_device(const void *dData, size_t size, size_t segmentBytes, int bits,
                                  uint16_t *dSymbolCounts, void *cudaStream);

/* ------------------------------------------------------------------------------------------------ mt_ encoder (device) */

/* Produces an mt_rANS32xN_16w stream on the GPU. Same stream format as mt_rANS32xNN_16w_encode_<b>
 * (src/mt_rANS32x64_16w_encode.cpp:140-380; signature src/mt_rANS32x64_16w.h:9-14) and the same per-block histogram
 * (observe_hist + normalize_hist over the block's bytes), but with a FIXED block size (0 = 65536 = the reference's
 * MinBlockSize; a multiple of the state count, at most 2^25) and every block encoded from fresh states, so blocks
 * are independent GPU work and the result always has ~n / blockSize blocks whatever the data's stationarity. Any
 * reference decoder decodes it; for inputs of at most one block it is byte-identical to the reference encoder's
 * output (constant inputs excepted: no single-symbol run blocks are emitted). Returns the compressed length, 0 on
 * error (length < stateCount, outCapacity too small, no GPU). */
size_t hsr_encode_mt(int stateCount, int bits, const uint8_t *pInData, size_t length, uint8_t *pOutData, size_t outCapacity,
                     size_t blockSize);
/* Device-pointer form; dOut should hold hsr_encode_mt_bound() bytes. Synchronises cudaStream once. */
size_t hsr_encode_mt_device(int stateCount, int bits, const void *dIn, size_t length, void *dOut, size_t outCapacity, size_t blockSize,
                            void *cudaStream);
/* The same producer with the reference's block-split POLICY decided on the device: blocks are
 * whole numbers of 64 KiB segments, grown while coding the next segment with the block's histogram costs less than
 * a histogram of its own plus half a header (_CanExtendHist, src/mt_rANS32x64_16w_encode.cpp:61-136), and stretches
 * of one repeated byte become 8-byte run blocks (:171-187, :300-305). maxBlockSize (0 = 262144; a multiple of 65536,
 * at most 2^25) bounds a block and is also the grain at which the split runs in parallel, so the stream keeps at
 * least length / maxBlockSize independent blocks for the GPU decoder. hsr_encode_mt_bound(N, length, 0) bounds the size. */
size_t hsr_encode_mt_policy(int stateCount, int bits, const uint8_t *pInData, size_t length, uint8_t *pOutData, size_t outCapacity,
                            size_t maxBlockSize);
size_t hsr_encode_mt_policy_device(int stateCount, int bits, const void *dIn, size_t length, void *dOut, size_t outCapacity,
                                   size_t maxBlockSize, void *cudaStream);
/* Hard upper bound of the stream size for `length` input bytes (one 16-bit word per symbol + headers). */
size_t hsr_encode_mt_bound(int stateCount, size_t length, size_t blockSize);
/* hsr_encode_mt_device (policy == 0; blockSize as there) or hsr_encode_mt_policy_device (policy != 0; blockSize is its
 * maxBlockSize) that also hands out what the encoder knows anyway: the decoder's unit table, one hsr_block_t per block
 * in chain order, written to dIndex (device memory with room for indexCapacity >= hsr_encode_mt_index_bound()
 * records); *numUnits receives the count. Feed both to hsr_stream_from_device_indexed(). */
size_t hsr_encode_mt_device_indexed(int stateCount, int bits, const void *dIn, size_t length, void *dOut, size_t outCapacity,
                                    size_t blockSize, int policy, hsr_block_t *dIndex, size_t indexCapacity, size_t *numUnits,
                                    void *cudaStream);
size_t hsr_encode_mt_index_bound(int stateCount, size_t length, size_t blockSize);

/* ------------------------------------------------------------------------------------------------ synthetic inputs */

/* Deterministic Zipf(s) bytes over 256 symbols. segmentBytes == 0: one rank->byte permutation
 * for the whole buffer ("iid"); otherwise the permutation is re-drawn every segmentBytes ("pw64k" = 65536).
 * Host-side, multi-threaded; not on the decode path. */
int hsr_synth_zipf(uint8_t *out, size_t n, double s, uint64_t seed, size_t segmentBytes);

#ifdef __cplusplus
}
#endif
#endif /* HSRANS_B200_H */
