// hsr_kernels_aux.cu — the two remaining 32-bit-state / 16-bit-word raw layouts of the reference's registry
// (SURVEY.md §8f rank 4), on the same tables, word ring and unit records as the main kernels:
//
//   rANS32x16_16w        src/rANS32x16_16w.cpp:162-271: the raw format with 16 interleaved states. Lanes 0..15 of
//                        the warp own the states, lanes 16..31 idle; byte position inside a row of 16 is
//                        { 0-3, 8-11, 4-7, 12-15 } (:211). Words are handed out by the same ballot/popc prefix.
//   rANS32x32_32blk_16w  src/rans32x32_32blk_16w.cpp:183-301: 32 states, but every state reads its words from its
//                        OWN sub-stream (u32 blockSize[31] after the states, :223-231), so there is no shared
//                        cursor at all: each lane walks a private read head through global memory, its next two
//                        words always preloaded into registers (L1-resident lines, prefetched two lines ahead).
//
// Both are single recurrences (one warp per stream, latency-bound like every raw stream); throughput comes from
// batches (hsr_decode_batch), which these kernels serve through the same persistent unit loop.
#include <type_traits>

#include "hsr_kernels.cuh"

namespace hsr {

// lane -> byte position inside a row of 16 (src/rANS32x16_16w.cpp:211)
__device__ __forceinline__ uint32_t idx2idx16_lane(uint32_t l) { return (l & 3u) | ((l & 4u) << 1) | ((l & 8u) >> 1); }

struct UnitView {
  const uint8_t *base, *end;
  uint8_t *out;
  uint64_t count;
  uint32_t kind, tail, streamId;
};

__device__ __forceinline__ bool next_unit(const DecodeParams &p, uint32_t lane, UnitView *u)
{
  uint32_t b = 0;
  if (lane == 0)
    b = atomicAdd(p.work, 1u);
  b = __reduce_add_sync(kFull, b); // lane 0's claim in a uniform register: a provably warp-uniform loop exit (see units_kernel_body)
  if (b >= p.numBlocks)
    return false;
  const hsr_block_t *blk = p.blocks + b;
  u->base = p.in + (__ldg(&blk->inOffset) - p.inBase);
  u->end = p.in + (__ldg(&blk->inEnd) - p.inBase);
  u->out = p.out + (__ldg(&blk->outOffset) - p.outBase);
  u->count = __ldg(&blk->count);
  u->kind = __ldg(&blk->kind);
  u->tail = __ldg(&blk->tail);
  u->streamId = __ldg(&blk->reserved);
  return true;
}

// ---------------------------------------------------------------------------------------------- rANS32x16_16w

// which symbol step a unit uses: the fast one-lookup table of this kernel, or the rank tables that are always built
template <int TK>
__device__ __forceinline__ int step_mode(const TableInfo &info) { return TK == TK_WIDE ? 3 : (TK == TK_PACKED && !info.degenerate) ? 0 : 2; }

template <int BITS, int TK>
__device__ __forceinline__ void raw16_kernel_body(const DecodeParams &p)
{
  using L = WarpLayout<BITS, 32, TK>; // a row of 16 consumes at most 32 bytes: inside the 32-state ring's overlap
  uint32_t sw;
  if constexpr (L::kDynamic)
    sw = dynamic_smem_base();
  else
    sw = declare_smem<L::kBytes>();
  const uint32_t lane = lane_id();
  const uint32_t ltMask = lanemask_lt();
  const bool live = lane < 16u;
  const uint32_t lanePos = idx2idx16_lane(lane & 15u);

  Decoder<BITS, 32, TK> dec;
  dec.init(sw);
  Ring<L> ring;
#if HSR_RING_TMA
  ring.init(sw + L::kOffRing, sw + L::kOffBar, lane);
#endif

  units_enter();
  UnitView u;
  while (next_unit(p, lane, &u)) {
    if (u.kind != 2u) { // only whole raw streams exist for this codec
      raise(p.status, p.streamStatus, u.streamId, HSR_ERR_INTERNAL, lane);
      continue;
    }
    const uint8_t *countsPtr = u.base; // counts, then u32 states[16], then words (src/rANS32x16_16w.cpp:183-203)
    const uint8_t *statesPtr = u.base + 512;
    const uint8_t *words = u.base + 512 + 4 * 16;
    const TableInfo info = build_tables<BITS, 32, TK>(sw, countsPtr, lane);
    if (!info.ok) {
      raise(p.status, p.streamStatus, u.streamId, HSR_ERR_HIST, lane);
      continue;
    }
    // idle lanes hold a state that never asks for a word; their lookups stay inside the tables
    uint32_t x = live ? ldg_u32_a2(statesPtr + 4 * lane) : 0x80000000u;
#if HSR_RING_TMA
    ring.start(words, u.end, lane);
#else
    ring.start(sw + L::kOffRing, words, u.end, lane);
#endif
    ring.start_wait();
    const uint64_t rows = (u.count - u.tail) / 16u;
    uint8_t *outLane = u.out + lanePos;
    auto run = [&](auto modeTag) {
      constexpr int kMode = decltype(modeTag)::value;
#pragma unroll 4
      for (uint64_t r = 0; r < rows; r++) { // :213-238
        ring.advance_if_needed(lane);
        const uint32_t s = dec.template symbol_step<kMode>(x);
        if (live)
          st_global_u8(outLane, s);
        else
          x = 0x80000000u; // idle lanes: pinned above the consume point, so the unmasked hand-out skips them
        dec.renorm(x, ring.wp, ltMask);
        outLane += 16;
      }
    };
    const int mode = step_mode<TK>(info);
    if (mode == 3) run(std::integral_constant<int, 3>{});
    else if (mode == 0) run(std::integral_constant<int, 0>{});
    else run(std::integral_constant<int, 2>{});
    if (u.tail) { // :240-268
      ring.advance_if_needed(lane);
      const bool a = live && lanePos < u.tail;
      uint32_t t = x;
      const uint32_t s = dec.template symbol_step_rank<false>(t);
      if (a) {
        x = t;
        st_global_u8(outLane, s);
      }
      dec.renorm_masked(x, ring.wp, a, ltMask);
    }
    ring.drain();
    if (ring.cursor() > ring.glimit)
      raise(p.status, p.streamStatus, u.streamId, HSR_ERR_OVERRUN, lane);
  }
  units_leave(p.work, lane);
}

// ---------------------------------------------------------------------------------------------- rANS32x32_32blk_16w

template <int BITS, int TK>
__device__ __forceinline__ void blk32_kernel_body(const DecodeParams &p)
{
  using L = WarpLayout<BITS, 32, TK>;
  uint32_t sw;
  if constexpr (L::kDynamic)
    sw = dynamic_smem_base();
  else
    sw = declare_smem<L::kBytes>();
  const uint32_t lane = lane_id();
  const uint32_t lanePos = idx2idx_lane(lane); // same permutation as rANS32x32_16w (src/rans32x32_32blk_16w.cpp:235)

  Decoder<BITS, 32, TK> dec;
  dec.init(sw);

  units_enter();
  UnitView u;
  while (next_unit(p, lane, &u)) {
    if (u.kind != 3u) {
      raise(p.status, p.streamStatus, u.streamId, HSR_ERR_INTERNAL, lane);
      continue;
    }
    const uint8_t *countsPtr = u.base; // counts, u32 states[32], u32 blockSize[31], then the 32 sub-streams (:205-231)
    const uint8_t *statesPtr = u.base + 512;
    const uint8_t *sizesPtr = statesPtr + 4 * 32;
    const uint8_t *data = sizesPtr + 4 * 31;
    const TableInfo info = build_tables<BITS, 32, TK>(sw, countsPtr, lane);
    if (!info.ok) {
      raise(p.status, p.streamStatus, u.streamId, HSR_ERR_HIST, lane);
      continue;
    }
    uint32_t x = ldg_u32_a2(statesPtr + 4 * lane);

    // pReadHead[j] = pReadHead[j - 1] + blockSize[j - 1] (:223-231): exclusive prefix sum of the 31 sizes
    const uint64_t mine = lane < 31u ? (uint64_t)ldg_u32_a2(sizesPtr + 4 * lane) : 0ull;
    uint64_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint64_t up = __shfl_up_sync(kFull, incl, d);
      if (lane >= (uint32_t)d)
        incl += up;
    }
    const uint64_t avail = (uint64_t)(u.end - data);
    const uint64_t startOff = incl - mine;
    if (__any_sync(kFull, startOff > avail)) { // a sub-stream would begin past the end of the stream
      raise(p.status, p.streamStatus, u.streamId, HSR_ERR_OVERRUN, lane);
      continue;
    }
    // Sub-streams are sequences of 16-bit words: an odd size field (corrupt header) would make every later read head
    // odd, and a misaligned 16-bit load is a sticky fault on the GPU where the reference merely reads unaligned.
    if (__any_sync(kFull, (startOff & 1ull) != 0)) {
      raise(p.status, p.streamStatus, u.streamId, HSR_ERR_ALIGN, lane);
      continue;
    }
    // Each lane always holds its next TWO words in registers. A renormalisation takes the first, promotes the second
    // and requests the one after it — in straight-line code that every lane executes every row (re-reading the same
    // word when none was taken). The address of that load depends on this row's decision, but its value is only
    // needed at the lane's renormalisation after the next one, so the load (an L1 hit: the lane's 128-byte line is
    // prefetched two lines ahead) never sits on the symbol -> renormalise -> symbol chain.
    // Everything below is selects and predicated instructions: a lone warp pays ~100 cycles for every divergent
    // branch region in its loop (measured: 254 cycles per row this way, 344-474 with branches around the refill).
    const uint8_t *rd = data + startOff;
    const uint8_t *const last = u.end - 2; // highest address a word may be read from
    asm volatile("prefetch.global.L1 [%0];" ::"l"(rd));
    if (rd + 128 <= last) asm volatile("prefetch.global.L1 [%0];" ::"l"(rd + 128));
    uint32_t w0 = rd <= last ? ldg_u16(rd) : 0u;
    uint32_t w1 = rd + 2 <= last ? ldg_u16(rd + 2) : 0u;
    bool bad = false;

    auto renorm = [&](bool may) {
      const bool take = may && x < kConsumePoint16;
      x = take ? ((x << 16) | w0) : x;
      bad = bad || (take && rd > last); // the word came from beyond the end of the stream (read as zero)
      w0 = take ? w1 : w0;
      rd += take ? 2 : 0;
      if (take && (reinterpret_cast<uintptr_t>(rd) & 127u) == 0 && rd + 256 <= last)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rd + 256));
      w1 = rd + 2 <= last ? ldg_u16(rd + 2) : 0u;
    };

    const uint64_t rows = (u.count - u.tail) / 32u;
    uint8_t *outLane = u.out + lanePos;
    auto run = [&](auto modeTag) {
      constexpr int kMode = decltype(modeTag)::value;
#pragma unroll 2
      for (uint64_t r = 0; r < rows; r++) { // :241-269
        const uint32_t s = dec.template symbol_step<kMode>(x);
        st_global_u8(outLane, s);
        renorm(true);
        outLane += 32;
      }
    };
    const int mode = step_mode<TK>(info);
    if (mode == 3) run(std::integral_constant<int, 3>{});
    else if (mode == 0) run(std::integral_constant<int, 0>{});
    else run(std::integral_constant<int, 2>{});
    if (u.tail) { // :271-298
      const bool a = lanePos < u.tail;
      uint32_t t = x;
      const uint32_t s = dec.template symbol_step_rank<false>(t);
      if (a) {
        x = t;
        st_global_u8(outLane, s);
      }
      renorm(a);
    }
    if (__any_sync(kFull, bad))
      raise(p.status, p.streamStatus, u.streamId, HSR_ERR_OVERRUN, lane);
  }
  units_leave(p.work, lane);
}

// ---------------------------------------------------------------------------------------------- instantiation

// per bit width: the bitmap-rank kernel (any number of units) and a one-lookup kernel for launches the host judges
// latency-bound or small-tabled (packed u32/slot up to 12 bits, wide tables in dynamic shared memory above)
#define HSR_AUX_DEFINE(BITS, FAST)                                                                                    \
  __global__ void __launch_bounds__(32, 16) raw16_b##BITS(DecodeParams p) { raw16_kernel_body<BITS, TK_RANK>(p); }    \
  __global__ void __launch_bounds__(32, 16) blk32_b##BITS(DecodeParams p) { blk32_kernel_body<BITS, TK_RANK>(p); }    \
  __global__ void __launch_bounds__(32, 16) raw16_fast_b##BITS(DecodeParams p) { raw16_kernel_body<BITS, FAST>(p); }  \
  __global__ void __launch_bounds__(32, 16) blk32_fast_b##BITS(DecodeParams p) { blk32_kernel_body<BITS, FAST>(p); }

HSR_AUX_DEFINE(10, TK_PACKED) HSR_AUX_DEFINE(11, TK_PACKED) HSR_AUX_DEFINE(12, TK_PACKED)
HSR_AUX_DEFINE(13, TK_WIDE) HSR_AUX_DEFINE(14, TK_WIDE) HSR_AUX_DEFINE(15, TK_WIDE)

#define HSR_AUX_ENTRY(name, BITS, TK) { (const void *)name##BITS, nullptr, WarpLayout<BITS, 32, TK>::kBytes, WarpLayout<BITS, 32, TK>::kDynamic ? 1 : 0 }
#define HSR_AUX_ROW(prefix, BITS, FAST) { HSR_AUX_ENTRY(prefix##_b, BITS, TK_RANK), HSR_AUX_ENTRY(prefix##_fast_b, BITS, FAST) }

extern const KernelEntry kKernelsRaw16[6][2] = { HSR_AUX_ROW(raw16, 10, TK_PACKED), HSR_AUX_ROW(raw16, 11, TK_PACKED), HSR_AUX_ROW(raw16, 12, TK_PACKED),
                                                 HSR_AUX_ROW(raw16, 13, TK_WIDE), HSR_AUX_ROW(raw16, 14, TK_WIDE), HSR_AUX_ROW(raw16, 15, TK_WIDE) };
extern const KernelEntry kKernelsBlk32[6][2] = { HSR_AUX_ROW(blk32, 10, TK_PACKED), HSR_AUX_ROW(blk32, 11, TK_PACKED), HSR_AUX_ROW(blk32, 12, TK_PACKED),
                                                 HSR_AUX_ROW(blk32, 13, TK_WIDE), HSR_AUX_ROW(blk32, 14, TK_WIDE), HSR_AUX_ROW(blk32, 15, TK_WIDE) };

} // namespace hsr
