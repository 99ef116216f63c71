// hsr_device.cuh — warp-level rANS32xN_16w decoder for sm_100a.
//
// One interleaved rANS state per warp lane (N = 64: lane l owns states l and l + 32). Semantics restated from
// the reference's scalar section decoder (src/block_codec32.h:162-206, src/block_codec64.h:173-217,
// src/rANS32x32_16w.cpp:17-30): per row of N symbols, in state order j,
//     slot = x & (2^b - 1); s = slotToSymbol[slot]; out[i + idx2idx(j)] = s
//     x = (x >> b) * freq[s] + slot - cumul[s]
//     if (x < 2^15) x = (x << 16) | *readHead++
// The AVX movemask/popcnt + shuffle-LUT word hand-out (src/rANS32x32_16w.cpp:1229-1290) becomes
// __ballot_sync + __popc(mask & lanemask_lt) on a warp-wide word cursor.
//
// Shared-memory tables (private layouts; only the decoded bytes have to match the reference):
//   TK_RANK   bitmap-rank table, any bits: one u32 per 16 slots {8*starts_before_group : 16 | start_bitmap : 16}
//             and one 8-byte entry per present symbol {freq - 2^b, (-cumul) << 8 | symbol}.  2^(b-2) + 2 KB
//             (10 KB at 15 bits vs 33 KB for the reference's hist_dec2_t, src/hist.h:32-37) and O(256 + 2^b/16)
//             to build instead of O(2^b).
//   TK_PACKED one u32 per slot {freq : 12 | slot - cumul : 12 | symbol : 8}, bits <= 12, one lookup on the chain
//             (the reference's hist_dec_pack_t idea, src/hist.h:46-50, with the bias pre-subtracted).
// Compressed words are staged by cp.async (LDGSTS, 16 B per lane) into a per-warp ring of overlapping linear
// segments, so the data-dependent word reads are LDS, never exposed DRAM latency.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/hsrans_b200.h"

namespace hsr {

constexpr uint32_t kConsumePoint16 = 1u << 15; // src/rans.h:8
constexpr unsigned kFull = 0xffffffffu;

enum TableKind : int { TK_RANK = 1, TK_PACKED = 2 };

// ---------------------------------------------------------------------------------------------- small helpers

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t lanemask_lt()
{
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// lane -> byte position inside a group of 32 (src/block_codec32.h:22): [0-3,16-19,4-7,20-23,8-11,24-27,12-15,28-31]
__device__ __forceinline__ uint32_t idx2idx_lane(uint32_t l) { return (l & 3u) | ((l & 4u) << 2) | ((l & 24u) >> 1); }

// every header field is only 2-byte aligned (src/mt_rANS32x64_16w_decode.cpp:43,57,64)
__device__ __forceinline__ uint32_t ldg_u16(const uint8_t *p) { return __ldg(reinterpret_cast<const uint16_t *>(p)); }
__device__ __forceinline__ uint32_t ldg_u32_a2(const uint8_t *p) { return ldg_u16(p) | (ldg_u16(p + 2) << 16); }
__device__ __forceinline__ uint64_t ldg_u64_a2(const uint8_t *p) { return (uint64_t)ldg_u32_a2(p) | ((uint64_t)ldg_u32_a2(p + 4) << 32); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds_u64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u64(uint32_t a, uint2 v) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t srcBytes)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(srcBytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int Pending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(Pending) : "memory"); }

__device__ __forceinline__ void st_global_u8(uint8_t *p, uint32_t v) { asm volatile("st.global.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// ---------------------------------------------------------------------------------------------- layout

template <int BITS, int N, int TK>
struct WarpLayout {
  static_assert(BITS >= 10 && BITS <= 15, "probability bits 10..15");
  static_assert(N == 32 || N == 64, "32 or 64 interleaved states");
  static_assert(TK == TK_RANK || (TK == TK_PACKED && BITS <= 12), "packed slot table only up to 12 bits");

  static constexpr int kSlots = 1 << BITS;
  static constexpr int kGroups = kSlots / 16;       // bitmap-rank groups of 16 slots
  static constexpr int kGrpBytes = kGroups * 4;
  static constexpr int kEntBytes = 257 * 8 + 8;     // entries are indexed 1..256 (starts up to and including the slot)
  static constexpr int kPackedBytes = TK == TK_PACKED ? kSlots * 4 : 0;

  // word ring: kBufs linear segments of kSeg bytes; consecutive segments overlap by one worst-case row
  static constexpr int kSeg = 512;                  // one 16-byte cp.async per lane
  static constexpr int kBufs = 4;
  static constexpr int kOverlap = 2 * N;            // a row consumes at most N words
  static constexpr int kStride = kSeg - kOverlap;
  static constexpr int kRingBytes = kSeg * kBufs;

  static constexpr int kOffGrp = 0;
  static constexpr int kOffEnt = kOffGrp + kGrpBytes;
  static constexpr int kOffPacked = kOffEnt + kEntBytes;
  static constexpr int kOffRing = kOffPacked + kPackedBytes;
  static constexpr int kBytes = kOffRing + kRingBytes; // per warp, multiple of 16
  static_assert(kBytes % 16 == 0, "per-warp shared memory must stay 16-byte aligned");
};

// ---------------------------------------------------------------------------------------------- word ring

// Segment k holds stream bytes [k*kStride, k*kStride + kSeg) relative to `gbase` (16-byte aligned, at or just
// below the first word). A row may read up to kOverlap bytes past the cursor, so the decoder leaves segment k
// once cursor - k*kStride >= kStride, which is exactly where segment k+1 begins.
template <class L>
struct WordRing {
  uint32_t sbuf;        // shared address of buffer 0
  const uint8_t *gbase; // 16-byte aligned
  uint32_t glimit;      // readable bytes from gbase (never read past the caller's inLength)
  uint32_t seg;         // current segment
  uint32_t segStart;    // seg * kStride
  uint32_t cur;         // cursor, bytes from gbase

  __device__ __forceinline__ void issue(uint32_t k, uint32_t lane) const
  {
    const uint32_t srcOff = k * L::kStride + lane * 16u;
    const uint32_t dst = sbuf + (k % L::kBufs) * L::kSeg + lane * 16u;
    uint32_t bytes = srcOff < glimit ? glimit - srcOff : 0u;
    bytes = bytes > 16u ? 16u : bytes;
    cp_async16(dst, gbase + (bytes ? srcOff : 0u), bytes); // bytes == 0: pure zero fill, no global read
    cp_async_commit();
  }

  __device__ __forceinline__ void start(uint32_t sbuf_, const uint8_t *firstWord, const uint8_t *streamEnd, uint32_t lane)
  {
    sbuf = sbuf_;
    const uintptr_t a = reinterpret_cast<uintptr_t>(firstWord);
    gbase = reinterpret_cast<const uint8_t *>(a & ~(uintptr_t)15);
    cur = (uint32_t)(a & 15);
    const uint64_t avail = (uint64_t)(streamEnd - gbase);
    glimit = avail > 0xffffffffull ? 0xffffffffu : (uint32_t)avail;
    seg = 0;
    segStart = 0;
    __syncwarp(); // every lane is done with whatever lived in the ring before
#pragma unroll
    for (uint32_t k = 0; k + 2 <= (uint32_t)L::kBufs; k++)
      issue(k, lane);
    cp_async_wait<L::kBufs - 2>();
    __syncwarp();
  }

  // call once per row, before any word of the row is read
  __device__ __forceinline__ void advance_if_needed(uint32_t lane)
  {
    if (cur - segStart >= (uint32_t)L::kStride) {
      seg += 1;
      segStart += L::kStride;
      issue(seg + L::kBufs - 2, lane); // lands in the buffer of segment seg-2, abandoned one segment ago
      cp_async_wait<L::kBufs - 2>();   // segment `seg` has landed for this lane ...
      __syncwarp();                    // ... and for all the others
    }
  }

  __device__ __forceinline__ uint32_t cursor_addr() const { return sbuf + (seg % L::kBufs) * L::kSeg + (cur - segStart); }

  __device__ __forceinline__ void drain() const { cp_async_wait<0>(); }
};

// ---------------------------------------------------------------------------------------------- table build

// Builds the warp's tables from 256 u16 counts at `counts` (2-byte aligned global memory).
// Returns false (warp-uniform) unless the counts sum to 2^BITS (src/hist.cpp:308-324).
template <int BITS, int N, int TK>
__device__ __forceinline__ bool build_tables(uint8_t *smemWarp, const uint8_t *counts, uint32_t lane)
{
  using L = WarpLayout<BITS, N, TK>;
  const uint32_t sGrp = smem_u32(smemWarp + L::kOffGrp);
  const uint32_t sEnt = smem_u32(smemWarp + L::kOffEnt);

  // lane l owns symbols 8l .. 8l+7
  uint32_t freq[8];
  uint32_t sum = 0, present = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    freq[k] = ldg_u16(counts + 2 * (lane * 8 + k));
    sum += freq[k];
    present += freq[k] != 0;
  }
  // one inclusive scan carries both running totals: frequencies (<= 2^16 per lane... kept in the low 20 bits)
  // and present-symbol counts (high 12 bits)
  uint32_t packed = sum | (present << 20);
  uint32_t incl = packed;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t up = __shfl_up_sync(kFull, incl, d);
    if (lane >= (uint32_t)d)
      incl += up;
  }
  const uint32_t total = __shfl_sync(kFull, incl, 31) & 0xfffffu;
  // a corrupt histogram can carry up to 256 * 65535 in the low field; 20 bits hold 2^20 - 1 < that, so check
  // the per-lane sums for overflow as well
  const bool laneOk = sum <= (uint32_t)L::kSlots;
  if (!__all_sync(kFull, laneOk) || total != (uint32_t)L::kSlots)
    return false;

  __syncwarp();
  // clear the bitmap groups
  for (uint32_t o = lane * 16u; o < (uint32_t)L::kGrpBytes; o += 512u)
    sts_v4(sGrp + o, make_uint4(0, 0, 0, 0));
  __syncwarp();

  uint32_t excl = incl - packed;
  uint32_t cumul = excl & 0xfffffu;
  uint32_t rank = excl >> 20; // present symbols before mine
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (freq[k]) {
      // symbol start bit
      atomicOr(reinterpret_cast<unsigned *>(smemWarp + L::kOffGrp) + (cumul >> 4), 1u << (cumul & 15u));
      // entry index = number of starts at or below any slot of this symbol = rank + 1
      const uint32_t sym = lane * 8u + (uint32_t)k;
      const uint32_t fm = freq[k] - (uint32_t)L::kSlots;             // x' = (x >> b) * fm + x - cumul
      const uint32_t w1 = ((0u - cumul) << 8) | sym;
      sts_u64(sEnt + (rank + 1u) * 8u, make_uint2(fm, w1));
      rank += 1;
      cumul += freq[k];
    }
  }
  __syncwarp();

  // per-group prefix of start counts; lane owns kGroups/32 consecutive groups
  constexpr int kPer = L::kGroups / 32;
  uint32_t local = 0;
  const uint32_t gBase = sGrp + lane * (uint32_t)kPer * 4u;
#pragma unroll 4
  for (int k = 0; k < kPer; k++)
    local += __popc(lds_u32(gBase + k * 4u));
  uint32_t scan = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t up = __shfl_up_sync(kFull, scan, d);
    if (lane >= (uint32_t)d)
      scan += up;
  }
  uint32_t before = scan - local;
#pragma unroll 4
  for (int k = 0; k < kPer; k++) {
    const uint32_t bm = lds_u32(gBase + k * 4u);
    sts_u32(gBase + k * 4u, ((before * 8u) << 16) | bm);
    before += __popc(bm);
  }
  __syncwarp();

  if constexpr (TK == TK_PACKED) {
    // expand to one u32 per slot: {freq:12 | slot - cumul:12 | symbol:8}
    const uint32_t sPk = smem_u32(smemWarp + L::kOffPacked);
    for (uint32_t slot = lane; slot < (uint32_t)L::kSlots; slot += 32u) {
      const uint32_t e = lds_u32(sGrp + ((slot >> 4) << 2));
      const uint32_t cnt = __popc(e << (31u - (slot & 15u)));
      const uint2 en = lds_u64(sEnt + (e >> 16) + cnt * 8u);
      const uint32_t f = en.x + (uint32_t)L::kSlots;
      const uint32_t bias = slot + (uint32_t)((int32_t)en.y >> 8);
      sts_u32(sPk + slot * 4u, (f << 20) | (bias << 8) | (en.y & 0xffu));
    }
    __syncwarp();
  }
  return true;
}

// true if some symbol owns the whole range (freq == 2^BITS): the packed 12-bit field cannot hold it
// (same limit as the reference's packed table, src/hist.cpp:304); callers route such tables to TK_RANK math.
template <int BITS, int N, int TK>
__device__ __forceinline__ bool table_is_degenerate(const uint8_t *smemWarp)
{
  using L = WarpLayout<BITS, N, TK>;
  const uint2 en = lds_u64(smem_u32(smemWarp + L::kOffEnt) + 8u);
  return en.x == 0u; // first present symbol has freq - 2^b == 0
}

// ---------------------------------------------------------------------------------------------- decode

template <int BITS, int N, int TK>
struct Decoder {
  using L = WarpLayout<BITS, N, TK>;

  uint32_t sGrp, sEnt, sPk;
  bool degenerate; // warp-uniform: packed lookups would be wrong, use the rank table

  __device__ __forceinline__ void init(uint8_t *smemWarp)
  {
    sGrp = smem_u32(smemWarp + L::kOffGrp);
    sEnt = smem_u32(smemWarp + L::kOffEnt);
    sPk = smem_u32(smemWarp + L::kOffPacked);
    degenerate = false;
  }

  // symbol lookup + state update for one state; returns the symbol in the low byte
  __device__ __forceinline__ uint32_t symbol_step_rank(uint32_t &x) const
  {
    const uint32_t e = lds_u32(sGrp + ((x >> 2) & (uint32_t)((L::kGroups - 1) << 2)));
    const uint32_t cnt = __popc(e << ((~x & 15u) | 16u)); // starts at or below the slot, inside its group
    const uint2 en = lds_u64(sEnt + (e >> 16) + cnt * 8u);
    x = (x >> BITS) * en.x + x;                           // (x >> b) * freq + slot, since en.x = freq - 2^b
    x += (uint32_t)((int32_t)en.y >> 8);                  // - cumul
    return en.y;
  }

  __device__ __forceinline__ uint32_t symbol_step_packed(uint32_t &x) const
  {
    const uint32_t e = lds_u32(sPk + ((x & (uint32_t)(L::kSlots - 1)) << 2));
    x = (x >> BITS) * (e >> 20) + ((e >> 8) & 0xfffu);
    return e;
  }

  __device__ __forceinline__ uint32_t symbol_step(uint32_t &x) const
  {
    if constexpr (TK == TK_PACKED)
      return symbol_step_packed(x);
    else
      return symbol_step_rank(x);
  }

  // renormalise the states of one half-row; `wordAddr` is the shared address of the warp cursor
  __device__ __forceinline__ void renorm(uint32_t &x, uint32_t &wordAddr, uint32_t &cur, bool active, uint32_t ltMask) const
  {
    const bool need = active && x < kConsumePoint16;
    const uint32_t m = __ballot_sync(kFull, need);
    if (need) {
      const uint32_t w = lds_u16(wordAddr + 2u * __popc(m & ltMask));
      x = (x << 16) | w;
    }
    const uint32_t adv = 2u * __popc(m);
    wordAddr += adv;
    cur += adv;
  }

  // full rows: `rows` rows of N symbols starting at out (already offset by the lane's byte position)
  template <bool kDegenerate>
  __device__ __forceinline__ void rows_impl(uint32_t &x0, uint32_t &x1, WordRing<L> &ring, uint8_t *outLane, uint64_t rows,
                                            uint32_t lane, uint32_t ltMask) const
  {
    for (uint64_t r = 0; r < rows; r++) {
      ring.advance_if_needed(lane);
      uint32_t wa = ring.cursor_addr();
      uint32_t cur = ring.cur;
      uint32_t s0, s1 = 0;
      if constexpr (kDegenerate || TK == TK_RANK) {
        s0 = symbol_step_rank(x0);
        if constexpr (N == 64)
          s1 = symbol_step_rank(x1);
      } else {
        s0 = symbol_step_packed(x0);
        if constexpr (N == 64)
          s1 = symbol_step_packed(x1);
      }
      st_global_u8(outLane, s0);
      renorm(x0, wa, cur, true, ltMask);
      if constexpr (N == 64) {
        st_global_u8(outLane + 32, s1);
        renorm(x1, wa, cur, true, ltMask);
      }
      ring.cur = cur;
      outLane += N;
    }
  }

  __device__ __forceinline__ void rows(uint32_t &x0, uint32_t &x1, WordRing<L> &ring, uint8_t *outLane, uint64_t nrows,
                                       uint32_t lane, uint32_t ltMask) const
  {
    if constexpr (TK == TK_PACKED) {
      if (degenerate) {
        rows_impl<true>(x0, x1, ring, outLane, nrows, lane, ltMask);
        return;
      }
    }
    rows_impl<false>(x0, x1, ring, outLane, nrows, lane, ltMask);
  }

  // the < N leftover symbols (src/rANS32x32_16w.cpp:238-266): lanes whose byte position is inside the buffer
  __device__ __forceinline__ void tail(uint32_t &x0, uint32_t &x1, WordRing<L> &ring, uint8_t *outLane, uint32_t lanePos,
                                       uint32_t left, uint32_t lane, uint32_t ltMask) const
  {
    ring.advance_if_needed(lane);
    uint32_t wa = ring.cursor_addr();
    uint32_t cur = ring.cur;
    const bool a0 = lanePos < left;
    uint32_t t0 = x0;
    const uint32_t s0 = symbol_step_rank(t0); // the rank table is always built
    if (a0) {
      x0 = t0;
      st_global_u8(outLane, s0);
    }
    renorm(x0, wa, cur, a0, ltMask);
    if constexpr (N == 64) {
      const bool a1 = lanePos + 32u < left;
      uint32_t t1 = x1;
      const uint32_t s1 = symbol_step_rank(t1);
      if (a1) {
        x1 = t1;
        st_global_u8(outLane + 32, s1);
      }
      renorm(x1, wa, cur, a1, ltMask);
    }
    ring.cur = cur;
  }
};

// ---------------------------------------------------------------------------------------------- fills

// single-symbol runs (src/block_rANS32x32_16w_decode.cpp:58-66): memset by one warp
__device__ __forceinline__ void warp_fill(uint8_t *dst, uint32_t symbol, uint64_t count, uint32_t lane)
{
  const uint32_t v = symbol * 0x01010101u;
  uint64_t head = (16u - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
  if (head > count)
    head = count;
  for (uint64_t i = lane; i < head; i += 32)
    dst[i] = (uint8_t)symbol;
  uint8_t *body = dst + head;
  const uint64_t vecs = (count - head) / 16;
  uint4 *b4 = reinterpret_cast<uint4 *>(body);
  for (uint64_t i = lane; i < vecs; i += 32)
    b4[i] = make_uint4(v, v, v, v);
  for (uint64_t i = head + vecs * 16 + lane; i < count; i += 32)
    dst[i] = (uint8_t)symbol;
}

} // namespace hsr
