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
// One CTA = one warp, so every table lives at a compile-time shared-memory address and lookups are
// LDS [reg + imm]. Shared-memory tables (private layouts; only the decoded bytes have to match the reference):
//   TK_RANK   bitmap-rank table, any bits:
//               grp[2^b / 16]  u32  {symbol starts before this group : 16 | start bitmap of its 16 slots : 16}
//               ent[256]       u32  per PRESENT symbol, in slot order {-cumul : 16 | freq - 2^b : 16}, both signed
//               sym[256]       u8   rank -> symbol (skipped when all 256 symbols are present: rank == symbol)
//             2^(b-2) + 1.25 KB (9.25 KB at 15 bits vs 33 KB for the reference's hist_dec2_t, src/hist.h:32-37)
//             and O(256 + 2^b/16) to build instead of O(2^b).
//   TK_PACKED one u32 per slot {freq : 12 | symbol : 8 | slot - cumul : 12}, bits <= 12, one lookup on the chain
//             (the reference's hist_dec_pack_t idea, src/hist.h:46-50, with the bias pre-subtracted).
//   TK_WIDE   bits >= 13, for launches with at most one unit per SM (a single raw / block_ stream, a handful of huge
//             mt_ blocks): one u32 per slot {freq : 16 | slot - cumul : 16} plus one u8 per slot (symbol) — two
//             INDEPENDENT lookups, one on the chain (the reference's hist_dec3_t idea, src/hist.h:39-44). 5 * 2^b
//             bytes = 160 KB at 15 bits: dynamic shared memory, one CTA per SM, which is all such a launch can use.
// Compressed words are staged by cp.async (LDGSTS, 16 B per lane) into a ring of overlapping linear segments, so
// the data-dependent word reads are LDS, never exposed DRAM latency.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/hsrans_b200.h"

// Output path experiment: 1 = stage decoded bytes in a 512-byte shared-memory tile and flush with 16-byte vector
// stores; 0 = one st.global.u8 per lane and half-row (32 lanes = one full 32-byte sector). Measured on B200
// (profiles/r1/): see the note at rows_impl.
#ifndef HSR_OUT_TILE
#define HSR_OUT_TILE 0
#endif

// Build-time knobs of the hot loop (defaults = what measured best on B200, profiles/r1/variants_ring_unroll.jsonl):
//   HSR_ROW_UNROLL  rows per trip of the hot loop: the loop counter, the branch and the 64-bit output pointer
//                   amortise over 4 rows (2 -> 4: +1.8 %, 8: no further gain)
//   HSR_RING2       1: two-buffer word ring with incremental addressing (512 B less shared memory per CTA = 20
//                   instead of 19 resident CTAs per SM at 15 bits, 24 instead of 31 instructions per refill: +6.3 %)
//                   0: the three-buffer ring with modulo addressing it replaced
#ifndef HSR_ROW_UNROLL
#define HSR_ROW_UNROLL 4
#endif
#ifndef HSR_RING2
#define HSR_RING2 1
#endif
//   HSR_GRP24       1: at 15 bits and N = 32 the bitmap-rank groups cover 24 slots instead of 16 ({8-bit prefix | 24-bit
//                   bitmap} still fills one u32): the group table shrinks from 8 KB to 5.5 KB, so 26 instead of 20
//                   one-warp CTAs fit an SM, for a multiply-shift division by 24 on the lookup path (symbol_step_rank).
//                   N = 32 has one state per lane and is short of warps, not of pipe cycles: 660 -> 740 GB/s (723 -> 816
//                   with overlapping launches). N = 64 (two states per lane, data-pipe-bound) gains nothing from the
//                   extra warps and pays for the extra instructions (917 -> 839, 956 -> 964 overlapped), so it keeps
//                   16-slot groups (profiles/r2/variants_grp24.jsonl).
#ifndef HSR_GRP24
#define HSR_GRP24 1
#endif
//   HSR_RING_TMA    1 (with HSR_RING2=0): the word ring is fed by TMA bulk copies (measured slower, see WordRingTma)
#ifndef HSR_RING_TMA
#define HSR_RING_TMA 0
#endif
//   HSR_QUAD_STORE  1: four half-rows of decoded bytes leave as one 32-bit store per lane after a 4x4 byte transpose
//                   inside each lane quad (see rows_impl); 0: one byte store per lane and half-row
#ifndef HSR_QUAD_STORE
#define HSR_QUAD_STORE 0
#endif

namespace hsr {

constexpr uint32_t kConsumePoint16 = 1u << 15; // src/rans.h:8
constexpr unsigned kFull = 0xffffffffu;
constexpr int kRowUnroll = HSR_ROW_UNROLL;

enum TableKind : int { TK_RANK = 1, TK_PACKED = 2, TK_WIDE = 3 };

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

// The CTA's shared memory is declared in PTX, not C++: `mov.u32 r, hsr_smem_arr` is then a link-time constant,
// ptxas keeps it in a uniform register and every table lookup becomes LDS [reg + UR + imm] with no address
// arithmetic. (A C++ __shared__ array reached through generic pointers drags the cluster shared-window base,
// S2R SR_CgaCtaId + LEA, and one extra IADD per lookup through the hot loop.) Call exactly once per kernel.
template <int BYTES>
__device__ __forceinline__ uint32_t declare_smem()
{
  uint32_t base;
  asm volatile(".shared .align 16 .b8 hsr_smem_arr[%1];\n\tmov.u32 %0, hsr_smem_arr;" : "=r"(base) : "n"(BYTES));
  return base;
}
// kernels whose tables exceed the 48 KB static limit take the whole CTA allocation as dynamic shared memory
__device__ __forceinline__ uint32_t dynamic_smem_base()
{
  extern __shared__ __align__(16) uint8_t hsr_dynamic_smem[];
  return (uint32_t)__cvta_generic_to_shared(hsr_dynamic_smem);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void atoms_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t srcBytes)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(srcBytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int Pending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(Pending) : "memory"); }

// mbarrier + TMA bulk copy (cp.async.bulk -> UBLKCP): one elected lane moves a whole ring segment
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ uint4 lds_v4(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void st_global_v4(uint8_t *p, uint4 v) { asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void st_global_u32(uint8_t *p, uint32_t v) { asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_global_u8(uint8_t *p, uint32_t v) { asm volatile("st.global.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// ---------------------------------------------------------------------------------------------- layout

template <int BITS, int N, int TK>
struct WarpLayout {
  static_assert(BITS >= 10 && BITS <= 15, "probability bits 10..15");
  static_assert(N == 32 || N == 64, "32 or 64 interleaved states");
  static_assert(TK == TK_RANK || (TK == TK_PACKED && BITS <= 12) || (TK == TK_WIDE && BITS >= 13),
                "packed slot table up to 12 bits, wide slot tables from 13 bits");

  static constexpr int kSlots = 1 << BITS;
  // bitmap-rank groups: one u32 {symbol starts before the group | start bitmap of its slots} per kGrpSlots slots.
  // 16 slots: {16 | 16}; 24 slots (15 bits, N = 32): {8 | 24} — at most 255 starts precede a group (slot 0's is implicit)
  static constexpr int kGrpSlots = (HSR_GRP24 && BITS == 15 && N == 32) ? 24 : 16;
  static constexpr int kPrefixShift = 32 - (kGrpSlots == 24 ? 8 : 16);
  static constexpr int kGroupsUsed = (kSlots + kGrpSlots - 1) / kGrpSlots;
  static constexpr int kGroups = kGrpSlots == 16 ? kGroupsUsed : (kGroupsUsed + 127) / 128 * 128; // scan granularity
  static constexpr int kGrpBytes = kGroups * 4;
  // group of a slot; 43691 / 2^20 = 1 / 23.99983: exact floor(slot / 24) for every slot < 2^15
  static __device__ __forceinline__ uint32_t grp_of(uint32_t slot) { return kGrpSlots == 16 ? slot >> 4 : (slot * 43691u) >> 20; }
  static constexpr int kEntBytes = 256 * 4;
  static constexpr int kSymBytes = 256;
  static constexpr int kPackedBytes = TK == TK_PACKED ? kSlots * 4 : 0;
  static constexpr int kWideBytes = TK == TK_WIDE ? kSlots * 5 : 0; // u32 {freq, bias} per slot, then u8 symbol per slot
  static constexpr bool kDynamic = TK == TK_WIDE;

  // word ring: kBufs linear segments of kSeg bytes; consecutive segments overlap by one worst-case row
  static constexpr int kSeg = 512;                  // one 16-byte cp.async per lane
  static constexpr int kBufs = HSR_RING2 ? 2 : 3;
  static constexpr int kOverlap = 2 * N;            // a row consumes at most N words
  static constexpr int kStride = kSeg - kOverlap;
  static constexpr int kRingBytes = kSeg * kBufs;

  static constexpr int kOffGrp = 0;
  static constexpr int kOffEnt = kOffGrp + kGrpBytes;
  static constexpr int kOffSym = kOffEnt + kEntBytes;
  static constexpr int kOffPacked = kOffSym + kSymBytes;
  static constexpr int kOffWide = kOffPacked + kPackedBytes;
  static constexpr int kOffWideSym = kOffWide + (TK == TK_WIDE ? kSlots * 4 : 0);
  static constexpr int kOffRing = kOffWide + kWideBytes;
  static constexpr int kOffBar = kOffRing + kRingBytes;   // one mbarrier per ring buffer (TMA bulk copies only)
  static constexpr int kOffTile = kOffBar + (HSR_RING_TMA ? 32 : 0); // output staging tile (HSR_OUT_TILE experiments)
  static constexpr int kTileBytes = HSR_OUT_TILE ? 512 : 0;
  static constexpr int kBytes = kOffTile + kTileBytes;    // per warp (= per CTA), multiple of 16
  static_assert(kBytes % 16 == 0 && kBytes <= (kDynamic ? 227 : 48) * 1024, "shared memory budget");
};

// ---------------------------------------------------------------------------------------------- word ring

// Segment k holds stream bytes [k*kStride, k*kStride + kSeg) relative to `gbase` (16-byte aligned, at or just
// below the first word). A row may read up to kOverlap bytes past the cursor, so the decoder leaves segment k
// once cursor - k*kStride >= kStride, which is exactly where segment k+1 begins.
template <class L>
struct WordRing {
  uint32_t sbuf;        // shared address of buffer 0 (constant)
  const uint8_t *gbase; // 16-byte aligned
  uint32_t glimit;      // readable bytes from gbase (never read past the caller's inLength)
  uint32_t seg;         // current segment
  uint32_t wp;          // shared address of the word cursor (inside segment `seg`'s buffer)
  uint32_t wlimit;      // leave the segment once wp reaches this (buffer start + kStride)

  __device__ __forceinline__ uint32_t buf(uint32_t k) const { return sbuf + (k % L::kBufs) * L::kSeg; }

  __device__ __forceinline__ void issue(uint32_t k, uint32_t lane) const
  {
    const uint32_t srcOff = k * L::kStride + lane * 16u;
    uint32_t bytes = srcOff < glimit ? glimit - srcOff : 0u;
    bytes = bytes > 16u ? 16u : bytes;
    cp_async16(buf(k) + lane * 16u, gbase + (bytes ? srcOff : 0u), bytes); // bytes == 0: zero fill only
    cp_async_commit();
  }

  __device__ __forceinline__ void start(uint32_t sbuf_, const uint8_t *firstWord, const uint8_t *streamEnd, uint32_t lane)
  {
    sbuf = sbuf_;
    const uintptr_t a = reinterpret_cast<uintptr_t>(firstWord);
    gbase = reinterpret_cast<const uint8_t *>(a & ~(uintptr_t)15);
    const uint64_t avail = (uint64_t)(streamEnd - gbase);
    glimit = avail > 0xffffffffull ? 0xffffffffu : (uint32_t)avail;
    seg = 0;
    wp = sbuf + (uint32_t)(a & 15);
    wlimit = sbuf + L::kStride;
    __syncwarp(); // every lane is done with whatever lived in the ring before
#pragma unroll
    for (uint32_t k = 0; k + 2 <= (uint32_t)L::kBufs; k++)
      issue(k, lane);
    cp_async_wait<L::kBufs - 2>();
    __syncwarp();
  }

  __device__ __forceinline__ void start_wait() const {}

  // call once per row, before any word of the row is read
  __device__ __forceinline__ void advance_if_needed(uint32_t lane)
  {
    if (wp >= wlimit) {
      const uint32_t into = wp - wlimit; // offset inside the next segment
      seg += 1;
      issue(seg + L::kBufs - 2, lane); // lands in the buffer of segment seg-2, abandoned one segment ago
      cp_async_wait<L::kBufs - 2>();   // segment `seg` has landed for this lane ...
      __syncwarp();                    // ... and for all the others
      const uint32_t b = buf(seg);
      wp = b + into;
      wlimit = b + L::kStride;
    }
  }

  // bytes consumed from gbase so far
  __device__ __forceinline__ uint32_t cursor() const { return seg * L::kStride + (wp - buf(seg)); }

  __device__ __forceinline__ void drain() const { cp_async_wait<0>(); }
};

// Two-buffer ring: while the warp reads segment k from one buffer, segment k + 1 is in flight into the other —
// the buffer the warp left when it crossed into k. The LDS.U16 of the last row read there were issued before the
// LDGSTS that refills it and have long returned when its data arrives. All addressing is incremental (no modulo):
// per lane a source offset and a destination that flips between the two buffers.
template <class L>
struct WordRing2 {
  const uint8_t *gbase; // 16-byte aligned
  uint32_t glimit;      // readable bytes from gbase
  uint32_t srcOff;      // this lane's source offset of the next segment to issue
  uint32_t dstNext;     // this lane's shared destination of the next segment to issue
  uint32_t dstSum;      // dst(buffer 0) + dst(buffer 1)
  uint32_t cur;         // shared address of the buffer being read
  uint32_t curSum;      // buffer 0 + buffer 1
  uint32_t base;        // stream bytes (from gbase) in front of the current buffer
  uint32_t wp, wlimit;

  __device__ __forceinline__ void issue()
  {
    uint32_t bytes = srcOff < glimit ? glimit - srcOff : 0u;
    bytes = bytes > 16u ? 16u : bytes;
    cp_async16(dstNext, gbase + (bytes ? srcOff : 0u), bytes); // bytes == 0: zero fill only
    cp_async_commit();
    srcOff += (uint32_t)L::kStride;
    dstNext = dstSum - dstNext;
  }

  __device__ __forceinline__ void start(uint32_t sbuf, const uint8_t *firstWord, const uint8_t *streamEnd, uint32_t lane)
  {
    const uintptr_t a = reinterpret_cast<uintptr_t>(firstWord);
    gbase = reinterpret_cast<const uint8_t *>(a & ~(uintptr_t)15);
    const uint64_t avail = (uint64_t)(streamEnd - gbase);
    glimit = avail > 0xffffffffull ? 0xffffffffu : (uint32_t)avail;
    srcOff = lane * 16u;
    dstNext = sbuf + lane * 16u;
    dstSum = 2u * dstNext + (uint32_t)L::kSeg;
    cur = sbuf;
    curSum = 2u * sbuf + (uint32_t)L::kSeg;
    base = 0;
    wp = sbuf + (uint32_t)(a & 15);
    wlimit = sbuf + L::kStride;
    __syncwarp(); // every lane is done with whatever lived in the ring before
    issue();
    issue();
  }
  // start() only posts the first two segments; call this before the first word is read (the table build of the
  // block sits between the two, so its global loads and the ring's overlap)
  __device__ __forceinline__ void start_wait()
  {
    cp_async_wait<1>();
    __syncwarp();
  }

  __device__ __forceinline__ void advance_if_needed(uint32_t)
  {
    if (wp >= wlimit) {
      const uint32_t into = wp - wlimit; // offset inside the next segment
      __syncwarp();
      issue();              // two segments ahead of the one just left, into its buffer
      cp_async_wait<1>();   // the segment being entered has landed for this lane ...
      __syncwarp();         // ... and for all the others
      cur = curSum - cur;
      base += (uint32_t)L::kStride;
      wp = cur + into;
      wlimit = cur + L::kStride;
    }
  }

  __device__ __forceinline__ uint32_t cursor() const { return base + (wp - cur); }
  __device__ __forceinline__ void drain() const { cp_async_wait<0>(); }
};

// The same ring fed by the TMA: a full segment is one cp.async.bulk (UBLKCP) issued by lane 0 and tracked by an
// mbarrier per buffer; only the last, partial segment of a block still goes through the clamped, zero-filling
// LDGSTS path above so that nothing past the block's end is ever read. Segments are numbered globally across the
// blocks a warp decodes (buffer = g % kBufs, barrier phase parity = (g / kBufs) & 1), and a new block first waits
// for whatever the previous one left in flight, so a barrier is never re-armed before its phase completed.
template <class L>
struct WordRingTma {
  uint32_t sbuf, sbar;
  const uint8_t *gbase;
  uint32_t glimit;
  uint32_t seg;      // current segment of this block
  uint32_t g0;       // global number of this block's segment 0
  uint32_t gIssued;  // next global segment number to issue
  uint32_t gWaited;  // every global segment below this has been waited for
  uint32_t wp, wlimit;
  uint32_t stuck;    // set if a barrier never completed (reported, never hangs)

  __device__ __forceinline__ uint32_t buf(uint32_t g) const { return sbuf + (g % L::kBufs) * L::kSeg; }
  __device__ __forceinline__ uint32_t bar(uint32_t g) const { return sbar + (g % L::kBufs) * 8u; }

  __device__ __forceinline__ void init(uint32_t sbuf_, uint32_t sbar_, uint32_t lane)
  {
    sbuf = sbuf_; sbar = sbar_;
    gIssued = gWaited = 0; stuck = 0;
    if (lane == 0) {
#pragma unroll
      for (uint32_t b = 0; b < (uint32_t)L::kBufs; b++) mbar_init(sbar + b * 8u, 1u);
      fence_mbar_init();
    }
    __syncwarp();
  }

  __device__ __forceinline__ void issue(uint32_t lane)
  {
    const uint32_t g = gIssued++;
    const uint32_t srcOff = (g - g0) * L::kStride;
    if (srcOff + (uint32_t)L::kSeg <= glimit) {
      if (lane == 0) {
        mbar_arrive_expect_tx(bar(g), (uint32_t)L::kSeg);
        bulk_copy_g2s(buf(g), gbase + srcOff, (uint32_t)L::kSeg, bar(g));
      }
    } else { // partial or empty segment at the end of the block: clamped 16-byte copies, zero fill beyond the limit
      const uint32_t off = srcOff + lane * 16u;
      uint32_t bytes = off < glimit ? glimit - off : 0u;
      bytes = bytes > 16u ? 16u : bytes;
      cp_async16(buf(g) + lane * 16u, gbase + (bytes ? off : 0u), bytes);
      cp_async_commit();
      cp_async_wait<0>();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(g));
    }
  }

  __device__ __forceinline__ void wait_next()
  {
    const uint32_t g = gWaited++;
    const uint32_t b = bar(g), parity = (g / L::kBufs) & 1u;
    uint32_t spins = 0;
    while (!mbar_try_wait(b, parity)) {
      if (++spins > (1u << 22)) { stuck = 1; break; }
    }
  }

  __device__ __forceinline__ void start(const uint8_t *firstWord, const uint8_t *streamEnd, uint32_t lane)
  {
    while (gWaited < gIssued) wait_next(); // segments the previous block prefetched and never used
    __syncwarp();                          // every lane is done with whatever lived in the ring before
    const uintptr_t a = reinterpret_cast<uintptr_t>(firstWord);
    gbase = reinterpret_cast<const uint8_t *>(a & ~(uintptr_t)15);
    const uint64_t avail = (uint64_t)(streamEnd - gbase);
    glimit = avail > 0xffffffffull ? 0xffffffffu : (uint32_t)avail;
    seg = 0;
    g0 = gIssued;
#pragma unroll
    for (uint32_t k = 0; k + 2 <= (uint32_t)L::kBufs; k++)
      issue(lane);
    wait_next();
    wp = buf(g0) + (uint32_t)(a & 15);
    wlimit = buf(g0) + L::kStride;
  }

  __device__ __forceinline__ void start_wait() const {}

  __device__ __forceinline__ void advance_if_needed(uint32_t lane)
  {
    if (wp >= wlimit) {
      const uint32_t into = wp - wlimit;
      seg += 1;
      issue(lane);  // segment seg + kBufs - 2 lands in the buffer of segment seg - 2, abandoned one segment ago
      wait_next();  // segment seg
      const uint32_t b = buf(g0 + seg);
      wp = b + into;
      wlimit = b + L::kStride;
    }
  }

  __device__ __forceinline__ uint32_t cursor() const { return seg * L::kStride + (wp - buf(g0 + seg)); }

  __device__ __forceinline__ void drain()
  {
    while (gWaited < gIssued) wait_next();
  }
};

// Measured on B200 (profiles/r1/sweep_tma_ring_experiment.jsonl): the TMA ring is correct (all parity tests pass) but
// 26-31 % SLOWER than the LDGSTS ring (571 vs 773 GB/s at mt_64x15, 838 vs 1220 at 10 bits): a 512-byte segment is
// too small to amortise the mbarrier try_wait round trip per segment, and the extra live state costs registers
// (92 in the packed 10-bit kernel). A leaner two-buffer TMA ring (one pending bulk copy per warp, 16-byte granular
// sizes) measured 726 GB/s against 903 for the two-buffer LDGSTS ring (profiles/r1/variants_ring_unroll.jsonl): small
// bulk copies at ~3000 warps x 1 per 1.5 us are simply not what the TMA is good at. It stays available
// (-DHSR_RING_TMA=1 -DHSR_RING2=0) for larger-segment experiments; LDGSTS is the default.
template <class L>
#if HSR_RING_TMA
using Ring = WordRingTma<L>;
#elif HSR_RING2
using Ring = WordRing2<L>;
#else
using Ring = WordRing<L>;
#endif

// ---------------------------------------------------------------------------------------------- table build

struct TableInfo {
  bool ok;          // counts sum to 2^BITS (src/hist.cpp:308-324)
  bool allPresent;  // every symbol has a non-zero count: rank == symbol
  bool degenerate;  // one symbol owns the whole range (freq == 2^BITS)
};

// rank of the symbol that owns `slot` (table expansion; the hot loop has its own fused form in symbol_step_rank)
template <class L>
__device__ __forceinline__ uint32_t rank_of_slot(uint32_t sGrp, uint32_t slot)
{
  const uint32_t g = L::grp_of(slot);
  const uint32_t w = lds_u32(sGrp + (g << 2));
  const uint32_t pos = slot - g * (uint32_t)L::kGrpSlots;
  return (w >> L::kPrefixShift) + __popc(w << (31u - pos)); // starts at or below the slot
}

// Builds the warp's tables from 256 u16 counts at `counts` (2-byte aligned global memory). Warp-uniform result.
template <int BITS, int N, int TK>
__device__ __forceinline__ TableInfo build_tables(uint32_t smemWarp, const uint8_t *counts, uint32_t lane)
{
  using L = WarpLayout<BITS, N, TK>;
  const uint32_t sGrp = smemWarp + L::kOffGrp;
  const uint32_t sEnt = smemWarp + L::kOffEnt;
  const uint32_t sSym = smemWarp + L::kOffSym;
  TableInfo info{false, false, false};

  // lane l owns symbols 8l .. 8l+7
  uint32_t freq[8];
  uint32_t sum = 0, present = 0, big = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    freq[k] = ldg_u16(counts + 2 * (lane * 8 + k));
    sum += freq[k];
    present += freq[k] != 0;
    big |= freq[k] == (uint32_t)L::kSlots;
  }
  // one inclusive scan carries both running totals: frequencies (low 20 bits) and present-symbol counts (high 12)
  const uint32_t packed = sum | (present << 20);
  uint32_t incl = packed;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t up = __shfl_up_sync(kFull, incl, d);
    if (lane >= (uint32_t)d)
      incl += up;
  }
  const uint32_t last = __shfl_sync(kFull, incl, 31);
  // a corrupt histogram can carry up to 256 * 65535; per-lane sums above 2^BITS are rejected first so the
  // 20-bit field cannot overflow unnoticed (32 * 2^15 = 2^20 wraps to 0 != 2^BITS)
  const bool laneOk = sum <= (uint32_t)L::kSlots;
  if (!__all_sync(kFull, laneOk) || (last & 0xfffffu) != (uint32_t)L::kSlots)
    return info;
  info.ok = true;
  info.allPresent = (last >> 20) == 256u;
  info.degenerate = __any_sync(kFull, big != 0);

  __syncwarp();
  for (uint32_t o = lane * 16u; o < (uint32_t)L::kGrpBytes; o += 512u) // clear the bitmap groups
    sts_v4(sGrp + o, make_uint4(0, 0, 0, 0));
  __syncwarp();

  const uint32_t excl = incl - packed;
  uint32_t cumul = excl & 0xfffffu;
  uint32_t rank = excl >> 20; // present symbols before mine
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (freq[k]) {
      if (cumul) { // the start at slot 0 is implicit, so that (starts before or at a slot) == rank of its symbol
        const uint32_t g = L::grp_of(cumul);
        atoms_or(sGrp + (g << 2), 1u << (cumul - g * (uint32_t)L::kGrpSlots));
      }
      // x' = x + (x >> b) * (freq - 2^b) - cumul; both fields are signed 16-bit
      const uint32_t e = (((0u - cumul) & 0xffffu) << 16) | ((freq[k] - (uint32_t)L::kSlots) & 0xffffu);
      sts_u32(sEnt + rank * 4u, e);
      sts_u8(sSym + rank, lane * 8u + (uint32_t)k);
      rank += 1;
      cumul += freq[k];
    }
  }
  __syncwarp();

  // prefix of start counts over the groups. 128 groups per step where the table has that many: lane l owns groups
  // 4l .. 4l+3 (one conflict-free 16-byte load and store), so one warp scan serves four groups per lane — the
  // shuffles of that scan share the shared-memory data pipe with every lookup of the other resident warps.
  uint32_t carry = 0;
  if constexpr (L::kGroups >= 128) {
#pragma unroll 2
    for (uint32_t g0 = 0; g0 < (uint32_t)L::kGroups; g0 += 128u) {
      const uint32_t a = sGrp + (g0 + 4u * lane) * 4u;
      const uint4 bm = lds_v4(a);
      const uint32_t c0 = __popc(bm.x), c1 = __popc(bm.y), c2 = __popc(bm.z), c3 = __popc(bm.w);
      const uint32_t tot = c0 + c1 + c2 + c3;
      uint32_t scan = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(kFull, scan, d);
        if (lane >= (uint32_t)d)
          scan += up;
      }
      const uint32_t e0 = carry + scan - tot, e1 = e0 + c0, e2 = e1 + c1, e3 = e2 + c2;
      constexpr int kP = L::kPrefixShift;
      sts_v4(a, make_uint4((e0 << kP) | bm.x, (e1 << kP) | bm.y, (e2 << kP) | bm.z, (e3 << kP) | bm.w));
      carry += __shfl_sync(kFull, scan, 31);
    }
  } else
#pragma unroll 2
  for (uint32_t g0 = 0; g0 < (uint32_t)L::kGroups; g0 += 32u) {
    const uint32_t a = sGrp + (g0 + lane) * 4u;
    const uint32_t bm = lds_u32(a);
    const uint32_t c = __popc(bm);
    uint32_t scan = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(kFull, scan, d);
      if (lane >= (uint32_t)d)
        scan += up;
    }
    sts_u32(a, ((carry + scan - c) << L::kPrefixShift) | bm);
    carry += __shfl_sync(kFull, scan, 31);
  }
  __syncwarp();

  if constexpr (TK == TK_PACKED) {
    // expand to one u32 per slot: {freq:12 | symbol:8 | slot - cumul:12}
    const uint32_t sPk = smemWarp + L::kOffPacked;
    for (uint32_t slot = lane; slot < (uint32_t)L::kSlots; slot += 32u) {
      const uint32_t rank = rank_of_slot<L>(sGrp, slot);
      const uint32_t e = lds_u32(sEnt + rank * 4u);
      const uint32_t s = lds_u8(sSym + rank);
      const uint32_t f = (uint32_t)L::kSlots + (uint32_t)(int32_t)(int16_t)(e & 0xffffu);
      const uint32_t bias = (slot + (uint32_t)((int32_t)e >> 16)) & 0xfffu;
      sts_u32(sPk + slot * 4u, (f << 20) | (s << 12) | bias);
    }
    __syncwarp();
  }
  if constexpr (TK == TK_WIDE) {
    // expand to {freq:16 | slot - cumul:16} and the symbol per slot
    const uint32_t sW = smemWarp + L::kOffWide, sWs = smemWarp + L::kOffWideSym;
    for (uint32_t slot = lane; slot < (uint32_t)L::kSlots; slot += 32u) {
      const uint32_t rank = rank_of_slot<L>(sGrp, slot);
      const uint32_t e = lds_u32(sEnt + rank * 4u);
      const uint32_t f = (uint32_t)L::kSlots + (uint32_t)(int32_t)(int16_t)(e & 0xffffu);
      const uint32_t bias = (slot + (uint32_t)((int32_t)e >> 16)) & 0xffffu;
      sts_u32(sW + slot * 4u, (f << 16) | bias);
      sts_u8(sWs + slot, lds_u8(sSym + rank));
    }
    __syncwarp();
  }
  return info;
}

// ---------------------------------------------------------------------------------------------- decode

template <int BITS, int N, int TK>
struct Decoder {
  using L = WarpLayout<BITS, N, TK>;

  uint32_t sGrp, sEnt, sSym, sPk, sW, sWs, sTile; // shared addresses, compile-time constants after inlining

  __device__ __forceinline__ void init(uint32_t smemWarp)
  {
    sTile = smemWarp + L::kOffTile;
    sGrp = smemWarp + L::kOffGrp;
    sEnt = smemWarp + L::kOffEnt;
    sSym = smemWarp + L::kOffSym;
    sPk = smemWarp + L::kOffPacked;
    sW = smemWarp + L::kOffWide;
    sWs = smemWarp + L::kOffWideSym;
  }

  // symbol lookup + state update for one state; returns the symbol (low byte significant)
  template <bool kAllPresent>
  __device__ __forceinline__ uint32_t symbol_step_rank(uint32_t &x) const
  {
    uint32_t rank; // symbol starts at or below the slot = rank of its owner
    if constexpr (L::kGrpSlots == 16) {
      const uint32_t g = lds_u32(sGrp + ((x >> 2) & (uint32_t)((L::kGroups - 1) << 2)));
      uint32_t sh; // 31 - (x & 15) = (~x & 15) | 16 as ONE lop3 (immLut 0xAE = (~a & b) | c)
      asm("lop3.b32 %0, %1, 15, 16, 0xAE;" : "=r"(sh) : "r"(x));
      rank = (g >> 16) + __popc(g << sh);
    } else {
      // 24-slot groups: gi = slot / 24 by multiply-shift; the bit position is pos = slot - 24 gi, and the shift that
      // brings bit `pos` to bit 31, 31 - pos, equals (24 gi + ~x) mod 32 (x = slot mod 32) — one multiply-add feeding a
      // wrapping funnel shift, no subtraction, no mask
      const uint32_t gi = ((x & (uint32_t)(L::kSlots - 1)) * 43691u) >> 20;
      const uint32_t g = lds_u32(sGrp + (gi << 2));
      const uint32_t sh = gi * 24u + ~x;
      rank = (g >> 24) + __popc(__funnelshift_l(0u, g, sh));
    }
    const uint32_t e = lds_u32(sEnt + rank * 4u);
    x = (x >> BITS) * (uint32_t)(int32_t)(int16_t)(e & 0xffffu) + x; // (x >> b) * freq + slot
    x += (uint32_t)((int32_t)e >> 16);                               // - cumul
    if constexpr (kAllPresent)
      return rank;
    else
      return lds_u8(sSym + rank);
  }

  __device__ __forceinline__ uint32_t symbol_step_packed(uint32_t &x) const
  {
    const uint32_t e = lds_u32(sPk + ((x & (uint32_t)(L::kSlots - 1)) << 2));
    x = (x >> BITS) * (e >> 20) + (e & 0xfffu);
    return e >> 12;
  }

  __device__ __forceinline__ uint32_t symbol_step_wide(uint32_t &x) const
  {
    const uint32_t slot = x & (uint32_t)(L::kSlots - 1);
    const uint32_t e = lds_u32(sW + (slot << 2));
    const uint32_t s = lds_u8(sWs + slot); // off the chain
    x = (x >> BITS) * (e >> 16) + (e & 0xffffu);
    return s;
  }

  // kMode 0: packed table, 1: rank table with all symbols present, 2: rank table with the rank->symbol map, 3: wide
  template <int kMode>
  __device__ __forceinline__ uint32_t symbol_step(uint32_t &x) const
  {
    if constexpr (kMode == 0)
      return symbol_step_packed(x);
    else if constexpr (kMode == 3)
      return symbol_step_wide(x);
    else
      return symbol_step_rank<kMode == 1>(x);
  }

  // Renormalise one full half-row; `wp` is the shared address of the warp cursor in the word ring. Written in
  // PTX so that exactly one predicate feeds the ballot, the word load and the merge:
  //   p = x < 2^15; m = ballot(p); w = words[wp + 2 * popc(m & lanemask_lt)]; if (p) x = x << 16 | w; wp += 2 * popc(m)
  __device__ __forceinline__ void renorm(uint32_t &x, uint32_t &wp, uint32_t ltMask) const
  {
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 ".reg .u32 m, r, w, c;\n\t"
                 "setp.lt.u32 p, %0, 32768;\n\t"
                 "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
                 "and.b32 r, m, %2;\n\t"
                 "popc.b32 r, r;\n\t"
                 "mad.lo.u32 r, r, 2, %1;\n\t"
                 "ld.shared.u16 w, [r];\n\t"
                 "@p mad.lo.u32 %0, %0, 65536, w;\n\t"
                 "popc.b32 c, m;\n\t"
                 "mad.lo.u32 %1, c, 2, %1;\n\t"
                 "}"
                 : "+r"(x), "+r"(wp)
                 : "r"(ltMask)
                 : "memory");
  }

  // the same for the ragged last row: only `active` lanes hold a symbol
  __device__ __forceinline__ void renorm_masked(uint32_t &x, uint32_t &wp, bool active, uint32_t ltMask) const
  {
    const bool need = active && x < kConsumePoint16;
    const uint32_t m = __ballot_sync(kFull, need);
    const uint32_t w = lds_u16(wp + 2u * __popc(m & ltMask));
    x = need ? ((x << 16) | w) : x;
    wp += 2u * __popc(m);
  }

  // mode 0: packed table, 1: rank table with all symbols present, 2: rank table with the rank->symbol map
  template <int kMode>
  __device__ __forceinline__ void rows_impl(uint32_t &x0, uint32_t &x1, Ring<L> &ring, uint8_t *outLane, uint32_t rows,
                                            uint32_t lane, uint32_t ltMask) const
  {
    uint32_t r = 0;
#if HSR_OUT_TILE
    // tile variant: kTileRows rows are staged in shared memory, then flushed as 32 x 16-byte stores
    constexpr uint32_t kTileRows = 512 / N;
    if ((reinterpret_cast<uintptr_t>(outLane - idx2idx_lane(lane)) & 15) == 0) {
      const uint32_t lanePos = idx2idx_lane(lane);
      for (; r + kTileRows <= rows; r += kTileRows) {
#pragma unroll 2
        for (uint32_t t = 0; t < kTileRows; t++) {
          ring.advance_if_needed(lane);
          uint32_t s0, s1 = 0;
          if constexpr (kMode == 0) {
            s0 = symbol_step_packed(x0);
            if constexpr (N == 64) s1 = symbol_step_packed(x1);
          } else {
            s0 = symbol_step_rank<kMode == 1>(x0);
            if constexpr (N == 64) s1 = symbol_step_rank<kMode == 1>(x1);
          }
          sts_u8(sTile + t * N + lanePos, s0);
          renorm(x0, ring.wp, ltMask);
          if constexpr (N == 64) {
            sts_u8(sTile + t * N + 32u + lanePos, s1);
            renorm(x1, ring.wp, ltMask);
          }
        }
        __syncwarp();
        st_global_v4(outLane - lanePos + lane * 16u, lds_v4(sTile + lane * 16u));
        __syncwarp();
        outLane += 512;
      }
    }
#endif
#if HSR_QUAD_STORE
    // Quad-transposed stores: four half-rows of symbols (2 rows at N = 64, 4 rows at N = 32) are packed into one
    // register per lane and transposed inside each group of four lanes (two SHFL.BFLY + two PRMT), after which lane
    // 4m + h holds four CONSECUTIVE output bytes of half-row h — lanes 4m .. 4m+3 own byte positions p .. p+3 of a
    // half-row (idx2idx keeps the low two lane bits) — and one 32-bit store per lane replaces four byte stores:
    // 2 x 1.1 + 1.5 cycles of the shared-memory/LSU data pipe instead of 4 x 1.4 (profiles/r1/ubench_lsu.jsonl).
    if ((reinterpret_cast<uintptr_t>(outLane - idx2idx_lane(lane)) & 3) == 0) {
      constexpr uint32_t kGroupRows = N == 64 ? 2 : 4;
      const uint32_t selA = (lane & 1u) ? 0x3715u : 0x6240u, selB = (lane & 2u) ? 0x3276u : 0x5410u;
      uint8_t *outQuad = outLane - idx2idx_lane(lane) + 32u * (lane & 3u) + idx2idx_lane(lane & ~3u);
#pragma unroll(kRowUnroll / kGroupRows > 0 ? kRowUnroll / kGroupRows : 1)
      for (; r + kGroupRows <= rows; r += kGroupRows) {
        uint32_t s[4];
#pragma unroll
        for (uint32_t t = 0; t < kGroupRows; t++) {
          ring.advance_if_needed(lane);
          if constexpr (N == 64) {
            if constexpr (kMode == 0) {
              s[2 * t] = symbol_step_packed(x0);
              s[2 * t + 1] = symbol_step_packed(x1);
            } else {
              s[2 * t] = symbol_step_rank<kMode == 1>(x0);
              s[2 * t + 1] = symbol_step_rank<kMode == 1>(x1);
            }
            renorm(x0, ring.wp, ltMask);
            renorm(x1, ring.wp, ltMask);
          } else {
            if constexpr (kMode == 0)
              s[t] = symbol_step_packed(x0);
            else
              s[t] = symbol_step_rank<kMode == 1>(x0);
            renorm(x0, ring.wp, ltMask);
          }
        }
        uint32_t a;
        if constexpr (kMode == 0) { // the packed entry carries freq above the symbol byte: pick bytes
          a = __byte_perm(__byte_perm(s[0], s[1], 0x0040), __byte_perm(s[2], s[3], 0x0040), 0x5410);
        } else {                    // ranks / symbols are < 256: multiply-adds on the FMA pipe
          a = (s[3] * 256u + s[2]) * 65536u + (s[1] * 256u + s[0]);
        }
        const uint32_t b = __byte_perm(a, __shfl_xor_sync(kFull, a, 1), selA);
        const uint32_t c = __byte_perm(b, __shfl_xor_sync(kFull, b, 2), selB);
        st_global_u32(outQuad, c);
        outQuad += 128;
      }
      outLane += (size_t)r * N;
    }
#endif
#pragma unroll kRowUnroll
    for (; r < rows; r++) {
      ring.advance_if_needed(lane);
      uint32_t s0, s1 = 0;
      s0 = symbol_step<kMode>(x0);
      if constexpr (N == 64)
        s1 = symbol_step<kMode>(x1);
      st_global_u8(outLane, s0);
      renorm(x0, ring.wp, ltMask);
      if constexpr (N == 64) {
        st_global_u8(outLane + 32, s1);
        renorm(x1, ring.wp, ltMask);
      }
      outLane += N;
    }
  }

  __device__ __forceinline__ void rows(const TableInfo &info, uint32_t &x0, uint32_t &x1, Ring<L> &ring, uint8_t *outLane,
                                       uint64_t nrows, uint32_t lane, uint32_t ltMask) const
  {
    while (nrows) { // 64-bit row counts are split so the hot loop keeps a 32-bit counter
      const uint32_t chunk = nrows > 0x40000000ull ? 0x40000000u : (uint32_t)nrows;
      if (TK == TK_WIDE)
        rows_impl<3>(x0, x1, ring, outLane, chunk, lane, ltMask);
      else if (TK == TK_PACKED && !info.degenerate)
        rows_impl<0>(x0, x1, ring, outLane, chunk, lane, ltMask);
      else if (info.allPresent)
        rows_impl<1>(x0, x1, ring, outLane, chunk, lane, ltMask);
      else
        rows_impl<2>(x0, x1, ring, outLane, chunk, lane, ltMask);
      outLane += (uint64_t)chunk * N;
      nrows -= chunk;
    }
  }

  // the < N leftover symbols (src/rANS32x32_16w.cpp:238-266): lanes whose byte position is inside the buffer
  __device__ __forceinline__ void tail(uint32_t &x0, uint32_t &x1, Ring<L> &ring, uint8_t *outLane, uint32_t lanePos,
                                       uint32_t left, uint32_t lane, uint32_t ltMask) const
  {
    ring.advance_if_needed(lane);
    const bool a0 = lanePos < left;
    uint32_t t0 = x0;
    const uint32_t s0 = symbol_step_rank<false>(t0); // the rank table is always built
    if (a0) {
      x0 = t0;
      st_global_u8(outLane, s0);
    }
    renorm_masked(x0, ring.wp, a0, ltMask);
    if constexpr (N == 64) {
      const bool a1 = lanePos + 32u < left;
      uint32_t t1 = x1;
      const uint32_t s1 = symbol_step_rank<false>(t1);
      if (a1) {
        x1 = t1;
        st_global_u8(outLane + 32, s1);
      }
      renorm_masked(x1, ring.wp, a1, ltMask);
    }
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
