// hsr_kernels.cuh — kernel bodies shared by the per-(bits, N, table) instantiation units.
#pragma once

#include "hsr_device.cuh"

namespace hsr {

// Everything a decode launch needs. Offsets inside hsr_block_t are absolute stream / output offsets; `in` and
// `out` point at stream offset inBase and decoded offset outBase (shards hold only their slice).
struct DecodeParams {
  const uint8_t *in;
  uint64_t inBase;
  uint8_t *out;
  uint64_t outBase;
  const hsr_block_t *blocks;
  uint32_t numBlocks;
  uint32_t *work;         // this launch's slot of the device's ring: [0] next unit to claim, [1] CTAs that have left
  uint32_t *status;       // status bits (OR over all units)
  uint32_t *streamStatus; // optional: status bits per stream, indexed by hsr_block_t::reserved (batch decode)
};

// block_ framing: every stream is one sequential recurrence; a batch gives one warp to each stream
struct BlockStreamDesc {
  uint64_t inOffset, inLength; // stream bytes at in + inOffset
  uint64_t outOffset, n;       // decoded bytes at out + outOffset
};

struct BlockStreamParams {
  const uint8_t *in;
  uint8_t *out;
  const BlockStreamDesc *streams; // device array, or nullptr: `single` is the only stream
  BlockStreamDesc single;
  uint32_t numStreams;
  uint32_t *counter;      // [0] work counter, [1] status bits
  uint32_t *streamStatus; // optional, per stream
};

__device__ __forceinline__ void raise(uint32_t *status, uint32_t *streamStatus, uint32_t stream, uint32_t bits, uint32_t lane)
{
  if (lane == 0) {
    atomicOr(status, bits);
    if (streamStatus)
      atomicOr(streamStatus + stream, bits);
  }
}

// Units kernels overlap back to back (programmatic dependent launch): every CTA releases the NEXT launch of the stream
// as its first instruction, so that launch's CTAs take over SM slots one by one as this grid's persistent warps run out
// of units — the drain of one decode (a last, partial round of the grid: ~6 % of a 15 k-block step, most of a step that
// has fewer blocks than the GPU has warp slots) is filled with the start of the next. Units kernels never read what
// another units kernel wrote, so nothing waits at the start; a CTA waits for the previous grid only right before it
// exits, which keeps completion in stream order. Launches whose predecessor in the stream is anything else (a copy, a
// memset, a foreign kernel) serialise as usual.
__device__ __forceinline__ void units_enter() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// The work counter is a launch-private slot of a ring in device memory ({next unit, CTAs that have left}); the last CTA
// to leave hands the slot back zeroed, so no memset sits between two launches and launches in flight never share one.
__device__ __forceinline__ void units_leave(uint32_t *work, uint32_t lane)
{
  if (lane == 0) {
    const uint32_t left = atomicAdd(work + 1, 1u);
    if (left + 1u == gridDim.x) {
      atomicExch(work, 0u);
      atomicExch(work + 1, 0u);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------- mt_ / raw

// Persistent one-warp CTAs pull units (mt_ blocks, fills, or one raw stream) from a global counter. The CTA's
// static shared memory holds the tables of its current block + its word ring at compile-time addresses.
template <int BITS, int N, int TK>
__device__ __forceinline__ void units_kernel_body(const DecodeParams &p)
{
  using L = WarpLayout<BITS, N, TK>;
  uint32_t sw;
  if constexpr (L::kDynamic)
    sw = dynamic_smem_base();
  else
    sw = declare_smem<L::kBytes>();
  const uint32_t lane = lane_id();
  const uint32_t ltMask = lanemask_lt();
  const uint32_t lanePos = idx2idx_lane(lane);

  Decoder<BITS, N, TK> dec;
  dec.init(sw);
  Ring<L> ring;
#if HSR_RING_TMA
  ring.init(sw + L::kOffRing, sw + L::kOffBar, lane);
#endif

  // While plenty of units remain, the next one is claimed as soon as the current one is known, so the atomic's round
  // trip hides behind the decode. Near the end (less than two rounds of the grid left) units are claimed only when a
  // warp is free: a warp that sat on a pre-claimed unit would leave others idle — with as many units as CTAs (a
  // batch of raw streams) the early CTAs would take two units each and the late ones none.
  units_enter();
  uint32_t claimed = 0; // meaningful on lane 0 only, 0 elsewhere
  bool ahead = true;
  if (lane == 0)
    claimed = atomicAdd(p.work, 1u);
  for (;;) {
    if (!ahead && lane == 0)
      claimed = atomicAdd(p.work, 1u);
    // lanes other than 0 hold 0, so the warp sum IS lane 0's claim — and, unlike a shuffle, REDUX delivers it in a
    // uniform register: the loop exit below is then provably warp-uniform. With a shuffle ptxas has to assume that
    // some lanes leave the loop early and stay alive in units_leave(), and wraps every vote / shuffle of the hot loop
    // in WARPSYNC + a convergence barrier (+7 % instructions, 1.100 -> 1.238 ms on the 1 GB step).
    const uint32_t b = __reduce_add_sync(kFull, claimed);
    if (b >= p.numBlocks)
      break;
    ahead = (uint64_t)b + 2ull * gridDim.x < p.numBlocks;
    if (ahead && lane == 0)
      claimed = atomicAdd(p.work, 1u);

    const hsr_block_t *blk = p.blocks + b;
    const uint64_t inOffset = __ldg(&blk->inOffset);
    const uint64_t inEnd = __ldg(&blk->inEnd);
    const uint64_t outOffset = __ldg(&blk->outOffset);
    const uint64_t count = __ldg(&blk->count);
    const uint32_t kind = __ldg(&blk->kind);
    const uint32_t tailCount = __ldg(&blk->tail);
    const uint32_t streamId = __ldg(&blk->reserved);
    uint8_t *out = p.out + (outOffset - p.outBase);

    if (kind == 1u) {
      warp_fill(out, __ldg(&blk->symbol), count, lane);
      continue;
    }

    const uint8_t *base = p.in + (inOffset - p.inBase);
    const uint8_t *end = p.in + (inEnd - p.inBase);
    const uint8_t *statesPtr = kind == 0u ? base : base + 512;     // mt_: states then counts; raw: counts then states
    const uint8_t *countsPtr = kind == 0u ? base + 4 * N : base;
    const uint8_t *words = base + 4 * N + 512;

    // states and the first word segments are requested before the table build, whose own loads they overlap
    uint32_t x0 = ldg_u32_a2(statesPtr + 4 * lane);
    uint32_t x1 = 0;
    if constexpr (N == 64)
      x1 = ldg_u32_a2(statesPtr + 4 * (lane + 32));
#if HSR_RING_TMA
    ring.start(words, end, lane);
#else
    ring.start(sw + L::kOffRing, words, end, lane);
#endif

    const TableInfo info = build_tables<BITS, N, TK>(sw, countsPtr, lane);
    if (!info.ok) {
      ring.start_wait();
      ring.drain();
      raise(p.status, p.streamStatus, streamId, HSR_ERR_HIST, lane);
      continue;
    }
    ring.start_wait();

    const uint64_t rows = (count - tailCount) / N;
    uint8_t *outLane = out + lanePos;
    dec.rows(info, x0, x1, ring, outLane, rows, lane, ltMask);
    if (tailCount)
      dec.tail(x0, x1, ring, outLane + rows * N, lanePos, tailCount, lane, ltMask);
    ring.drain();
    if (ring.cursor() > ring.glimit)
      raise(p.status, p.streamStatus, streamId, HSR_ERR_OVERRUN, lane);
#if HSR_RING_TMA
    if (ring.stuck)
      raise(p.status, p.streamStatus, streamId, HSR_ERR_INTERNAL, lane);
#endif
  }
  units_leave(p.work, lane);
}

// ---------------------------------------------------------------------------------------------- block_

// block_ framing (src/block_rANS32x32_16w_decode.cpp:18-142): the N states are read once and carried through
// every block; each block header sits in-band at the word cursor, so block k+1 cannot be located before block k
// has been decoded. One warp walks the whole stream, rebuilding its tables between sections.
template <int BITS, int N, int TK>
__device__ __forceinline__ void block_stream_decode(const BlockStreamParams &p, const BlockStreamDesc &d, uint32_t streamId, uint32_t sw,
                                                    uint32_t lane, uint32_t ltMask, uint32_t lanePos, Ring<WarpLayout<BITS, N, TK>> &ring)
{
  using L = WarpLayout<BITS, N, TK>;

  Decoder<BITS, N, TK> dec;
  dec.init(sw);
  TableInfo info{false, false, false};

  const uint8_t *in = p.in + d.inOffset;
  const uint8_t *streamEnd = in + d.inLength;
  uint8_t *outBase = p.out + d.outOffset;
  uint64_t pos = 16;
  uint32_t x0 = ldg_u32_a2(in + pos + 4 * lane);
  uint32_t x1 = 0;
  if constexpr (N == 64)
    x1 = ldg_u32_a2(in + pos + 4 * (lane + 32));
  pos += 4 * N;

  const uint64_t n = d.n;
  const uint64_t outLengthInStates = n - N + 1;
  uint64_t i = 0;
  bool haveHist = false;
  const uint32_t sRing = sw + L::kOffRing;

  do {
    if (pos + 8 > d.inLength) {
      raise(p.counter + 1, p.streamStatus, streamId, HSR_ERR_OVERRUN, lane);
      return;
    }
    const uint64_t v = ldg_u64_a2(in + pos);
    pos += 8;
    if (v >> 63) { // single-symbol run (:58-66)
      const uint32_t symbol = (uint32_t)(v >> 54) & 0xffu;
      const uint64_t size = v & ((1ull << 54) - 1);
      if (size > n - i) {
        raise(p.counter + 1, p.streamStatus, streamId, HSR_ERR_BOUNDS, lane);
        return;
      }
      warp_fill(outBase + i, symbol, size, lane);
      i += size;
    } else {
      if (pos + 512 > d.inLength) {
        raise(p.counter + 1, p.streamStatus, streamId, HSR_ERR_OVERRUN, lane);
        return;
      }
      info = build_tables<BITS, N, TK>(sw, in + pos, lane); // :69-76
      if (!info.ok) {
        raise(p.counter + 1, p.streamStatus, streamId, HSR_ERR_HIST, lane);
        return;
      }
      haveHist = true;
      pos += 512;

      uint64_t blockEnd = i + v; // :78-83
      if (blockEnd > outLengthInStates)
        blockEnd = outLengthInStates;
      else if (blockEnd & (N - 1)) {
        raise(p.counter + 1, p.streamStatus, streamId, HSR_ERR_ALIGN, lane);
        return;
      }
      const uint64_t rows = blockEnd > i ? (blockEnd - i + N - 1) / N : 0;
#if HSR_RING_TMA
      ring.start(in + pos, streamEnd, lane);
#else
      ring.start(sRing, in + pos, streamEnd, lane);
#endif
      ring.start_wait();
      dec.rows(info, x0, x1, ring, outBase + i + lanePos, rows, lane, ltMask);
      ring.drain();
      if (ring.cursor() > ring.glimit) {
        raise(p.counter + 1, p.streamStatus, streamId, HSR_ERR_OVERRUN, lane);
        return;
      }
      pos = (uint64_t)(ring.gbase - in) + ring.cursor();
      i += rows * N;
    }
    if (i > outLengthInStates) { // :88-94
      if (i >= n)
        return;
      break;
    }
  } while (i < outLengthInStates);

  if (i < n) { // :98-139
    if (!haveHist) {
      raise(p.counter + 1, p.streamStatus, streamId, HSR_ERR_HIST, lane);
      return;
    }
#if HSR_RING_TMA
    ring.start(in + pos, streamEnd, lane);
#else
    ring.start(sRing, in + pos, streamEnd, lane);
#endif
    ring.start_wait();
    dec.tail(x0, x1, ring, outBase + i + lanePos, lanePos, (uint32_t)(n - i), lane, ltMask);
    ring.drain();
  }
}

template <int BITS, int N, int TK>
__device__ __forceinline__ void block_kernel_body(const BlockStreamParams &p)
{
  using L = WarpLayout<BITS, N, TK>;
  uint32_t sw;
  if constexpr (L::kDynamic)
    sw = dynamic_smem_base();
  else
    sw = declare_smem<L::kBytes>();
  const uint32_t lane = lane_id();
  const uint32_t ltMask = lanemask_lt();
  const uint32_t lanePos = idx2idx_lane(lane);
  Ring<L> ring;
#if HSR_RING_TMA
  ring.init(sw + L::kOffRing, sw + L::kOffBar, lane);
#endif
  if (p.streams == nullptr) {
    block_stream_decode<BITS, N, TK>(p, p.single, 0u, sw, lane, ltMask, lanePos, ring);
    return;
  }
  for (;;) { // persistent: streams are handed out by an atomic counter
    uint32_t s = 0;
    if (lane == 0)
      s = atomicAdd(p.counter, 1u);
    s = __shfl_sync(kFull, s, 0);
    if (s >= p.numStreams)
      break;
    BlockStreamDesc d;
    d.inOffset = __ldg(&p.streams[s].inOffset);
    d.inLength = __ldg(&p.streams[s].inLength);
    d.outOffset = __ldg(&p.streams[s].outOffset);
    d.n = __ldg(&p.streams[s].n);
    block_stream_decode<BITS, N, TK>(p, d, s, sw, lane, ltMask, lanePos, ring);
  }
}

// ---------------------------------------------------------------------------------------------- launch table

typedef void (*units_kernel_t)(DecodeParams);
typedef void (*block_kernel_t)(BlockStreamParams);

struct KernelEntry {
  const void *units; // mt_ blocks / fills / one raw stream; one warp per CTA, persistent
  const void *block; // block_ framing, one warp
  int smemBytes;     // shared memory per CTA
  int dynamic;       // 1: smemBytes is DYNAMIC shared memory (tables beyond the 48 KB static limit), passed at launch
};

// defined in hsr_kernels_n32.cu / hsr_kernels_n64.cu; index [bits - 10][table - 1]
extern const KernelEntry kKernels32[6][3];
extern const KernelEntry kKernels64[6][3];
// defined in hsr_kernels_aux.cu; index [bits - 10][0 = bitmap-rank tables, 1 = one-lookup tables] (units kernel only)
extern const KernelEntry kKernelsRaw16[6][2]; // rANS32x16_16w
extern const KernelEntry kKernelsBlk32[6][2]; // rANS32x32_32blk_16w

} // namespace hsr
