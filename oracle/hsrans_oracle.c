/*
 * hsrans_oracle.c — plain C restatement of the hypersonic-rANS decode path (see hsrans_oracle.h).
 * TEST INFRASTRUCTURE ONLY; never linked into or called from the product.
 *
 * Build with FP contraction off (oracle/Makefile passes -ffp-contract=off): hsro_normalize_hist must do a
 * separate float multiply and add like the reference build (src/hist.cpp:60-64).
 *
 * Differences from the reference, all on malformed input only (the reference has undefined behaviour there):
 *   - every read of the compressed stream is bounds-checked against inLength; running off the end returns 0;
 *   - histogram sums are checked in 32 bits (the scalar reference sums in uint16_t, src/hist.cpp:332);
 *   - single-symbol (memset) blocks that would write past the decoded length return 0;
 *   - decoded lengths smaller than the state count return 0 (size_t wrap in src/rANS32x32_16w.cpp:206).
 */
#include "hsrans_oracle.h"

#include <string.h>

#define CONSUME_POINT16 (1u << 15) /* src/rans.h:8 */

static uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* ------------------------------------------------------------------ histogram (src/hist.cpp) */

void hsro_observe_hist(uint32_t hist[256], const uint8_t *data, size_t size)
{
  memset(hist, 0, sizeof(uint32_t) * 256);
  for (size_t i = 0; i < size; i++)
    hist[data[i]]++;
}

/* src/hist.cpp:112-129 */
static void heapify(uint8_t *idx, const uint16_t *val, int64_t n, int64_t i)
{
  for (;;) {
    const int64_t left = 2 * i + 1, right = 2 * i + 2;
    int64_t largest = i;
    if (left < n && val[idx[left]] > val[idx[largest]])
      largest = left;
    if (right < n && val[idx[right]] > val[idx[largest]])
      largest = right;
    if (largest == i)
      return;
    uint8_t t = idx[i]; idx[i] = idx[largest]; idx[largest] = t;
    i = largest; /* tail call in the reference */
  }
}

void hsro_normalize_hist(hsro_hist_t *h, const uint32_t hist[256], size_t dataBytes, uint32_t bits)
{
  const uint32_t total = 1u << bits;
  uint16_t capped[256];
  size_t cappedSum = 0;

  /* src/hist.cpp:60-70: one float divide, then per symbol a float MUL followed by a float ADD */
  const float mul = (float)total / (float)dataBytes;
  for (size_t i = 0; i < 256; i++) {
    volatile float prod = (float)hist[i] * mul; /* volatile: forbid fusing with the add */
    capped[i] = (uint16_t)(prod + 0.5f);
    if (capped[i] == 0 && hist[i])
      capped[i] = 1;
    cappedSum += capped[i];
  }

  if (cappedSum != total) { /* src/hist.cpp:103 */
    uint8_t sorted[256];
    for (size_t i = 0; i < 256; i++)
      sorted[i] = (uint8_t)i;
    /* src/hist.cpp:131-141 heap sort, ascending by capped[] */
    for (int64_t i = 256 / 2 - 1; i >= 0; i--)
      heapify(sorted, capped, 256, i);
    for (int64_t i = 255; i >= 0; i--) {
      uint8_t t = sorted[0]; sorted[0] = sorted[i]; sorted[i] = t;
      heapify(sorted, capped, i, 0);
    }

    size_t minTwo = 0; /* src/hist.cpp:145-154 */
    for (size_t i = 0; i < 256; i++)
      if (capped[sorted[i]] >= 2) { minTwo = i; break; }

    while (cappedSum > total) { /* src/hist.cpp:156-176 "stealing" */
      for (size_t i = minTwo; i < 256; i++) {
        capped[sorted[i]]--;
        cappedSum--;
        if (cappedSum == total)
          goto ready;
      }
      for (size_t i = minTwo; i < 256; i++)
        if (capped[sorted[i]] >= 2) { minTwo = i; break; }
    }
    while (cappedSum < total) { /* src/hist.cpp:178-198 "charity" */
      for (int64_t i = 255; i >= (int64_t)minTwo; i--) {
        capped[sorted[i]]++;
        cappedSum++;
        if (cappedSum == total)
          goto ready;
      }
      for (size_t i = minTwo; i < 256; i++)
        if (capped[sorted[i]] >= 2) { minTwo = i; break; }
    }
  }
ready:;
  size_t counter = 0; /* src/hist.cpp:201-209 */
  for (size_t i = 0; i < 256; i++) {
    h->cumul[i] = (uint16_t)counter;
    h->symbolCount[i] = capped[i];
    counter += capped[i];
  }
}

void hsro_make_hist(hsro_hist_t *h, const uint8_t *data, size_t size, uint32_t bits)
{
  uint32_t hist[256];
  hsro_observe_hist(hist, data, size);
  hsro_normalize_hist(h, hist, size, bits);
}

int hsro_complete_hist(hsro_hist_t *h, uint32_t bits)
{
  uint32_t counter = 0;
  for (size_t i = 0; i < 256; i++) {
    h->cumul[i] = (uint16_t)counter;
    counter += h->symbolCount[i];
  }
  return counter == (1u << bits);
}

int hsro_make_cumul_inv(hsro_hist_t *h, uint32_t bits, uint8_t *cumulInv)
{
  if (!hsro_complete_hist(h, bits))
    return 0;
  const uint32_t total = 1u << bits;
  uint8_t sym = 0; /* src/hist.cpp:343-351 */
  for (uint32_t i = 0; i < total; i++) {
    while (sym != 0xFF && (!h->symbolCount[sym] || h->cumul[sym + 1] <= i))
      sym++;
    cumulInv[i] = sym;
  }
  return 1;
}

/* ------------------------------------------------------------------ shared pieces */

uint32_t hsro_idx2idx(uint32_t j)
{
  /* closed form of the tables at src/block_codec32.h:22 and src/block_codec64.h:22-28 */
  const uint32_t l = j & 31u;
  return (j & ~31u) | (l & 3u) | ((l & 4u) << 2) | ((l & 24u) >> 1);
}

uint32_t hsro_idx2idx16(uint32_t j)
{
  /* src/rANS32x16_16w.cpp:54 (encoder) and :211 (decoder): { 0-3, 8-11, 4-7, 12-15 } */
  return (j & 3u) | ((j & 4u) << 1) | ((j & 8u) >> 1);
}

static inline uint32_t lane_pos(uint32_t N, uint32_t j) { return N == 16 ? hsro_idx2idx16(j) : hsro_idx2idx(j); }

size_t hsro_capacity(uint32_t family, uint32_t N, size_t inputSize)
{
  if (family == HSRO_RAW) /* also src/rANS32x16_16w.cpp:10-13 */
    return inputSize + N + sizeof(uint16_t) * 256 + sizeof(uint32_t) * N + sizeof(uint64_t) * 2;
  if (family == HSRO_RAW32BLK) /* src/rans32x32_32blk_16w.cpp:10-13 */
    return inputSize + N + sizeof(uint16_t) * 256 + sizeof(uint32_t) * N * 2 + sizeof(uint64_t) * 2;
  const size_t minMinBlockSize = (size_t)1 << 15; /* src/block_rANS32x32_16w_encode.cpp:12-13 */
  const size_t baseSize = 2 * sizeof(uint64_t) + 256 * sizeof(uint16_t) + inputSize + N * sizeof(uint32_t);
  const size_t blockCount = (inputSize + minMinBlockSize) / minMinBlockSize + 1;
  const size_t perBlock = family == HSRO_MT ? sizeof(uint64_t) * 2 + 256 * sizeof(uint16_t) + N * sizeof(uint32_t)
                                            : sizeof(uint64_t) + 256 * sizeof(uint16_t);
  return baseSize + blockCount * perBlock;
}

typedef struct dec_state {
  uint32_t N, bits;
  uint32_t states[64];
  hsro_hist_t hist;
  uint8_t cumulInv[1 << 15];
  const uint8_t *rd;    /* read head (2-byte aligned relative to the stream start only) */
  const uint8_t *rdEnd; /* one past the last readable byte */
  int overrun;
} dec_state_t;

/* one symbol of lane j: src/rANS32x32_16w.cpp:17-30 + the renormalisation at :228-232 */
static inline uint8_t step(dec_state_t *d, uint32_t j)
{
  const uint32_t mask = (1u << d->bits) - 1;
  uint32_t x = d->states[j];
  const uint32_t slot = x & mask;
  const uint8_t s = d->cumulInv[slot];
  x = (x >> d->bits) * (uint32_t)d->hist.symbolCount[s] + slot - (uint32_t)d->hist.cumul[s];
  if (x < CONSUME_POINT16) {
    if (d->rd + 2 > d->rdEnd) {
      d->overrun = 1;
    } else {
      x = (x << 16) | rd16(d->rd);
      d->rd += 2;
    }
  }
  d->states[j] = x;
  return s;
}

/* src/block_codec32.h:162-206 / src/block_codec64.h:173-217 (decode_section, scalar) */
static size_t rows(dec_state_t *d, uint8_t *out, size_t i, size_t end)
{
  for (; i < end; i += d->N) {
    for (uint32_t j = 0; j < d->N; j++)
      out[i + lane_pos(d->N, j)] = step(d, j);
    if (d->overrun)
      return i;
  }
  return i;
}

/* the < N leftover symbols: src/rANS32x32_16w.cpp:238-266 */
static void tail(dec_state_t *d, uint8_t *out, size_t i, size_t n)
{
  for (uint32_t j = 0; j < d->N; j++) {
    const uint32_t index = lane_pos(d->N, j);
    if (i + index < n)
      out[i + index] = step(d, j);
  }
}

typedef struct header {
  uint64_t n, compLen;
} header_t;

static int read_header(uint32_t N, uint32_t bits, const uint8_t *in, size_t inLength, size_t outCapacity,
                       header_t *h)
{
  if (!(N == 16 || N == 32 || N == 64) || bits < 10 || bits > 15)
    return 0;
  if (inLength < 16 + 4 * (size_t)N + 512) /* src/rANS32x32_16w.cpp:164, src/rANS32x16_16w.cpp:165 */
    return 0;
  h->n = rd64(in);
  if (h->n > outCapacity) /* :173 */
    return 0;
  h->compLen = rd64(in + 8);
  if (inLength < h->compLen) /* :179 */
    return 0;
  if (h->n < N) /* reference: size_t wrap at :206, undefined; refuse */
    return 0;
  return 1;
}

/* ------------------------------------------------------------------ raw */

size_t hsro_decode_raw(uint32_t N, uint32_t bits, const uint8_t *in, size_t inLength, uint8_t *out,
                       size_t outCapacity)
{
  header_t h;
  if (!read_header(N, bits, in, inLength, outCapacity, &h))
    return 0;

  static dec_state_t sd; /* 33 KB; not re-entrant, fine for a test oracle */
  dec_state_t *d = &sd;
  memset(d, 0, sizeof(*d));
  d->N = N; d->bits = bits;
  for (size_t i = 0; i < 256; i++) /* :183-187 */
    d->hist.symbolCount[i] = rd16(in + 16 + 2 * i);
  if (!hsro_make_cumul_inv(&d->hist, bits, d->cumulInv)) /* :189 */
    return 0;
  for (uint32_t j = 0; j < N; j++) /* :194-198 */
    d->states[j] = rd32(in + 16 + 512 + 4 * j);
  d->rd = in + 16 + 512 + 4 * (size_t)N;
  d->rdEnd = in + inLength;

  const size_t outLengthInStates = h.n - N + 1; /* :206 */
  size_t i = rows(d, out, 0, outLengthInStates);
  if (d->overrun)
    return 0;
  tail(d, out, i, h.n);
  if (d->overrun)
    return 0;
  return h.n;
}

/* ------------------------------------------------------------------ rANS32x32_32blk_16w */

/* src/rans32x32_32blk_16w.cpp:183-301. Same header as raw, then u32 blockSize[31] (bytes of the word sub-stream of
 * states 0..30; the last one needs none, :160-167) and the 32 sub-streams back to back: state j reads its words
 * from its OWN head pReadHead[j] (:224-233), so there is no shared cursor. */
static int step_32blk(dec_state_t *d, const uint8_t **rd, const uint8_t *end, uint32_t j, uint8_t *sym)
{
  const uint32_t mask = (1u << d->bits) - 1;
  uint32_t x = d->states[j];
  const uint32_t slot = x & mask;
  const uint8_t s = d->cumulInv[slot]; /* :17-30 */
  *sym = s;
  x = (x >> d->bits) * (uint32_t)d->hist.symbolCount[s] + slot - (uint32_t)d->hist.cumul[s];
  if (x < CONSUME_POINT16) { /* :258-262: this state's own read head */
    if (rd[j] + 2 > end)
      return 0;
    x = (x << 16) | rd16(rd[j]);
    rd[j] += 2;
  }
  d->states[j] = x;
  return 1;
}

size_t hsro_decode_32blk(uint32_t bits, const uint8_t *in, size_t inLength, uint8_t *out, size_t outCapacity)
{
  const uint32_t N = 32;
  if (bits < 10 || bits > 15)
    return 0;
  if (inLength < 16 + 4 * (size_t)(2 * N - 1) + 512) /* :185 */
    return 0;
  const uint64_t n = rd64(in);
  if (n > outCapacity) /* :194 */
    return 0;
  const uint64_t compLen = rd64(in + 8);
  if (inLength < compLen) /* :200 */
    return 0;
  if (n < N) /* size_t wrap at :238, undefined in the reference; refuse */
    return 0;

  static dec_state_t sd;
  dec_state_t *d = &sd;
  memset(d, 0, sizeof(*d));
  d->N = N; d->bits = bits;
  for (size_t i = 0; i < 256; i++) /* :205-209 */
    d->hist.symbolCount[i] = rd16(in + 16 + 2 * i);
  if (!hsro_make_cumul_inv(&d->hist, bits, d->cumulInv)) /* :211 */
    return 0;
  for (uint32_t j = 0; j < N; j++) /* :216-220 */
    d->states[j] = rd32(in + 16 + 512 + 4 * j);

  const uint8_t *sizes = in + 16 + 512 + 4 * (size_t)N;
  const uint8_t *end = in + inLength;
  const uint8_t *rd[32];
  rd[0] = sizes + 4 * (size_t)(N - 1); /* :223 */
  for (uint32_t j = 1; j < N; j++) {   /* :225-231 */
    const uint32_t blockSize = rd32(sizes + 4 * (size_t)(j - 1));
    if ((size_t)(end - rd[j - 1]) < blockSize)
      return 0; /* the reference would read out of bounds here */
    rd[j] = rd[j - 1] + blockSize;
  }

  const size_t outLengthInStates = n - N + 1; /* :238 */
  size_t i = 0;
  for (; i < outLengthInStates; i += N) /* full rows, :241-269 */
    for (uint32_t j = 0; j < N; j++)
      if (!step_32blk(d, rd, end, j, &out[i + hsro_idx2idx(j)]))
        return 0;
  for (uint32_t j = 0; j < N; j++) { /* the < 32 leftover symbols, :271-298 */
    const uint32_t index = hsro_idx2idx(j);
    if (i + index < n && !step_32blk(d, rd, end, j, &out[i + index]))
      return 0;
  }
  return n;
}

/* ------------------------------------------------------------------ block_ and mt_ */

static size_t decode_blocked(int mt, uint32_t N, uint32_t bits, const uint8_t *in, size_t inLength, uint8_t *out,
                             size_t outCapacity)
{
  header_t h;
  if (!read_header(N, bits, in, inLength, outCapacity, &h))
    return 0;

  static dec_state_t sd;
  dec_state_t *d = &sd;
  memset(d, 0, sizeof(*d));
  d->N = N; d->bits = bits;
  d->rdEnd = in + inLength;
  d->rd = in + 16;
  int haveHist = 0;

  if (!mt) { /* block_: states once (src/block_rANS32x32_16w_decode.cpp:42-46) */
    for (uint32_t j = 0; j < N; j++)
      d->states[j] = rd32(d->rd + 4 * j);
    d->rd += 4 * (size_t)N;
  }

  const size_t outLengthInStates = h.n - N + 1;
  size_t i = 0;

  do {
    if (d->rd + 8 > d->rdEnd)
      return 0;
    const uint64_t blockSizeVal = rd64(d->rd); /* block_ :55, mt_ :43 */
    d->rd += 8;

    if (blockSizeVal & ((uint64_t)1 << 63)) { /* single symbol run: block_ :58-66, mt_ :46-54 */
      const uint8_t symbol = (uint8_t)((blockSizeVal >> 54) & 0xFF);
      const uint64_t blockSize = blockSizeVal & (((uint64_t)1 << 54) - 1);
      if (blockSize > h.n - i)
        return 0;
      memset(out + i, symbol, blockSize);
      i += blockSize;
    } else {
      const uint8_t *after = NULL;
      if (mt) { /* mt_ :57-66 */
        if (d->rd + 8 + 4 * (size_t)N > d->rdEnd)
          return 0;
        const uint64_t skip = rd64(d->rd);
        d->rd += 8;
        if (skip >= (uint64_t)(d->rdEnd - d->rd) / 2)
          return 0;
        after = d->rd + 2 * (skip + 1);
        for (uint32_t j = 0; j < N; j++)
          d->states[j] = rd32(d->rd + 4 * j);
        d->rd += 4 * (size_t)N;
      }
      if (d->rd + 512 > d->rdEnd)
        return 0;
      for (size_t j = 0; j < 256; j++) /* block_ :69-73, mt_ :68-72 */
        d->hist.symbolCount[j] = rd16(d->rd + 2 * j);
      d->rd += 512;
      if (!hsro_make_cumul_inv(&d->hist, bits, d->cumulInv)) /* _init_from_hist */
        return 0;
      haveHist = 1;

      uint64_t blockEnd = i + blockSizeVal; /* block_ :78-83, mt_ :77-82 */
      if (blockEnd > outLengthInStates)
        blockEnd = outLengthInStates;
      else if ((blockEnd & (N - 1)) != 0)
        return 0;

      i = rows(d, out, i, blockEnd);
      if (d->overrun)
        return 0;
      if (mt) {
        if (i > outLengthInStates) { /* mt_ :86-92 */
          if (i >= h.n)
            return h.n;
          break;
        }
        d->rd = after; /* mt_ :94 */
      }
    }
    if (!mt && i > outLengthInStates) { /* block_ :88-94 */
      if (i >= h.n)
        return h.n;
      break;
    }
  } while (i < outLengthInStates);

  if (i < h.n) { /* block_ :98-139, mt_ :99-130 */
    if (!haveHist)
      return 0; /* reference: zero histogram fails inplace_make_hist_dec */
    tail(d, out, i, h.n);
    if (d->overrun)
      return 0;
  }
  return h.n;
}

size_t hsro_decode_block(uint32_t N, uint32_t bits, const uint8_t *in, size_t inLength, uint8_t *out,
                         size_t outCapacity)
{
  if (N == 16)
    return 0; /* no 16-state block_ codec exists */
  return decode_blocked(0, N, bits, in, inLength, out, outCapacity);
}

size_t hsro_decode_mt(uint32_t N, uint32_t bits, const uint8_t *in, size_t inLength, uint8_t *out,
                      size_t outCapacity)
{
  if (N == 16)
    return 0;
  return decode_blocked(1, N, bits, in, inLength, out, outCapacity);
}

size_t hsro_decode(uint32_t family, uint32_t N, uint32_t bits, const uint8_t *in, size_t inLength, uint8_t *out,
                   size_t outCapacity)
{
  switch (family) {
  case HSRO_RAW: return hsro_decode_raw(N, bits, in, inLength, out, outCapacity);
  case HSRO_BLOCK: return hsro_decode_block(N, bits, in, inLength, out, outCapacity);
  case HSRO_MT: return hsro_decode_mt(N, bits, in, inLength, out, outCapacity);
  case HSRO_RAW32BLK: return N == 32 ? hsro_decode_32blk(bits, in, inLength, out, outCapacity) : 0;
  default: return 0;
  }
}

size_t hsro_mt_walk(uint32_t N, const uint8_t *in, size_t inLength, hsro_mt_block_t *blocks, size_t maxBlocks)
{
  if (inLength < 16 + 4 * (size_t)N + 512)
    return (size_t)-1;
  const uint64_t n = rd64(in);
  if (n < N)
    return (size_t)-1;
  const uint64_t outLengthInStates = n - N + 1;
  size_t pos = 16, count = 0;
  uint64_t i = 0;
  do {
    if (pos + 8 > inLength)
      return (size_t)-1;
    const uint64_t v = rd64(in + pos);
    hsro_mt_block_t b;
    b.inOffset = pos;
    b.outOffset = i;
    if (v & ((uint64_t)1 << 63)) {
      b.size = v & (((uint64_t)1 << 54) - 1);
      b.kind = 1 | (((v >> 54) & 0xFF) << 8);
      if (b.size > n - i)
        return (size_t)-1;
      pos += 8;
      i += b.size;
    } else {
      if (pos + 16 > inLength)
        return (size_t)-1;
      const uint64_t skip = rd64(in + pos + 8);
      if (skip >= (inLength - (pos + 16)) / 2)
        return (size_t)-1;
      uint64_t end = i + v;
      if (end > outLengthInStates)
        end = outLengthInStates;
      else if (end & (N - 1))
        return (size_t)-1;
      /* decode_section advances in whole rows until i >= end */
      const uint64_t rowsN = end > i ? (end - i + N - 1) / N : 0;
      b.size = rowsN * N;
      b.kind = 0;
      pos = pos + 16 + 2 * (skip + 1);
      i += b.size;
    }
    if (count < maxBlocks)
      blocks[count] = b;
    count++;
  } while (i < outLengthInStates);
  return count;
}

/* ------------------------------------------------------------------ raw encoder (stream producer twin) */

size_t hsro_encode_raw(uint32_t N, uint32_t bits, const uint8_t *in, size_t length, uint8_t *out,
                       size_t outCapacity, const hsro_hist_t *hist)
{
  if (!(N == 16 || N == 32 || N == 64) || bits < 10 || bits > 15 || length == 0)
    return 0;
  if (outCapacity < hsro_capacity(HSRO_RAW, N, length)) /* src/rANS32x32_16w.cpp:37, src/rANS32x16_16w.cpp:37 */
    return 0;
  const uint32_t emitPoint = (CONSUME_POINT16 >> bits) << 16; /* :41 */
  uint32_t states[64];
  for (uint32_t j = 0; j < N; j++)
    states[j] = CONSUME_POINT16;

  /* words are written downward from the end of the buffer (:44-45), then moved behind the header */
  uint8_t *pEnd = out + outCapacity - 2;
  uint8_t *pStart = pEnd;

  int64_t i = (int64_t)length - 1; /* :55-57 */
  i &= ~(int64_t)(N - 1);
  i += N;

  for (; i >= (int64_t)N; i -= N) { /* first pass is the ragged row (:59-95), then full rows (:99-128) */
    for (int64_t j = (int64_t)N - 1; j >= 0; j--) {
      const int64_t pos = i - (int64_t)N + (int64_t)lane_pos(N, (uint32_t)j);
      if (pos >= (int64_t)length)
        continue;
      const uint8_t sym = in[pos];
      const uint32_t freq = hist->symbolCount[sym];
      const uint32_t max = emitPoint * freq;
      uint32_t x = states[j];
      if (x >= max) {
        const uint16_t w = (uint16_t)(x & 0xFFFF);
        memcpy(pStart, &w, 2);
        pStart -= 2;
        x >>= 16;
      }
      states[j] = ((x / freq) << bits) + (uint32_t)hist->cumul[sym] + (x % freq);
    }
  }

  size_t o = 0; /* :130-158 */
  const uint64_t len64 = length;
  memcpy(out + o, &len64, 8); o += 16;
  for (size_t j = 0; j < 256; j++) { memcpy(out + o, &hist->symbolCount[j], 2); o += 2; }
  for (uint32_t j = 0; j < N; j++) { memcpy(out + o, &states[j], 4); o += 4; }
  const size_t size = (size_t)(pEnd - pStart);
  memmove(out + o, pStart + 2, size);
  o += size;
  const uint64_t total = o;
  memcpy(out + 8, &total, 8);
  return o;
}
