// hsr_synth.cpp — deterministic synthetic inputs for tests and bench.py (host side, not on the decode path).
// Zipf(s) over 256 ranks, mapped to bytes through a pseudo-random permutation that is either fixed ("iid",
// stationary) or re-drawn every segmentBytes ("pw64k" when 65536), see SURVEY.md §8d. The output depends only on
// (n, s, seed, segmentBytes), never on the thread count.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/hsrans_b200.h"

namespace {

struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next()
  {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
};

struct Alias { // Walker alias table over 256 outcomes, 32-bit thresholds
  uint32_t threshold[256];
  uint8_t alias[256];
};

void build_alias(Alias &a, double s)
{
  double p[256], sum = 0;
  for (int r = 0; r < 256; r++) { p[r] = std::pow((double)(r + 1), -s); sum += p[r]; }
  double scaled[256];
  std::vector<int> small, large;
  for (int r = 0; r < 256; r++) { scaled[r] = p[r] / sum * 256.0; (scaled[r] < 1.0 ? small : large).push_back(r); }
  for (int r = 0; r < 256; r++) { a.threshold[r] = 0xffffffffu; a.alias[r] = (uint8_t)r; }
  while (!small.empty() && !large.empty()) {
    const int l = small.back(); small.pop_back();
    const int g = large.back(); large.pop_back();
    a.threshold[l] = (uint32_t)std::min(4294967295.0, scaled[l] * 4294967296.0);
    a.alias[l] = (uint8_t)g;
    scaled[g] = scaled[g] + scaled[l] - 1.0;
    (scaled[g] < 1.0 ? small : large).push_back(g);
  }
}

void make_perm(uint8_t perm[256], uint64_t seed, uint64_t index)
{
  SplitMix64 rng(seed * 0x2545f4914f6cdd1dull + index * 0x9e3779b97f4a7c15ull + 0x1234567ull);
  for (int i = 0; i < 256; i++) perm[i] = (uint8_t)i;
  for (int i = 255; i > 0; i--) {
    const uint32_t j = (uint32_t)(rng.next() % (uint64_t)(i + 1));
    std::swap(perm[i], perm[j]);
  }
}

constexpr size_t kChunk = 65536;

void fill_chunk(uint8_t *out, size_t begin, size_t end, const Alias &a, uint64_t seed, size_t segmentBytes)
{
  SplitMix64 rng(seed ^ (0xd6e8feb86659fd93ull * (uint64_t)(begin / kChunk + 1)));
  uint8_t perm[256];
  uint64_t permIndex = ~0ull;
  for (size_t i = begin; i < end; i++) {
    const uint64_t want = segmentBytes ? i / segmentBytes : 0;
    if (want != permIndex) { permIndex = want; make_perm(perm, seed, permIndex); }
    const uint64_t r = rng.next();
    const uint32_t col = (uint32_t)(r & 0xffu);
    const uint32_t u = (uint32_t)(r >> 32);
    const uint32_t rank = u <= a.threshold[col] ? col : a.alias[col];
    out[i] = perm[rank];
  }
}

} // namespace

extern "C" int hsr_synth_zipf(uint8_t *out, size_t n, double s, uint64_t seed, size_t segmentBytes)
{
  if (!out || s < 0) return -1;
  Alias a;
  build_alias(a, s);
  const size_t chunks = (n + kChunk - 1) / kChunk;
  unsigned threads = std::thread::hardware_concurrency();
  if (threads == 0) threads = 1;
  threads = (unsigned)std::min<size_t>(threads, std::max<size_t>(1, chunks / 16));
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < threads; t++) {
    pool.emplace_back([&, t]() {
      for (size_t c = t; c < chunks; c += threads)
        fill_chunk(out, c * kChunk, std::min(n, (c + 1) * kChunk), a, seed, segmentBytes);
    });
  }
  for (auto &th : pool) th.join();
  return 0;
}
