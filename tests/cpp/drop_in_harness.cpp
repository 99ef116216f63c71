// drop_in_harness.cpp — the reference harness's validation protocol (src/main.cpp:841-898, re-written, not copied)
// run against the C++ wrappers in hypersonic-rans_b200/cpp/hsrans_b200_codecs.hpp: a table of decodeFunc pointers
// exactly like `_Codecs[].decoders[]`, output poisoned with 0xCC, decode, compare size, memcmp.
//   usage: drop_in_harness <codec-name> <stream-file> <expected-file>      exit 0 = validated
//          drop_in_harness --list
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../hypersonic-rans_b200/cpp/hsrans_b200_codecs.hpp"

typedef size_t (*decodeFunc)(const uint8_t *pInData, const size_t inLength, uint8_t *pOutData, const size_t outCapacity); // src/main.cpp:149

struct entry_t { const char *name; decodeFunc func; };

#define ROWS(prefix) { #prefix "_15", prefix##_15 }, { #prefix "_14", prefix##_14 }, { #prefix "_13", prefix##_13 }, \
                     { #prefix "_12", prefix##_12 }, { #prefix "_11", prefix##_11 }, { #prefix "_10", prefix##_10 }

static const entry_t decoders[] = {
  ROWS(cuda_rANS32x32_16w_decode), ROWS(cuda_rANS32x64_16w_decode), ROWS(cuda_block_rANS32x32_16w_decode),
  ROWS(cuda_block_rANS32x64_16w_decode), ROWS(cuda_mt_rANS32x32_16w_decode), ROWS(cuda_mt_rANS32x64_16w_decode),
  ROWS(cuda_rANS32x16_16w_decode), ROWS(cuda_rANS32x32_32blk_16w_decode),
};

static std::vector<uint8_t> slurp(const char *path)
{
  std::vector<uint8_t> v;
  FILE *f = fopen(path, "rb");
  if (!f) return v;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  v.resize((size_t)n);
  if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) v.clear();
  fclose(f);
  return v;
}

int main(int argc, char **argv)
{
  if (argc == 2 && !strcmp(argv[1], "--list")) {
    for (const entry_t &e : decoders) puts(e.name);
    return 0;
  }
  if (argc != 4) { fprintf(stderr, "usage: %s <codec-name> <stream-file> <expected-file>\n", argv[0]); return 2; }
  const entry_t *codec = nullptr;
  for (const entry_t &e : decoders) if (!strcmp(e.name, argv[1])) codec = &e;
  if (!codec) { fprintf(stderr, "unknown codec %s\n", argv[1]); return 2; }
  const std::vector<uint8_t> stream = slurp(argv[2]), expected = slurp(argv[3]);
  if (stream.empty() || expected.empty()) { fprintf(stderr, "cannot read inputs\n"); return 2; }

  std::vector<uint8_t> out(expected.size() + 64);
  memset(out.data(), 0xCC, out.size());                                           // src/main.cpp:860
  const size_t decodedSize = codec->func(stream.data(), stream.size(), out.data(), expected.size());
  if (decodedSize != expected.size()) {                                           // src/main.cpp:891-897
    fprintf(stderr, "Failed to validate: decoded %zu of %zu bytes (%s)\n", decodedSize, expected.size(), hsr_last_error());
    return 1;
  }
  if (memcmp(out.data(), expected.data(), expected.size()) != 0) { fprintf(stderr, "Failed to validate: bytes differ\n"); return 1; }
  for (size_t i = expected.size(); i < out.size(); i++)
    if (out[i] != 0xCC) { fprintf(stderr, "wrote past the decoded length\n"); return 1; }
  printf("%s validated %zu bytes\n", codec->name, decodedSize);
  return 0;
}
