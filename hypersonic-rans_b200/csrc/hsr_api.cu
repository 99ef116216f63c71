// hsr_api.cu — host side of the C-ABI in include/hsrans_b200.h: header validation, the mt_ block index,
// sharding, device contexts, the host<->device pipeline and the kernel launches.
//
// Mirrors the framing logic of the reference's decode entry points (error returns included):
//   raw    src/rANS32x32_16w.cpp:161-201 (16 states: src/rANS32x16_16w.cpp:162-203; 32blk: src/rans32x32_32blk_16w.cpp:183-231)
//   block_ src/block_rANS32x32_16w_decode.cpp:18-96
//   mt_    src/mt_rANS32x64_16w_decode.cpp:12-97  (the serial header walk of :40-66,94 becomes hsr_mt_index)
// There is no CPU decode path in this file: without a CUDA device every compute entry point fails.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "hsr_kernels.cuh"

using namespace hsr;

// ------------------------------------------------------------------------------------------------ errors / options

static thread_local std::string g_err;

static void set_err(const char *fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

#define CU_TRY(call, onfail)                                                                              \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) {                                                                              \
      set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);               \
      (void)cudaGetLastError();                                                                           \
      onfail;                                                                                             \
    }                                                                                                     \
  } while (0)

// C++ exceptions (std::bad_alloc from the host-side vectors) must not cross the extern "C" boundary
template <class R, class F>
static R guarded(R onFail, F &&f) noexcept
{
  try {
    return f();
  } catch (const std::bad_alloc &) {
    set_err("out of host memory");
  } catch (...) {
    set_err("unexpected C++ exception");
  }
  return onFail;
}

static std::atomic<long> g_optTable{0}, g_optWarps{0}, g_optChunkMb{0}, g_optIndex{0}, g_optOverlap{1}, g_optContexts{4}, g_optBatchGroupMb{0};

// hsr_index.cu
bool hsr_parallel_mt_index(const uint8_t *dIn, uint64_t compLen, uint64_t n, int N, int bits, std::vector<hsr_block_t> *out, float *ms);

extern "C" int hsr_version(void) { return HSR_VERSION; }

extern "C" const char *hsr_last_error(void) { return g_err.c_str(); }

extern "C" int hsr_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int hsr_set_device(int device)
{
  CU_TRY(cudaSetDevice(device), return -1);
  return 0;
}

extern "C" int hsr_set_option(const char *key, long value)
{
  if (!key) return -1;
  if (!strcmp(key, "table")) { if (value < 0 || value > 3) return -1; g_optTable = value; return 0; }
  if (!strcmp(key, "warps")) { if (value < 0 || value > 32) return -1; g_optWarps = value; return 0; }
  if (!strcmp(key, "chunk_mb")) { if (value < 0) return -1; g_optChunkMb = value; return 0; }
  if (!strcmp(key, "index")) { if (value < 0 || value > 2) return -1; g_optIndex = value; return 0; }
  if (!strcmp(key, "overlap")) { if (value < 0 || value > 1) return -1; g_optOverlap = value; return 0; }
  if (!strcmp(key, "contexts")) { if (value < 1 || value > 16) return -1; g_optContexts = value; return 0; }
  if (!strcmp(key, "batch_group_mb")) { if (value < 0) return -1; g_optBatchGroupMb = value; return 0; }
  return -1;
}

extern "C" long hsr_get_option(const char *key)
{
  if (!key) return -1;
  if (!strcmp(key, "table")) return g_optTable;
  if (!strcmp(key, "warps")) return g_optWarps;
  if (!strcmp(key, "chunk_mb")) return g_optChunkMb;
  if (!strcmp(key, "index")) return g_optIndex;
  if (!strcmp(key, "overlap")) return g_optOverlap;
  if (!strcmp(key, "contexts")) return g_optContexts;
  if (!strcmp(key, "batch_group_mb")) return g_optBatchGroupMb;
  return -1;
}

extern "C" size_t hsr_capacity(int family, int N, size_t inputSize)
{
  const size_t n = (size_t)N;
  if (family == HSR_RAW) // src/rANS32x32_16w.cpp:10-13: buffer + one row of slack + histogram + states + lengths
    return inputSize + n + sizeof(uint16_t) * 256 + sizeof(uint32_t) * n + sizeof(uint64_t) * 2;
  if (family == HSR_RAW32BLK) // src/rans32x32_32blk_16w.cpp:10-13: + the sub-stream sizes
    return inputSize + n + sizeof(uint16_t) * 256 + sizeof(uint32_t) * n * 2 + sizeof(uint64_t) * 2;
  // src/block_rANS32x32_16w_encode.cpp:47-54 / src/mt_rANS32x64_16w_encode.cpp:50-57: one header per possible
  // block of MinMinBlockSize = 2^15 symbols; mt_ headers also carry the skip offset and a state snapshot
  const size_t base = 2 * sizeof(uint64_t) + 256 * sizeof(uint16_t) + inputSize + n * sizeof(uint32_t);
  const size_t blockCount = (inputSize + ((size_t)1 << 15)) / ((size_t)1 << 15) + 1;
  const size_t perBlock = family == HSR_MT ? sizeof(uint64_t) * 2 + 256 * sizeof(uint16_t) + n * sizeof(uint32_t)
                                           : sizeof(uint64_t) + 256 * sizeof(uint16_t);
  return base + blockCount * perBlock;
}

extern "C" void *hsr_host_alloc(size_t bytes)
{
  void *p = nullptr;
  CU_TRY(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable), return nullptr);
  return p;
}

extern "C" void hsr_host_free(void *p)
{
  if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------------ header + index

static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct Header {
  uint64_t n, compLen;
};

static bool valid_codec(int family, int N, int bits)
{
  if (bits < 10 || bits > 15) return false;
  if (family == HSR_RAW) return N == 16 || N == 32 || N == 64; // rANS32x16_16w, rANS32x32_16w, rANS32x64_16w
  if (family == HSR_RAW32BLK) return N == 32;                   // rANS32x32_32blk_16w
  return (family == HSR_BLOCK || family == HSR_MT) && (N == 32 || N == 64);
}

static inline bool is_raw_family(int family) { return family == HSR_RAW || family == HSR_RAW32BLK; }

// bytes in front of the first word: lengths, counts, states (+ the 31 sub-stream sizes of the 32blk layout,
// src/rans32x32_32blk_16w.cpp:185)
static inline uint64_t fixed_header_bytes(int family, int N)
{
  return 16 + 512 + 4 * (uint64_t)(family == HSR_RAW32BLK ? 2 * N - 1 : N);
}

// the checks every reference decoder starts with (src/rANS32x32_16w.cpp:164-180)
static bool read_header(int family, int N, const uint8_t *in, size_t inLength, size_t outCapacity, Header *h)
{
  if (!in || inLength < fixed_header_bytes(family, N)) { set_err("input shorter than the fixed header"); return false; }
  h->n = rd64(in);
  h->compLen = rd64(in + 8);
  if (h->n > outCapacity) { set_err("decoded length %llu exceeds outCapacity %zu", (unsigned long long)h->n, outCapacity); return false; }
  if (inLength < h->compLen) { set_err("inLength %zu < compressed length %llu", inLength, (unsigned long long)h->compLen); return false; }
  if (h->n < (uint64_t)N) { set_err("decoded length below the state count is undefined in the reference"); return false; }
  return true;
}

constexpr uint64_t kFillUnit = 4ull << 20; // single-symbol runs are cut into fills of this size
constexpr uint64_t kMaxUnitIn = 0xfff00000ull; // per-unit compressed bytes must fit the ring's 32-bit cursor

constexpr uint64_t kMaxDecoded = 1ull << 40; // 1 TiB: bounds the unit count a hostile header can ask for

// One pass over the mt_ header chain (src/mt_rANS32x64_16w_decode.cpp:40-66,94 — the serial walk the reference does on
// its calling thread). Every unit is handed to `emit`; returns the number of units or -1.
template <class Emit>
static long mt_walk(int N, const uint8_t *in, size_t inLength, Emit &&emit, hsr_block_t *lastCodedOut)
{
  if (!(N == 32 || N == 64)) { set_err("state count must be 32 or 64"); return -1; }
  if (!in || inLength < 16 + 4 * (size_t)N + 512) { set_err("input shorter than the fixed header"); return -1; }
  const uint64_t n = rd64(in);
  if (n < (uint64_t)N) { set_err("decoded length below the state count"); return -1; }
  if (n > kMaxDecoded) { set_err("decoded length above 1 TiB is not supported"); return -1; }
  const uint64_t outLengthInStates = n - N + 1;
  uint64_t pos = 16, i = 0;
  size_t count = 0;
  hsr_block_t pending{}; // the last coded block is held back: a ragged end of the stream becomes its tail
  bool havePending = false;
  auto flush = [&]() { if (havePending) { emit(pending, count - 1); havePending = false; } };
  do {
    if (pos + 8 > inLength) { set_err("mt_ chain runs past the input at offset %llu", (unsigned long long)pos); return -1; }
    const uint64_t v = rd64(in + pos);
    if (v >> 63) { // single-symbol run (src/mt_rANS32x64_16w_decode.cpp:46-54)
      const uint64_t size = v & ((1ull << 54) - 1);
      if (size > n - i) { set_err("single-symbol run overruns the decoded length"); return -1; }
      flush();
      for (uint64_t o = 0; o < size; o += kFillUnit) {
        hsr_block_t b{};
        b.inOffset = pos; b.inEnd = pos + 8; b.outOffset = i + o; b.count = std::min(kFillUnit, size - o);
        b.kind = 1; b.symbol = (uint32_t)(v >> 54) & 0xffu;
        emit(b, count);
        count++;
      }
      pos += 8;
      i += size;
    } else {
      if (pos + 16 + 4 * (uint64_t)N + 512 > inLength) { set_err("mt_ block header runs past the input"); return -1; }
      const uint64_t skip = rd64(in + pos + 8);
      if (skip >= (inLength - (pos + 16)) / 2) { set_err("mt_ skip offset runs past the input"); return -1; }
      uint64_t after = pos + 16 + 2 * (skip + 1); // :59
      uint64_t end = i + v; // :77-82
      if (end > outLengthInStates) end = outLengthInStates;
      else if (end & (uint64_t)(N - 1)) { set_err("mt_ block end not a multiple of the state count"); return -1; }
      const uint64_t rows = end > i ? (end - i + N - 1) / N : 0;
      // The encoder measures the first block it writes (the LAST of the chain) from its last word slot rather
      // than one past it (src/mt_rANS32x64_16w_encode.cpp:163,279), so that block's skip is 2 bytes short and
      // nothing reads it; its words simply run to the end of the stream.
      if (!(i + rows * N < outLengthInStates)) after = inLength;
      if (after < pos + 16 + 4 * (uint64_t)N + 512 || after > inLength) { set_err("mt_ skip offset inconsistent"); return -1; }
      if (after - pos > kMaxUnitIn) { set_err("mt_ block larger than 4 GiB compressed is not supported"); return -1; }
      flush();
      pending = hsr_block_t{};
      pending.inOffset = pos + 16; pending.inEnd = after; pending.outOffset = i; pending.count = rows * N; pending.kind = 0;
      havePending = true;
      count++;
      pos = after;
      i += rows * N;
    }
  } while (i < outLengthInStates);

  if (i < n) { // leftover < N symbols use the last coded block's states and cursor (:99-130)
    if (!havePending) { set_err("mt_ stream ends in a tail without a coded block"); return -1; }
    pending.tail = (uint32_t)(n - i);
    pending.count += n - i;
  }
  if (lastCodedOut && havePending) *lastCodedOut = pending;
  flush();
  return (long)count;
}

extern "C" long hsr_mt_index(int N, const uint8_t *in, size_t inLength, hsr_block_t *blocks, size_t maxBlocks)
{
  return mt_walk(N, in, inLength, [&](const hsr_block_t &b, size_t at) { if (blocks && at < maxBlocks) blocks[at] = b; }, nullptr);
}

// the same walk into a vector, one pass
static bool mt_index_vector(int N, const uint8_t *in, size_t inLength, std::vector<hsr_block_t> *out)
{
  out->clear();
  out->reserve((size_t)std::min<uint64_t>(inLength / 32768 + 64, 1u << 22));
  return mt_walk(N, in, inLength, [&](const hsr_block_t &b, size_t) { out->push_back(b); }, nullptr) >= 0;
}

extern "C" int hsr_mt_partition(const hsr_block_t *blocks, size_t count, int parts, size_t *firstUnit)
{
  if (!blocks || parts < 1 || !firstUnit) return -1;
  uint64_t total = 0;
  for (size_t k = 0; k < count; k++) total += (blocks[k].inEnd - blocks[k].inOffset) + blocks[k].count;
  size_t k = 0;
  uint64_t acc = 0;
  firstUnit[0] = 0;
  for (int p = 1; p < parts; p++) {
    const uint64_t target = total / parts * p + (total % parts) * p / parts;
    while (k < count && acc + ((blocks[k].inEnd - blocks[k].inOffset) + blocks[k].count) / 2 <= target) {
      acc += (blocks[k].inEnd - blocks[k].inOffset) + blocks[k].count;
      k++;
    }
    firstUnit[p] = k;
  }
  firstUnit[parts] = count;
  return 0;
}

// ------------------------------------------------------------------------------------------------ kernel launch

static int sm_count()
{
  static std::atomic<int> cached{0};
  int v = cached;
  if (v == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
      (void)cudaGetLastError();
      v = 148;
    }
    cached = v;
  }
  return v;
}

// `decodedBytes` = bytes the launch decodes in total (0 = unknown): the wide tables cost ~40 us to build per unit
static int pick_table(int bits, size_t units = (size_t)-1, uint64_t decodedBytes = 0)
{
  const long opt = g_optTable;
  if (opt == 1) return TK_RANK;
  if (opt == 2) return bits <= 12 ? TK_PACKED : TK_RANK;
  if (opt == 3) return bits >= 13 ? TK_WIDE : TK_PACKED;
  // at most one unit per SM (one raw / block_ stream, a few huge mt_ blocks): every unit can have an SM's whole
  // shared memory, so 13..15 bits also get a one-lookup table (TK_WIDE, 40..160 KB) and only the row latency counts
  // ... provided the units are long enough to pay for filling 2^bits slots (break-even near 100 KB per unit)
  if (bits >= 13 && units <= (size_t)sm_count() && decodedBytes / (units ? units : 1) >= (256u << 10)) return TK_WIDE;
  // measured on B200 (profiles/r1/sweep_1g_v5.jsonl): the packed slot table wins while it leaves >= 19 CTAs per
  // SM resident (4 / 8 KB at 10 / 11 bits); at 12 bits its 16 KB cost more occupancy than the second lookup costs
  // ... unless there are too few units to fill the GPU anyway (a single raw / block_ stream is ONE warp): then
  // only the row latency counts and one lookup on the chain beats two
  if (bits == 12 && units < 148u * 8u) return TK_PACKED;
  return bits <= 11 ? TK_PACKED : TK_RANK;
}

static const KernelEntry &kernel_entry(int family, int N, int bits, int table)
{
  if (family == HSR_RAW32BLK) return kKernelsBlk32[bits - 10][table == TK_RANK ? 0 : 1];
  if (N == 16) return kKernelsRaw16[bits - 10][table == TK_RANK ? 0 : 1];
  return (N == 32 ? kKernels32 : kKernels64)[bits - 10][table - 1];
}

struct LaunchInfo {
  int ctasPerSm = 0, smCount = 0;
};
static inline size_t dynamic_smem(const KernelEntry &ke) { return ke.dynamic ? (size_t)ke.smemBytes : 0; }

static std::mutex g_attrMutex;

// residency of a one-warp-CTA kernel, queried once per (kernel, device)
static bool prepare_kernel(const void *fn, LaunchInfo *li, size_t dynamicSmem = 0)
{
  struct Key { const void *fn; int dev; LaunchInfo li; };
  static std::vector<Key> cache;
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev), return false);
  std::lock_guard<std::mutex> lock(g_attrMutex);
  for (auto &k : cache)
    if (k.fn == fn && k.dev == dev) { *li = k.li; return true; }
  CU_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), return false);
  if (dynamicSmem)
    CU_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynamicSmem), return false);
  LaunchInfo out;
  CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&out.ctasPerSm, fn, 32, dynamicSmem), return false);
  CU_TRY(cudaDeviceGetAttribute(&out.smCount, cudaDevAttrMultiProcessorCount, dev), return false);
  if (out.ctasPerSm < 1) { set_err("kernel does not fit on an SM"); return false; }
  cache.push_back({fn, dev, out});
  *li = out;
  return true;
}

// Launch-private work slots of the units kernels: a ring of {next unit, CTAs that have left} pairs per device, all
// zero at rest — the last CTA of a launch hands its slot back zeroed (units_leave), so no memset sits between two
// launches and overlapping launches (programmatic dependent launch, several host threads, several CUDA streams) never
// share a counter. A units kernel in flight holds at least one resident CTA, and a device holds at most 148 x 32 of
// them, so 8192 slots cannot wrap onto a launch that is still running.
constexpr uint32_t kWorkSlots = 8192;
struct WorkRing {
  int device = -1;
  uint32_t *d = nullptr;
  std::atomic<uint32_t> next{0};
};

static uint32_t *claim_work_slot()
{
  static std::mutex mu;
  static std::vector<std::unique_ptr<WorkRing>> rings;
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev), return nullptr);
  WorkRing *r = nullptr;
  {
    std::lock_guard<std::mutex> lock(mu);
    for (auto &x : rings)
      if (x->device == dev) { r = x.get(); break; }
    if (!r) {
      std::unique_ptr<WorkRing> x(new WorkRing);
      x->device = dev;
      CU_TRY(cudaMalloc(&x->d, kWorkSlots * 2 * sizeof(uint32_t)), return nullptr);
      CU_TRY(cudaMemset(x->d, 0, kWorkSlots * 2 * sizeof(uint32_t)), return nullptr);
      rings.push_back(std::move(x));
      r = rings.back().get();
    }
  }
  return r->d + 2 * (r->next.fetch_add(1, std::memory_order_relaxed) % kWorkSlots);
}

// launches the units kernel over `numBlocks` records of a device-resident index; status bits are OR-ed into *dStatus
static int launch_units(int family, int N, int bits, const uint8_t *dIn, uint64_t inBase, uint8_t *dOut, uint64_t outBase,
                        const hsr_block_t *dBlocks, uint32_t numBlocks, uint32_t *dStatus, cudaStream_t st,
                        uint32_t *dStreamStatus = nullptr, uint64_t decodedBytes = 0)
{
  if (numBlocks == 0) return 0;
  const int table = pick_table(bits, numBlocks, decodedBytes);
  const KernelEntry &ke = kernel_entry(family, N, bits, table);
  uint32_t *dWork = claim_work_slot();
  if (!dWork) return -1;
  DecodeParams p{dIn, inBase, dOut, outBase, dBlocks, numBlocks, dWork, dStatus, dStreamStatus};
  void *args[] = {&p};
  LaunchInfo li;
  if (!prepare_kernel(ke.units, &li, dynamic_smem(ke))) return -1;
  // persistent one-warp CTAs: every SM filled to its residency limit, units handed out by an atomic counter
  const long optWarps = g_optWarps;
  const uint32_t perSm = optWarps > 0 ? std::min<uint32_t>((uint32_t)optWarps, (uint32_t)li.ctasPerSm) : (uint32_t)li.ctasPerSm;
  const uint32_t grid = std::min<uint32_t>(numBlocks, perSm * (uint32_t)li.smCount);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32);
  cfg.dynamicSmemBytes = dynamic_smem(ke);
  cfg.stream = st;
  // back-to-back units launches overlap: this one may start while the previous units kernel of the stream drains
  // (see units_enter in hsr_kernels.cuh); anything else in front of it serialises as usual
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_optOverlap ? 1u : 0u;
  CU_TRY(cudaLaunchKernelExC(&cfg, ke.units, args), return -1);
  return 1;
}

static int launch_block_stream(int N, int bits, const uint8_t *dIn, uint64_t inLength, uint8_t *dOut, uint64_t n,
                               uint32_t *dCounter, cudaStream_t st)
{
  const int table = pick_table(bits, 1, n);
  const KernelEntry &ke = kernel_entry(HSR_BLOCK, N, bits, table);
  BlockStreamParams p{dIn, dOut, nullptr, BlockStreamDesc{0, inLength, 0, n}, 1u, dCounter, nullptr};
  void *args[] = {&p};
  LaunchInfo li;
  if (!prepare_kernel(ke.block, &li, dynamic_smem(ke))) return -1;
  CU_TRY(cudaLaunchKernel(ke.block, dim3(1), dim3(32), args, dynamic_smem(ke), st), return -1);
  return 1;
}

// many independent block_ streams, one warp each
static int launch_block_batch(int N, int bits, const uint8_t *dIn, uint8_t *dOut, const BlockStreamDesc *dStreams, uint32_t count,
                              uint32_t *dCounter, uint32_t *dStreamStatus, cudaStream_t st, uint64_t decodedBytes = 0)
{
  if (count == 0) return 0;
  const int table = pick_table(bits, count, decodedBytes);
  const KernelEntry &ke = kernel_entry(HSR_BLOCK, N, bits, table);
  BlockStreamParams p{dIn, dOut, dStreams, BlockStreamDesc{0, 0, 0, 0}, count, dCounter, dStreamStatus};
  void *args[] = {&p};
  CU_TRY(cudaMemsetAsync(dCounter, 0, 4, st), return -1);
  LaunchInfo li;
  if (!prepare_kernel(ke.block, &li, dynamic_smem(ke))) return -1;
  const uint32_t grid = std::min<uint32_t>(count, (uint32_t)(li.ctasPerSm * li.smCount));
  CU_TRY(cudaLaunchKernel(ke.block, dim3(grid), dim3(32), args, dynamic_smem(ke), st), return -1);
  return 1;
}

// ------------------------------------------------------------------------------------------------ device mt_ walk

// Serial walk of the mt_ chain for streams that only exist in device memory. One warp: lanes 0..7 fetch the
// eight u16 of {size, skip} in one round trip per hop. Emits the same records as hsr_mt_index.
__global__ void mt_walk_kernel(const uint8_t *in, uint64_t inLength, uint32_t N, hsr_block_t *blocks, uint64_t maxBlocks,
                               unsigned long long *result /* [0] count, [1] error */)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t n = ldg_u64_a2(in);
  uint64_t err = 0, count = 0, pos = 16, i = 0;
  long long lastCoded = -1;
  if (n < N) err = 1;
  const uint64_t outLengthInStates = n - N + 1;
  while (!err) {
    if (pos + 8 > inLength) { err = 2; break; }
    uint32_t h = 0;
    if (lane < 8 && pos + 2 * lane + 2 <= inLength) h = ldg_u16(in + pos + 2 * lane);
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = __shfl_sync(kFull, h, k);
    const uint64_t v = (uint64_t)w[0] | ((uint64_t)w[1] << 16) | ((uint64_t)w[2] << 32) | ((uint64_t)w[3] << 48);
    const uint64_t skip = (uint64_t)w[4] | ((uint64_t)w[5] << 16) | ((uint64_t)w[6] << 32) | ((uint64_t)w[7] << 48);
    if (v >> 63) {
      const uint64_t size = v & ((1ull << 54) - 1);
      if (size > n - i) { err = 3; break; }
      for (uint64_t o = 0; o < size; o += kFillUnit) {
        if (lane == 0 && count < maxBlocks) {
          hsr_block_t b{};
          b.inOffset = pos; b.inEnd = pos + 8; b.outOffset = i + o; b.count = size - o < kFillUnit ? size - o : kFillUnit;
          b.kind = 1; b.symbol = (uint32_t)(v >> 54) & 0xffu;
          blocks[count] = b;
        }
        count++;
      }
      pos += 8;
      i += size;
    } else {
      if (pos + 16 + 4ull * N + 512 > inLength) { err = 4; break; }
      if (skip >= (inLength - (pos + 16)) / 2) { err = 5; break; }
      uint64_t after = pos + 16 + 2 * (skip + 1);
      uint64_t end = i + v;
      if (end > outLengthInStates) end = outLengthInStates;
      else if (end & (uint64_t)(N - 1)) { err = 7; break; }
      const uint64_t rows = end > i ? (end - i + N - 1) / N : 0;
      if (!(i + rows * N < outLengthInStates)) after = inLength; // last block: see hsr_mt_index
      if (after < pos + 16 + 4ull * N + 512 || after > inLength || after - pos > kMaxUnitIn) { err = 6; break; }
      if (lane == 0 && count < maxBlocks) {
        hsr_block_t b{};
        b.inOffset = pos + 16; b.inEnd = after; b.outOffset = i; b.count = rows * N; b.kind = 0;
        blocks[count] = b;
      }
      lastCoded = (long long)count;
      count++;
      pos = after;
      i += rows * N;
    }
    if (!(i < outLengthInStates)) break;
  }
  if (!err && i < n) {
    if (lastCoded < 0 || (uint64_t)lastCoded + 1 != count) err = 8;
    else if (lane == 0 && (uint64_t)lastCoded < maxBlocks) {
      blocks[lastCoded].tail = (uint32_t)(n - i);
      blocks[lastCoded].count += n - i;
    }
  }
  if (lane == 0) { result[0] = count; result[1] = err; }
}

// ------------------------------------------------------------------------------------------------ prepared streams

struct hsr_stream {
  int family = 0, N = 0, bits = 0, device = 0;
  uint64_t n = 0, compLen = 0;
  uint8_t *dIn = nullptr;   // device bytes of [inBase, inBase + inBytes)
  bool ownsIn = false;
  uint64_t inBase = 0, inBytes = 0;
  std::vector<hsr_block_t> blocks; // this shard's units (absolute offsets)
  hsr_block_t *dBlocks = nullptr;
  uint32_t *dCounter = nullptr;    // [0] work counter of the block_ kernels, [1] status bits of every kernel
  uint64_t outOffset = 0, outBytes = 0;
  uint64_t decodedTotal = 0; // bytes one decode_async produces (table choice)
  double indexMs = 0;
  // batch of independent streams (hsr_stream_upload_batch): block_ descriptors + per-stream status
  BlockStreamDesc *dDescs = nullptr;
  uint32_t numDescs = 0;
  bool batch = false;
};

static void stream_release(hsr_stream *s)
{
  if (!s) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(s->device);
  if (s->ownsIn && s->dIn) cudaFree(s->dIn);
  if (s->dBlocks) cudaFree(s->dBlocks);
  if (s->dDescs) cudaFree(s->dDescs);
  if (s->dCounter) cudaFree(s->dCounter);
  cudaSetDevice(prev);
  delete s;
}

extern "C" void hsr_stream_free(hsr_stream_t *s) { stream_release(s); }

static bool stream_finish(hsr_stream *s) // uploads the index, allocates the counters
{
  s->decodedTotal = s->batch ? s->outBytes : (s->outBytes ? s->outBytes : s->n);
  if (!s->blocks.empty()) {
    uint64_t sum = 0;
    for (const auto &b : s->blocks) sum += b.count;
    s->decodedTotal = sum;
  }
  CU_TRY(cudaMalloc(&s->dCounter, 16), return false);
  CU_TRY(cudaMemset(s->dCounter, 0, 16), return false);
  if (!s->blocks.empty()) {
    // Units are independent and carry absolute offsets, so the device copy may be in any order: longest first
    // (LPT), so that the persistent warps' last round consists of the short units and the SMs drain together.
    std::vector<hsr_block_t> order(s->blocks);
    std::stable_sort(order.begin(), order.end(), [](const hsr_block_t &a, const hsr_block_t &b) { return a.count > b.count; });
    CU_TRY(cudaMalloc(&s->dBlocks, order.size() * sizeof(hsr_block_t)), return false);
    CU_TRY(cudaMemcpy(s->dBlocks, order.data(), order.size() * sizeof(hsr_block_t), cudaMemcpyHostToDevice), return false);
  }
  return true;
}

static hsr_block_t raw_unit(int family, int N, const Header &h)
{
  hsr_block_t b{};
  b.inOffset = 16; // u16 counts[256], then u32 states[N], then words (src/rANS32x32_16w.cpp:183-200)
  b.inEnd = h.compLen;
  b.outOffset = 0;
  b.count = h.n;
  b.kind = family == HSR_RAW32BLK ? 3 : 2; // 3: + u32 blockSize[31], then 32 private sub-streams
  b.tail = (uint32_t)(h.n % (uint64_t)N);
  return b;
}

// Work list of a batch of independent streams (hsr_decode_batch / hsr_stream_upload_batch).
struct BatchPlan {
  std::vector<Header> hdr;
  std::vector<char> good;
  std::vector<hsr_block_t> units;      // raw / mt_: offsets relative to inBase / outBase
  std::vector<BlockStreamDesc> descs;  // block_: offsets relative to inLo / outLo
  std::vector<uint32_t> descStream;
  uint64_t inLo = ~0ull, inHi = 0, outLo = ~0ull, outHi = 0;
};

static bool plan_batch(int family, int N, const uint8_t *inBase, const hsr_batch_item_t *items, size_t count, BatchPlan *bp)
{
  bp->hdr.assign(count, Header{0, 0});
  bp->good.assign(count, 0);
  size_t nGood = 0;
  // per-stream header checks (src/rANS32x32_16w.cpp:164-180); bad streams get length 0 and are skipped
  for (size_t i = 0; i < count; i++) {
    const hsr_batch_item_t &it = items[i];
    // streams are sequences of 16-bit words: an odd offset would make every 16-bit device load of the stream misaligned
    if (it.inOffset & 1ull) { set_err("stream %zu: inOffset must be even", i); continue; }
    if (!read_header(family, N, inBase + it.inOffset, (size_t)it.inLength, (size_t)it.outCapacity, &bp->hdr[i])) continue;
    if (bp->hdr[i].compLen < fixed_header_bytes(family, N) || bp->hdr[i].compLen > kMaxUnitIn) continue;
    bp->good[i] = 1;
    nGood++;
    bp->inLo = std::min<uint64_t>(bp->inLo, it.inOffset & ~15ull);
    bp->inHi = std::max<uint64_t>(bp->inHi, it.inOffset + bp->hdr[i].compLen);
    bp->outLo = std::min<uint64_t>(bp->outLo, it.outOffset);
    bp->outHi = std::max<uint64_t>(bp->outHi, it.outOffset + bp->hdr[i].n);
  }
  if (nGood == 0) { set_err("no well-formed stream in the batch"); return false; }
  for (size_t i = 0; i < count; i++) {
    if (!bp->good[i]) continue;
    const hsr_batch_item_t &it = items[i];
    if (is_raw_family(family)) {
      hsr_block_t u = raw_unit(family, N, bp->hdr[i]);
      u.inOffset += it.inOffset; u.inEnd += it.inOffset; u.outOffset += it.outOffset; u.reserved = (uint32_t)i;
      bp->units.push_back(u);
    } else if (family == HSR_MT) {
      const long cnt = hsr_mt_index(N, inBase + it.inOffset, (size_t)bp->hdr[i].compLen, nullptr, 0);
      if (cnt < 0) { bp->good[i] = 0; continue; }
      const size_t at = bp->units.size();
      bp->units.resize(at + (size_t)cnt);
      hsr_mt_index(N, inBase + it.inOffset, (size_t)bp->hdr[i].compLen, bp->units.data() + at, (size_t)cnt);
      for (size_t k = at; k < bp->units.size(); k++) {
        bp->units[k].inOffset += it.inOffset; bp->units[k].inEnd += it.inOffset; bp->units[k].outOffset += it.outOffset;
        bp->units[k].reserved = (uint32_t)i;
      }
    } else {
      bp->descs.push_back(BlockStreamDesc{it.inOffset - bp->inLo, bp->hdr[i].compLen, it.outOffset - bp->outLo, bp->hdr[i].n});
      bp->descStream.push_back((uint32_t)i);
    }
  }
  return true;
}

static hsr_stream_t *stream_upload_impl(int family, int N, int bits, const uint8_t *in, size_t inLength, int shard, int shards)
{
  g_err.clear();
  if (!valid_codec(family, N, bits)) { set_err("unsupported codec (family %d, N %d, bits %d)", family, N, bits); return nullptr; }
  if (shards < 1 || shard < 0 || shard >= shards) { set_err("bad shard %d of %d", shard, shards); return nullptr; }
  if (family != HSR_MT && shards != 1) { set_err("only mt_ streams shard; raw and block_ are one recurrence"); return nullptr; }
  Header h;
  if (!read_header(family, N, in, inLength, (size_t)-1, &h)) return nullptr;
  if (h.compLen < fixed_header_bytes(family, N)) { set_err("compressed length field too small"); return nullptr; }
  if (h.n > kMaxDecoded) { set_err("decoded length above 1 TiB is not supported"); return nullptr; }

  std::unique_ptr<hsr_stream, void (*)(hsr_stream *)> s(new hsr_stream, stream_release);
  s->family = family; s->N = N; s->bits = bits; s->n = h.n; s->compLen = h.compLen;
  CU_TRY(cudaGetDevice(&s->device), return nullptr);

  uint64_t lo = 0, hi = h.compLen;
  if (family == HSR_MT) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<hsr_block_t> all;
    if (!mt_index_vector(N, in, (size_t)h.compLen, &all)) return nullptr;
    s->indexMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::vector<size_t> first((size_t)shards + 1);
    hsr_mt_partition(all.data(), all.size(), shards, first.data());
    s->blocks.assign(all.begin() + first[shard], all.begin() + first[shard + 1]);
    if (!s->blocks.empty()) {
      lo = s->blocks.front().inOffset & ~15ull;
      hi = s->blocks.back().inEnd;
      s->outOffset = s->blocks.front().outOffset;
      s->outBytes = s->blocks.back().outOffset + s->blocks.back().count - s->outOffset;
    } else {
      lo = hi = 0;
    }
  } else {
    if (is_raw_family(family)) {
      if (h.compLen > kMaxUnitIn) { set_err("raw streams above 4 GiB compressed are not supported"); return nullptr; }
      s->blocks.push_back(raw_unit(family, N, h));
    }
    s->outOffset = 0;
    s->outBytes = h.n;
  }
  s->inBase = lo;
  s->inBytes = hi - lo;
  if (s->inBytes) {
    CU_TRY(cudaMalloc(&s->dIn, (size_t)s->inBytes + 16), return nullptr);
    s->ownsIn = true;
    CU_TRY(cudaMemcpy(s->dIn, in + lo, (size_t)s->inBytes, cudaMemcpyHostToDevice), return nullptr);
  }
  if (!stream_finish(s.get())) return nullptr;
  return s.release();
}

static hsr_stream_t *stream_from_device_impl(int family, int N, int bits, const void *dInV, size_t inLength, const hsr_block_t *dIndex,
                                             size_t numUnits)
{
  g_err.clear();
  if (!valid_codec(family, N, bits)) { set_err("unsupported codec (family %d, N %d, bits %d)", family, N, bits); return nullptr; }
  const uint8_t *dIn = static_cast<const uint8_t *>(dInV);
  if (!dIn || (reinterpret_cast<uintptr_t>(dIn) & 15)) { set_err("device input must be 16-byte aligned"); return nullptr; }
  if (inLength < fixed_header_bytes(family, N)) { set_err("input shorter than the fixed header"); return nullptr; }
  uint8_t hdr[16];
  CU_TRY(cudaMemcpy(hdr, dIn, 16, cudaMemcpyDeviceToHost), return nullptr);
  Header h{rd64(hdr), rd64(hdr + 8)};
  if (inLength < h.compLen) { set_err("inLength %zu < compressed length %llu", inLength, (unsigned long long)h.compLen); return nullptr; }
  if (h.n < (uint64_t)N || h.compLen < fixed_header_bytes(family, N)) { set_err("malformed header"); return nullptr; }
  if (h.n > kMaxDecoded) { set_err("decoded length above 1 TiB is not supported"); return nullptr; }
  if (dIndex && family != HSR_MT) { set_err("a block index only applies to mt_ streams"); return nullptr; }

  std::unique_ptr<hsr_stream, void (*)(hsr_stream *)> s(new hsr_stream, stream_release);
  s->family = family; s->N = N; s->bits = bits; s->n = h.n; s->compLen = h.compLen;
  CU_TRY(cudaGetDevice(&s->device), return nullptr);
  s->dIn = const_cast<uint8_t *>(dIn);
  s->ownsIn = false;
  s->inBase = 0;
  s->inBytes = h.compLen;
  s->outOffset = 0;
  s->outBytes = h.n;

  if (is_raw_family(family)) {
    if (h.compLen > kMaxUnitIn) { set_err("raw streams above 4 GiB compressed are not supported"); return nullptr; }
    s->blocks.push_back(raw_unit(family, N, h));
  } else if (family == HSR_MT && dIndex) {
    // the producer's own block table (hsr_encode_mt_device_indexed): copied back and checked record by record — every
    // read and write of the kernels must stay inside [0, compLen) / [0, n) whatever the table claims
    const auto t0 = std::chrono::steady_clock::now();
    if (numUnits == 0 || numUnits > (size_t)(h.compLen / 8)) { set_err("implausible unit count"); return nullptr; }
    s->blocks.resize(numUnits);
    CU_TRY(cudaMemcpy(s->blocks.data(), dIndex, numUnits * sizeof(hsr_block_t), cudaMemcpyDeviceToHost), return nullptr);
    uint64_t at = 0;
    for (size_t k = 0; k < numUnits; k++) {
      const hsr_block_t &b = s->blocks[k];
      const bool coded = b.kind == 0;
      const bool ok = (b.kind == 0 || b.kind == 1) && (b.inOffset & 1ull) == 0 && b.inOffset >= 16 && b.inOffset < b.inEnd && b.inEnd <= h.compLen &&
                      b.inOffset + (coded ? 4ull * N + 512 : 8ull) <= b.inEnd && b.inEnd - b.inOffset <= kMaxUnitIn && b.outOffset == at &&
                      b.count <= h.n - at && b.tail < (uint32_t)N && (coded ? (b.count - b.tail) % (uint64_t)N == 0 && b.count >= b.tail : b.tail == 0) &&
                      (b.tail == 0 || at + b.count == h.n);
      if (!ok) { set_err("block index record %zu is inconsistent with the stream", k); return nullptr; }
      at += b.count;
    }
    if (at != h.n) { set_err("block index does not cover the decoded length"); return nullptr; }
    s->indexMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  } else if (family == HSR_MT) {
    // segment-parallel index first (hsr_index.cu); the serial walk below is the fallback and the error reporter
    const long indexMode = g_optIndex;
    float parallelMs = 0;
    if (indexMode != 1 && hsr_parallel_mt_index(dIn, h.compLen, h.n, N, bits, &s->blocks, &parallelMs)) {
      s->indexMs = parallelMs;
      if (!stream_finish(s.get())) return nullptr;
      return s.release();
    }
    if (indexMode == 2) { set_err("parallel index declined this stream"); return nullptr; }
    unsigned long long *dRes = nullptr;
    CU_TRY(cudaMalloc(&dRes, 16), return nullptr);
    uint64_t cap = std::max<uint64_t>(1024, h.n / 32768 + 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int attempt = 0; attempt < 2; attempt++) {
      hsr_block_t *dB = nullptr;
      CU_TRY(cudaMalloc(&dB, cap * sizeof(hsr_block_t)), { cudaFree(dRes); cudaEventDestroy(e0); cudaEventDestroy(e1); return nullptr; });
      cudaEventRecord(e0);
      mt_walk_kernel<<<1, 32>>>(dIn, h.compLen, (uint32_t)N, dB, cap, dRes);
      cudaEventRecord(e1);
      unsigned long long res[2] = {0, 0};
      CU_TRY(cudaMemcpy(res, dRes, 16, cudaMemcpyDeviceToHost), { cudaFree(dB); cudaFree(dRes); cudaEventDestroy(e0); cudaEventDestroy(e1); return nullptr; });
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      s->indexMs += ms;
      if (res[1]) { set_err("malformed mt_ chain (device walk error %llu)", res[1]); cudaFree(dB); cudaFree(dRes); cudaEventDestroy(e0); cudaEventDestroy(e1); return nullptr; }
      if (res[0] <= cap) {
        s->blocks.resize((size_t)res[0]);
        CU_TRY(cudaMemcpy(s->blocks.data(), dB, s->blocks.size() * sizeof(hsr_block_t), cudaMemcpyDeviceToHost), { cudaFree(dB); cudaFree(dRes); cudaEventDestroy(e0); cudaEventDestroy(e1); return nullptr; });
        cudaFree(dB);
        break;
      }
      cudaFree(dB);
      cap = res[0];
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dRes);
  }
  if (!stream_finish(s.get())) return nullptr;
  return s.release();
}

static hsr_stream_t *stream_upload_batch_impl(int family, int N, int bits, const uint8_t *inBase, const hsr_batch_item_t *items, size_t count)
{
  g_err.clear();
  if (!valid_codec(family, N, bits)) { set_err("unsupported codec (family %d, N %d, bits %d)", family, N, bits); return nullptr; }
  if (!inBase || !items || count == 0 || count > 0x7fffffffull) { set_err("bad batch arguments"); return nullptr; }
  BatchPlan plan;
  if (!plan_batch(family, N, inBase, items, count, &plan)) return nullptr;
  for (size_t i = 0; i < count; i++)
    if (!plan.good[i]) { set_err("stream %zu of the batch is malformed", i); return nullptr; }
  std::unique_ptr<hsr_stream, void (*)(hsr_stream *)> s(new hsr_stream, stream_release);
  s->family = family; s->N = N; s->bits = bits; s->batch = true;
  CU_TRY(cudaGetDevice(&s->device), return nullptr);
  s->n = plan.outHi;           // the output buffer is addressed from outBase: it must hold outHi bytes
  s->compLen = plan.inHi - plan.inLo;
  s->inBase = plan.inLo;
  s->inBytes = plan.inHi - plan.inLo;
  s->outOffset = 0;
  s->outBytes = plan.outHi;
  CU_TRY(cudaMalloc(&s->dIn, (size_t)s->inBytes + 16), return nullptr);
  s->ownsIn = true;
  CU_TRY(cudaMemcpy(s->dIn, inBase + plan.inLo, (size_t)s->inBytes, cudaMemcpyHostToDevice), return nullptr);
  if (family == HSR_BLOCK) {
    for (auto &d : plan.descs) d.outOffset += plan.outLo; // decode_async passes outBase itself
    s->numDescs = (uint32_t)plan.descs.size();
    CU_TRY(cudaMalloc(&s->dDescs, plan.descs.size() * sizeof(BlockStreamDesc)), return nullptr);
    CU_TRY(cudaMemcpy(s->dDescs, plan.descs.data(), plan.descs.size() * sizeof(BlockStreamDesc), cudaMemcpyHostToDevice), return nullptr);
  } else {
    s->blocks = std::move(plan.units);
  }
  if (!stream_finish(s.get())) return nullptr;
  return s.release();
}

extern "C" hsr_stream_t *hsr_stream_upload(int family, int N, int bits, const uint8_t *in, size_t inLength, int shard, int shards)
{
  return guarded<hsr_stream_t *>(nullptr, [&] { return stream_upload_impl(family, N, bits, in, inLength, shard, shards); });
}

extern "C" hsr_stream_t *hsr_stream_from_device(int family, int N, int bits, const void *dIn, size_t inLength)
{
  return guarded<hsr_stream_t *>(nullptr, [&] { return stream_from_device_impl(family, N, bits, dIn, inLength, nullptr, 0); });
}

extern "C" hsr_stream_t *hsr_stream_from_device_indexed(int N, int bits, const void *dIn, size_t inLength, const hsr_block_t *dIndex, size_t numUnits)
{
  if (!dIndex) { set_err("null block index"); return nullptr; }
  return guarded<hsr_stream_t *>(nullptr, [&] { return stream_from_device_impl(HSR_MT, N, bits, dIn, inLength, dIndex, numUnits); });
}

extern "C" hsr_stream_t *hsr_stream_upload_batch(int family, int N, int bits, const uint8_t *inBase, const hsr_batch_item_t *items,
                                                 size_t count)
{
  return guarded<hsr_stream_t *>(nullptr, [&] { return stream_upload_batch_impl(family, N, bits, inBase, items, count); });
}

extern "C" uint64_t hsr_stream_decoded_length(const hsr_stream_t *s) { return s ? s->n : 0; }
extern "C" uint64_t hsr_stream_shard_out_offset(const hsr_stream_t *s) { return s ? s->outOffset : 0; }
extern "C" uint64_t hsr_stream_shard_out_bytes(const hsr_stream_t *s) { return s ? s->outBytes : 0; }
extern "C" uint64_t hsr_stream_shard_in_bytes(const hsr_stream_t *s) { return s ? s->inBytes : 0; }
extern "C" uint64_t hsr_stream_units(const hsr_stream_t *s) { return s ? (s->family == HSR_BLOCK ? (s->batch ? s->numDescs : 1) : s->blocks.size()) : 0; }
extern "C" double hsr_stream_index_ms(const hsr_stream_t *s) { return s ? s->indexMs : 0.0; }

extern "C" int hsr_stream_copy_index(const hsr_stream_t *s, hsr_block_t *blocks, size_t maxBlocks)
{
  if (!s || !blocks) return -1;
  const size_t k = std::min(maxBlocks, s->blocks.size());
  memcpy(blocks, s->blocks.data(), k * sizeof(hsr_block_t));
  return (int)k;
}

extern "C" int hsr_stream_decode_async(hsr_stream_t *s, void *dOutV, size_t outCapacity, unsigned flags, void *cudaStream)
{
  if (!s || !dOutV) { set_err("null stream or output"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(cudaStream);
  uint8_t *dOut = static_cast<uint8_t *>(dOutV);
  const bool local = (flags & HSR_OUT_SHARD_LOCAL) != 0;
  const uint64_t need = local ? s->outBytes : s->n;
  if (outCapacity < need) { set_err("outCapacity %zu < %llu", outCapacity, (unsigned long long)need); return -1; }
  const uint64_t outBase = local ? s->outOffset : 0;
  if (s->family == HSR_BLOCK && s->batch)
    return launch_block_batch(s->N, s->bits, s->dIn, dOut, s->dDescs, s->numDescs, s->dCounter, nullptr, st, s->decodedTotal);
  if (s->family == HSR_BLOCK)
    return launch_block_stream(s->N, s->bits, s->dIn, s->compLen, dOut, s->n, s->dCounter, st);
  return launch_units(s->family, s->N, s->bits, s->dIn, s->inBase, dOut, outBase, s->dBlocks, (uint32_t)s->blocks.size(), s->dCounter + 1, st,
                      nullptr, s->decodedTotal);
}

extern "C" unsigned hsr_stream_status(hsr_stream_t *s)
{
  if (!s) return ~0u;
  uint32_t v[2] = {0, 0};
  if (cudaMemcpy(v, s->dCounter, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { (void)cudaGetLastError(); return ~0u; }
  const uint32_t zero = 0;
  cudaMemcpy(s->dCounter + 1, &zero, 4, cudaMemcpyHostToDevice);
  return v[1];
}

// ------------------------------------------------------------------------------------------------ host-pointer decode

// Scratch of one host-pointer decode in flight: three streams (copy-in, run, copy-out) and device buffers that grow
// on demand and are kept, like the reference harness keeps its three buffers for the whole run (src/main.cpp:125-127).
// Every device has a small POOL of these (option "contexts", default 4): the reference's decoders are re-entrant and
// are called from many threads at once (SURVEY.md §8b "Threading"), so K host threads decoding K streams each lease
// their own context and their copies and kernels overlap on the device; only a K+1-th caller waits.
struct DeviceCtx {
  int device = -1;
  bool busy = false;
  cudaStream_t sIn = nullptr, sRun = nullptr, sOut = nullptr;
  uint8_t *dIn = nullptr; size_t inCap = 0;
  uint8_t *dOut = nullptr; size_t outCap = 0;
  hsr_block_t *dBlocks = nullptr; size_t blocksCap = 0;
  hsr_block_t *hBlocks = nullptr; size_t hBlocksCap = 0; // pinned + mapped: kernels read the index straight from host memory
  uint32_t *dCounters = nullptr; size_t countersCap = 0; // 4 u32 per launch: block_ work counter, status, pad, pad
  uint32_t *hStatus = nullptr; size_t hStatusCap = 0;    // pinned + mapped: per-stream status words of a batch, read by the host
  std::vector<cudaEvent_t> evIn, evRun;
};

struct CtxPool {
  std::mutex mu;
  std::condition_variable cv;
  std::vector<std::unique_ptr<DeviceCtx>> all;
};
static CtxPool &ctx_pool()
{
  static CtxPool *p = new CtxPool; // never destroyed: worker threads may still hold leases at process exit
  return *p;
}

// RAII lease of an idle context of `device`; creates one while fewer than "contexts" exist, otherwise waits
class CtxLease {
 public:
  explicit CtxLease(int device)
  {
    CtxPool &pool = ctx_pool();
    std::unique_lock<std::mutex> lock(pool.mu);
    for (;;) {
      size_t have = 0;
      for (auto &c : pool.all) {
        if (c->device != device) continue;
        have++;
        if (!c->busy) { c->busy = true; ctx_ = c.get(); return; }
      }
      if (have < (size_t)std::max<long>(1, g_optContexts)) {
        std::unique_ptr<DeviceCtx> c(new DeviceCtx);
        c->device = device;
        if (cudaStreamCreateWithFlags(&c->sIn, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->sRun, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->sOut, cudaStreamNonBlocking) != cudaSuccess) {
          set_err("cannot create CUDA streams: %s", cudaGetErrorString(cudaGetLastError()));
          return;
        }
        c->busy = true;
        pool.all.push_back(std::move(c));
        ctx_ = pool.all.back().get();
        return;
      }
      pool.cv.wait(lock);
    }
  }
  ~CtxLease()
  {
    if (!ctx_) return;
    CtxPool &pool = ctx_pool();
    {
      std::lock_guard<std::mutex> lock(pool.mu);
      ctx_->busy = false;
    }
    pool.cv.notify_all();
  }
  CtxLease(const CtxLease &) = delete;
  CtxLease &operator=(const CtxLease &) = delete;
  DeviceCtx *get() const { return ctx_; }

 private:
  DeviceCtx *ctx_ = nullptr;
};

template <class T>
static bool grow(T *&p, size_t &cap, size_t need)
{
  if (need <= cap) return true;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  const size_t want = need + need / 8 + 256;
  CU_TRY(cudaMalloc(&p, want * sizeof(T)), return false);
  cap = want;
  return true;
}

// The block index of a host-pointer decode lives in pinned, device-mapped host memory. A separate H2D copy of it
// would queue on the copy-in engine BEHIND the stream pieces already in flight and hold back the first kernel
// until the whole stream has landed; 48 bytes per block over PCIe are nothing next to a block's decode time.
static bool grow_host_blocks(DeviceCtx *c, size_t need)
{
  if (need <= c->hBlocksCap) return true;
  if (c->hBlocks) cudaFreeHost(c->hBlocks);
  c->hBlocks = nullptr; c->hBlocksCap = 0;
  const size_t want = need + need / 4 + 1024;
  CU_TRY(cudaHostAlloc(&c->hBlocks, want * sizeof(hsr_block_t), cudaHostAllocMapped | cudaHostAllocPortable), return false);
  c->hBlocksCap = want;
  return true;
}

static bool ensure_events(DeviceCtx *c, size_t nChunks)
{
  while (c->evIn.size() < nChunks) {
    cudaEvent_t a, b;
    CU_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming), return false);
    CU_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming), return false);
    c->evIn.push_back(a); c->evRun.push_back(b);
  }
  return true;
}

// Host -> device copy of stream bytes [lo, hi) in fixed-size pieces on the copy-in stream; piece k covers bytes up
// to ends[k] and signals evIn[k]. Issued before the block index exists, so the header walk overlaps the DMA.
struct InFlight {
  uint64_t lo = 0;
  std::vector<uint64_t> ends;
  // Pageable input: a copy from unpinned memory blocks its caller until the driver has staged it, so the pieces are
  // issued by a helper thread while the calling thread walks the chain, launches ranges and drains their output —
  // otherwise the whole input would go up (65 ms per GB) before the first output byte came back (77 ms per GB): 142 ms
  // instead of ~80. `recorded` counts the pieces whose event has been recorded: a stream must not be told to wait for an
  // event that has not been recorded yet (that wait would be a no-op).
  std::thread helper;
  std::mutex mu;
  std::condition_variable cv;
  size_t recorded = 0;
  bool threaded = false, failed = false;
  std::string error;

  InFlight() = default;
  InFlight(const InFlight &) = delete;
  InFlight &operator=(const InFlight &) = delete;
  ~InFlight() { join(); }
  void join() { if (helper.joinable()) helper.join(); }
  // true once piece `k` has been handed to the copy-in stream and its event recorded
  bool wait_recorded(size_t k)
  {
    if (!threaded) return true;
    std::unique_lock<std::mutex> lock(mu);
    cv.wait(lock, [&] { return failed || recorded > k; });
    if (failed) { set_err("%s", error.c_str()); return false; }
    return true;
  }
};

// HSR_TRACE_PIPELINE=1: device timestamps of every copy piece and range of one host-pointer decode, printed to stderr
// when the call ends (diagnostics for the three-stream pipeline; timing events are only created when it is set).
struct PipeTrace {
  bool on = false;
  cudaEvent_t t0 = nullptr;
  struct Mark { const char *what; size_t index; uint64_t bytes; cudaEvent_t ev; };
  std::vector<Mark> marks;
  static bool enabled() { static const bool e = getenv("HSR_TRACE_PIPELINE") != nullptr; return e; }
  void start(cudaStream_t s) { on = enabled(); if (on) { cudaEventCreate(&t0); cudaEventRecord(t0, s); } }
  void mark(const char *what, size_t index, uint64_t bytes, cudaStream_t s)
  {
    if (!on) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s);
    marks.push_back(Mark{what, index, bytes, e});
  }
  void dump()
  {
    if (!on) return;
    for (const Mark &m : marks) {
      float ms = 0; cudaEventElapsedTime(&ms, t0, m.ev);
      fprintf(stderr, "[hsr pipeline] %8.3f ms  %-10s %3zu  %llu bytes\n", ms, m.what, m.index, (unsigned long long)m.bytes);
      cudaEventDestroy(m.ev);
    }
    cudaEventDestroy(t0);
    marks.clear(); on = false;
  }
};
static thread_local PipeTrace g_trace;

static bool host_memory_is_pageable(const void *p)
{
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { (void)cudaGetLastError(); return true; }
  return attr.type == cudaMemoryTypeUnregistered;
}

static bool start_h2d(DeviceCtx *c, const uint8_t *in, uint64_t lo, uint64_t hi, InFlight *fl)
{
  // pieces grow 2, 2, 4, 8, 16, 16, ... MiB: the first decode launch waits for 2 MiB instead of 16 (0.04 ms instead of
  // 0.3 ms of DMA), so the device -> host stream — the longer of the two directions — starts that much earlier
  const long optMb = g_optChunkMb;
  const uint64_t fixedPiece = optMb > 0 ? (uint64_t)optMb << 20 : 0;
  fl->lo = lo;
  fl->ends.clear();
  if (!grow(c->dIn, c->inCap, (size_t)(hi - lo) + 16)) return false;
  for (uint64_t a = lo; a < hi;) {
    const size_t k = fl->ends.size();
    const uint64_t piece = fixedPiece ? fixedPiece : (2ull << 20) << std::min<size_t>(k > 0 ? k - 1 : 0, 3);
    a = std::min(hi, a + piece);
    fl->ends.push_back(a);
  }
  if (!ensure_events(c, fl->ends.size())) return false;
  uint8_t *dIn = c->dIn;
  cudaStream_t sIn = c->sIn;
  const std::vector<cudaEvent_t> &evIn = c->evIn; // not resized again before the helper has finished (join in finish / ~InFlight)
  auto issue = [=, &evIn](size_t k, bool trace) -> cudaError_t {
    const uint64_t a = k ? fl->ends[k - 1] : lo, b = fl->ends[k];
    cudaError_t e = cudaMemcpyAsync(dIn + (a - lo), in + a, (size_t)(b - a), cudaMemcpyHostToDevice, sIn);
    if (e == cudaSuccess) e = cudaEventRecord(evIn[k], sIn);
    if (trace) g_trace.mark("h2d end", k, b - a, sIn);
    return e;
  };
  if (hi - lo >= (8ull << 20) && host_memory_is_pageable(in + lo)) {
    fl->threaded = true;
    const int device = c->device;
    fl->helper = std::thread([fl, issue, device] {
      cudaError_t e = cudaSetDevice(device);
      for (size_t k = 0; e == cudaSuccess && k < fl->ends.size(); k++) {
        e = issue(k, false);
        if (e != cudaSuccess) break;
        { std::lock_guard<std::mutex> lock(fl->mu); fl->recorded = k + 1; }
        fl->cv.notify_all();
      }
      if (e != cudaSuccess) {
        { std::lock_guard<std::mutex> lock(fl->mu); fl->failed = true; fl->error = std::string("host -> device copy failed: ") + cudaGetErrorString(e); }
        fl->cv.notify_all();
      }
    });
    return true;
  }
  g_trace.start(c->sIn);
  for (size_t k = 0; k < fl->ends.size(); k++) {
    const cudaError_t e = issue(k, true);
    if (e != cudaSuccess) { set_err("host -> device copy failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return false; }
  }
  return true;
}

// Decodes contiguous unit ranges whose compressed bytes are arriving through an InFlight copy: a range is launched as
// soon as the copy piece holding its last byte has landed, and its decoded bytes go back to the host while later
// ranges are still decoding (three streams: copy-in, run, copy-out). Ranges can be handed over while the header chain
// is still being walked (launch() from inside the walk), so the first decoded bytes leave the device after ~0.1 ms
// instead of after the whole 1.7 ms-per-GB walk; range sizes grow 6, 12, 24, 48, 48, ... MiB of traffic for the same
// reason.
struct UnitPipeline {
  DeviceCtx *c = nullptr;
  int family = 0, N = 0, bits = 0;
  uint8_t *out = nullptr;
  uint64_t outLo = 0;
  InFlight *fl = nullptr;
  size_t piece = 0, ranges = 0;

  static uint64_t range_bytes(size_t r) { return std::min<uint64_t>(48ull << 20, (6ull << 20) << std::min<size_t>(r, 8)); }
  uint64_t next_range_bytes() const { return range_bytes(ranges); }

  // decoded bytes [outLo_, outLo_ + outBytes) of the stream will be produced
  bool begin(DeviceCtx *ctx, int family_, int N_, int bits_, uint8_t *out_, uint64_t outLo_, uint64_t outBytes, InFlight *fl_)
  {
    c = ctx; family = family_; N = N_; bits = bits_; out = out_; outLo = outLo_; fl = fl_;
    piece = 0; ranges = 0;
    if (!grow(c->dOut, c->outCap, (size_t)outBytes + 16)) return false;
    if (!grow(c->dCounters, c->countersCap, 4)) return false;
    CU_TRY(cudaMemsetAsync(c->dCounters, 0, 16, c->sRun), return false); // [1]: status bits, OR over every range
    return true;
  }

  // units[0 .. count): device-visible (mapped host memory), consecutive in the stream
  bool launch(const hsr_block_t *units, size_t count)
  {
    if (count == 0) return true;
    while (c->evRun.size() <= ranges) {
      cudaEvent_t e;
      CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), return false);
      c->evRun.push_back(e);
    }
    const uint64_t needEnd = units[count - 1].inEnd;
    while (piece + 1 < fl->ends.size() && fl->ends[piece] < needEnd) piece++;
    if (!fl->wait_recorded(piece)) return false;
    CU_TRY(cudaStreamWaitEvent(c->sRun, c->evIn[piece], 0), return false);
    uint64_t rangeDecoded = 0;
    for (size_t k = 0; k < count; k++) rangeDecoded += units[k].count;
    g_trace.mark("run start", ranges, count, c->sRun);
    if (launch_units(family, N, bits, c->dIn, fl->lo, c->dOut, outLo, units, (uint32_t)count, c->dCounters + 1, c->sRun, nullptr, rangeDecoded) < 0)
      return false;
    CU_TRY(cudaEventRecord(c->evRun[ranges], c->sRun), return false);
    g_trace.mark("run end", ranges, rangeDecoded, c->sRun);
    CU_TRY(cudaStreamWaitEvent(c->sOut, c->evRun[ranges], 0), return false);
    const uint64_t oLo = units[0].outOffset, oHi = units[count - 1].outOffset + units[count - 1].count;
    g_trace.mark("d2h start", ranges, 0, c->sOut);
    CU_TRY(cudaMemcpyAsync(out + oLo, c->dOut + (oLo - outLo), (size_t)(oHi - oLo), cudaMemcpyDeviceToHost, c->sOut), return false);
    g_trace.mark("d2h end", ranges, oHi - oLo, c->sOut);
    ranges++;
    return true;
  }

  // waits for everything issued so far; false if a copy failed or a decoded unit was malformed
  bool finish()
  {
    uint32_t status[4] = {0, 0, 0, 0};
    bool ok = true;
    fl->join(); // pageable input: every piece has been issued once the helper is done
    if (fl->failed) ok = false;
    // the status word is read on the copy-out stream, which has waited for every launch
    if (cudaMemcpyAsync(status, c->dCounters, 16, cudaMemcpyDeviceToHost, c->sOut) != cudaSuccess) ok = false;
    if (cudaStreamSynchronize(c->sOut) != cudaSuccess) ok = false;
    if (cudaStreamSynchronize(c->sIn) != cudaSuccess) ok = false;
    if (cudaStreamSynchronize(c->sRun) != cudaSuccess) ok = false;
    g_trace.dump();
    if (!ok) { set_err("CUDA stream failed: %s", cudaGetErrorString(cudaGetLastError())); return false; }
    if (status[1]) { set_err("malformed stream (device status 0x%x)", status[1]); return false; }
    return true;
  }
};

// units [first, last) of a complete index, compressed bytes arriving through `fl`
static bool run_units_pipelined(DeviceCtx *c, int family, int N, int bits, uint8_t *out, const hsr_block_t *units, size_t first,
                                size_t last, InFlight &fl)
{
  if (first >= last) return true;
  const uint64_t outLo = units[first].outOffset, outHi = units[last - 1].outOffset + units[last - 1].count;
  const size_t count = last - first;
  const hsr_block_t *devBlocks = nullptr; // device-visible address of units[first]
  if (units >= c->hBlocks && units + last <= c->hBlocks + c->hBlocksCap) {
    devBlocks = units + first; // already in the mapped buffer (UVA: same address on the device)
  } else {
    if (!grow_host_blocks(c, count)) return false;
    memcpy(c->hBlocks, units + first, count * sizeof(hsr_block_t));
    devBlocks = c->hBlocks;
  }
  UnitPipeline pipe;
  if (!pipe.begin(c, family, N, bits, out, outLo, outHi - outLo, &fl)) return false;
  size_t a = 0;
  uint64_t acc = 0;
  for (size_t k = 0; k < count; k++) {
    acc += (devBlocks[k].inEnd - devBlocks[k].inOffset) + devBlocks[k].count;
    if (acc >= pipe.next_range_bytes() && k + 1 < count) {
      if (!pipe.launch(devBlocks + a, k + 1 - a)) { pipe.finish(); return false; }
      a = k + 1;
      acc = 0;
    }
  }
  const bool launched = pipe.launch(devBlocks + a, count - a);
  return pipe.finish() && launched;
}

// units [first, last) of an already indexed stream, from host memory, on `device`
static bool decode_units_from_host(int device, int family, int N, int bits, const uint8_t *in, uint8_t *out, const hsr_block_t *units,
                                   size_t first, size_t last)
{
  if (first >= last) return true;
  CU_TRY(cudaSetDevice(device), return false);
  CtxLease lease(device);
  DeviceCtx *c = lease.get();
  if (!c) return false;
  InFlight fl;
  if (!start_h2d(c, in, units[first].inOffset & ~15ull, units[last - 1].inEnd, &fl)) return false;
  return run_units_pipelined(c, family, N, bits, out, units, first, last, fl);
}

static bool decode_block_from_host(int device, int N, int bits, const uint8_t *in, const Header &h, uint8_t *out)
{
  CU_TRY(cudaSetDevice(device), return false);
  CtxLease lease(device);
  DeviceCtx *c = lease.get();
  if (!c) return false;
  if (!grow(c->dIn, c->inCap, (size_t)h.compLen + 16)) return false;
  if (!grow(c->dOut, c->outCap, (size_t)h.n + 16)) return false;
  if (!grow(c->dCounters, c->countersCap, 4)) return false;
  CU_TRY(cudaMemsetAsync(c->dCounters, 0, 16, c->sRun), return false);
  CU_TRY(cudaMemcpyAsync(c->dIn, in, (size_t)h.compLen, cudaMemcpyHostToDevice, c->sRun), return false);
  if (launch_block_stream(N, bits, c->dIn, h.compLen, c->dOut, h.n, c->dCounters, c->sRun) < 0) return false;
  CU_TRY(cudaMemcpyAsync(out, c->dOut, (size_t)h.n, cudaMemcpyDeviceToHost, c->sRun), return false);
  uint32_t status[4] = {0, 0, 0, 0};
  CU_TRY(cudaMemcpyAsync(status, c->dCounters, 16, cudaMemcpyDeviceToHost, c->sRun), return false);
  CU_TRY(cudaStreamSynchronize(c->sRun), return false);
  if (status[1]) { set_err("malformed stream (device status 0x%x)", status[1]); return false; }
  return true;
}

static size_t decode_impl(int family, int N, int bits, const uint8_t *in, size_t inLength, uint8_t *out, size_t outCapacity)
{
  g_err.clear();
  if (!valid_codec(family, N, bits)) { set_err("unsupported codec (family %d, N %d, bits %d)", family, N, bits); return 0; }
  Header h;
  if (!read_header(family, N, in, inLength, outCapacity, &h)) return 0;
  if (!out) { set_err("null output"); return 0; }
  if (h.compLen < fixed_header_bytes(family, N)) { set_err("compressed length field too small"); return 0; }
  int device = 0;
  CU_TRY(cudaGetDevice(&device), return 0);

  if (family == HSR_BLOCK)
    return decode_block_from_host(device, N, bits, in, h, out) ? (size_t)h.n : 0;
  if (is_raw_family(family)) {
    if (h.compLen > kMaxUnitIn) { set_err("raw streams above 4 GiB compressed are not supported"); return 0; }
    const hsr_block_t u = raw_unit(family, N, h);
    return decode_units_from_host(device, family, N, bits, in, out, &u, 0, 1) ? (size_t)h.n : 0;
  }
  // mt_: start moving the whole stream to the device and walk the header chain on the host meanwhile; every range of
  // units is launched from inside the walk, as soon as the walk has passed it
  CtxLease lease(device);
  DeviceCtx *c = lease.get();
  if (!c) return 0;
  InFlight fl;
  if (!start_h2d(c, in, 0, h.compLen, &fl)) return 0;
  if (!grow_host_blocks(c, (size_t)std::max<uint64_t>(64, h.n / 32768 + 64))) { cudaStreamSynchronize(c->sIn); return 0; }
  UnitPipeline pipe;
  if (!pipe.begin(c, family, N, bits, out, 0, h.n, &fl)) { cudaStreamSynchronize(c->sIn); return 0; }
  size_t launched = 0, have = 0;
  uint64_t acc = 0;
  bool overflow = false, failed = false;
  std::string launchErr;
  const long cnt = mt_walk(N, in, (size_t)h.compLen, [&](const hsr_block_t &b, size_t at) {
    if (overflow || failed) return;
    if (at >= c->hBlocksCap) { overflow = true; return; } // more units than the estimate: finished below, after a re-walk
    c->hBlocks[at] = b;
    have = at + 1;
    acc += (b.inEnd - b.inOffset) + b.count;
    if (acc >= pipe.next_range_bytes()) {
      if (!pipe.launch(c->hBlocks + launched, have - launched)) { failed = true; launchErr = g_err; }
      launched = have;
      acc = 0;
    }
  }, nullptr);
  if (cnt >= 0 && !failed && !overflow && !pipe.launch(c->hBlocks + launched, have - launched)) { failed = true; launchErr = g_err; }
  const std::string walkErr = g_err;
  const bool drained = pipe.finish();
  if (cnt < 0) { g_err = walkErr; return 0; } // ranges already decoded wrote only bytes of well-formed blocks; the call still fails
  if (failed) { g_err = launchErr; return 0; }
  if (!drained) return 0;
  if (overflow) { // rare: tiny blocks / long fills. The compressed bytes are all on the device by now.
    if (!grow_host_blocks(c, (size_t)cnt)) return 0;
    if (hsr_mt_index(N, in, (size_t)h.compLen, c->hBlocks, c->hBlocksCap) != cnt) return 0;
    if (!run_units_pipelined(c, family, N, bits, out, c->hBlocks, launched, (size_t)cnt, fl)) return 0;
  }
  return (size_t)h.n;
}

extern "C" size_t hsr_decode(int family, int N, int bits, const uint8_t *in, size_t inLength, uint8_t *out, size_t outCapacity)
{
  return guarded<size_t>(0, [&] { return decode_impl(family, N, bits, in, inLength, out, outCapacity); });
}

// Many independent streams of one codec in ONE launch: raw and block_ streams are a single recurrence each (one
// warp), so a batch is the only way they fill a GPU; mt_ streams simply contribute all their blocks.
static size_t decode_batch_impl(int family, int N, int bits, const uint8_t *inBase, uint8_t *outBase, const hsr_batch_item_t *items,
                                size_t count, uint64_t *decodedLengths)
{
  g_err.clear();
  if (!valid_codec(family, N, bits)) { set_err("unsupported codec (family %d, N %d, bits %d)", family, N, bits); return 0; }
  if (!inBase || !outBase || !items || !decodedLengths) { set_err("null argument"); return 0; }
  if (count == 0) return 0;
  if (count > 0x7fffffffull) { set_err("too many streams"); return 0; }
  for (size_t i = 0; i < count; i++) decodedLengths[i] = 0; // every early return below leaves "nothing decoded"
  int device = 0;
  CU_TRY(cudaGetDevice(&device), return 0);
  CtxLease lease(device);
  DeviceCtx *c = lease.get();
  if (!c) return 0;

  BatchPlan plan;
  if (!plan_batch(family, N, inBase, items, count, &plan)) return 0;
  std::vector<Header> &hdr = plan.hdr;
  std::vector<char> &good = plan.good;
  const uint64_t inLo = plan.inLo, inHi = plan.inHi, outLo = plan.outLo, outHi = plan.outHi;

  // The batch runs as a pipeline like a single mt_ stream does: streams are taken in input order and cut into groups
  // of growing traffic (UnitPipeline::range_bytes); the whole input range goes to the device in pieces, every group is
  // launched as soon as the piece holding its last byte has landed, and its decoded bytes go back while later groups
  // decode. Per-stream status words live in mapped host memory, so the host can tell which streams of a finished group
  // decoded cleanly (only those are copied back) without a device -> host copy in between.
  std::vector<uint32_t> order; // well-formed streams by input offset
  for (size_t i = 0; i < count; i++)
    if (good[i]) order.push_back((uint32_t)i);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return items[x].inOffset < items[y].inOffset; });
  std::vector<size_t> unitFirst(count + 1, 0); // plan.units / plan.descs hold the good streams in stream order
  {
    size_t at = 0, d = 0;
    for (size_t i = 0; i < count; i++) {
      unitFirst[i] = family == HSR_BLOCK ? d : at;
      if (!good[i]) continue;
      if (family == HSR_BLOCK) d++;
      else while (at < plan.units.size() && plan.units[at].reserved == (uint32_t)i) at++;
    }
    unitFirst[count] = family == HSR_BLOCK ? d : at;
  }
  struct Group { size_t sa, sb, ua, ub; uint64_t needEnd, decoded; };
  std::vector<Group> groups;
  std::vector<hsr_block_t> su;          // units in launch order
  std::vector<BlockStreamDesc> sd;      // block_ streams in launch order (position == index into `order`)
  {
    // A group's kernel cannot finish before its longest unit has been decoded by ONE warp (~0.4-0.8 GB/s), however
    // few units it holds, and consecutive groups run one after another: a group must therefore carry enough bytes for
    // its copies (~50 GB/s) to outlast that floor — 512 x its longest unit — or many small groups would serialise
    // their floors (2368 raw streams of 400 KB in 37 groups: 35 ms; in 8 groups: the copies' 21 ms).
    const long optGroupMb = g_optBatchGroupMb; // tests: a fixed group size instead of the rule above
    Group g{0, 0, 0, 0, 0, 0};
    uint64_t acc = 0, longest = 0;
    for (size_t k = 0; k < order.size(); k++) {
      const uint32_t i = order[k];
      if (family == HSR_BLOCK) {
        sd.push_back(plan.descs[unitFirst[i]]);
        longest = std::max<uint64_t>(longest, hdr[i].n);
      } else {
        for (size_t u = unitFirst[i]; u < unitFirst[i + 1]; u++) {
          su.push_back(plan.units[u]);
          if (plan.units[u].kind != 1u) longest = std::max<uint64_t>(longest, plan.units[u].count);
        }
      }
      g.needEnd = std::max<uint64_t>(g.needEnd, items[i].inOffset + hdr[i].compLen);
      g.decoded += hdr[i].n;
      acc += hdr[i].compLen + hdr[i].n;
      const uint64_t target = optGroupMb > 0 ? (uint64_t)optGroupMb << 20 : std::max<uint64_t>(UnitPipeline::range_bytes(groups.size()), 512 * longest);
      if (acc >= target || k + 1 == order.size()) {
        g.sb = k + 1;
        g.ub = family == HSR_BLOCK ? k + 1 : su.size();
        if (family != HSR_BLOCK) // longest units first inside a group
          std::stable_sort(su.begin() + (long)g.ua, su.begin() + (long)g.ub, [](const hsr_block_t &x, const hsr_block_t &y) { return x.count > y.count; });
        groups.push_back(g);
        g = Group{k + 1, k + 1, g.ub, g.ub, 0, 0};
        acc = 0;
        longest = 0;
      }
    }
  }

  InFlight fl;
  auto fail = [&]() -> size_t { // a CUDA call failed: drain what is in flight and report nothing decoded
    fl.join();
    cudaStreamSynchronize(c->sIn); cudaStreamSynchronize(c->sRun); cudaStreamSynchronize(c->sOut);
    for (size_t i = 0; i < count; i++) decodedLengths[i] = 0;
    return 0;
  };
  // the unit list goes up FIRST: a copy queued behind the input pieces would wait for all of them on the copy engine
  // (and with it every launch: measured, the first group then starts after the whole 14 ms of input)
  const size_t listBytes = family == HSR_BLOCK ? sd.size() * sizeof(BlockStreamDesc) : su.size() * sizeof(hsr_block_t);
  if (!grow(c->dBlocks, c->blocksCap, listBytes / sizeof(hsr_block_t) + 1)) return 0;
  CU_TRY(cudaMemcpyAsync(c->dBlocks, family == HSR_BLOCK ? (const void *)sd.data() : (const void *)su.data(), listBytes, cudaMemcpyHostToDevice, c->sRun),
         return fail());
  CU_TRY(cudaStreamSynchronize(c->sRun), return fail()); // su / sd are pageable: the copy has read them when this returns
  if (!start_h2d(c, inBase, inLo, inHi, &fl)) return fail();
  if (!grow(c->dOut, c->outCap, (size_t)(outHi - outLo) + 16)) return fail();
  if (!grow(c->dCounters, c->countersCap, 4)) return fail();
  if (count > c->hStatusCap) {
    if (c->hStatus) cudaFreeHost(c->hStatus);
    c->hStatus = nullptr; c->hStatusCap = 0;
    const size_t want = count + count / 4 + 256;
    CU_TRY(cudaHostAlloc(&c->hStatus, want * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable), return fail());
    c->hStatusCap = want;
  }
  memset(c->hStatus, 0, count * sizeof(uint32_t)); // raw / mt_: indexed by stream; block_: by position in `order`
  CU_TRY(cudaMemsetAsync(c->dCounters, 0, 16, c->sRun), return fail());
  while (c->evRun.size() < groups.size()) {
    cudaEvent_t e;
    CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), return fail());
    c->evRun.push_back(e);
  }
  size_t piece = 0;
  for (size_t gi = 0; gi < groups.size(); gi++) {
    const Group &g = groups[gi];
    while (piece + 1 < fl.ends.size() && fl.ends[piece] < g.needEnd) piece++;
    if (!fl.wait_recorded(piece)) return fail();
    CU_TRY(cudaStreamWaitEvent(c->sRun, c->evIn[piece], 0), return fail());
    int rc;
    if (family == HSR_BLOCK)
      rc = launch_block_batch(N, bits, c->dIn, c->dOut, reinterpret_cast<const BlockStreamDesc *>(c->dBlocks) + g.sa, (uint32_t)(g.sb - g.sa),
                              c->dCounters, c->hStatus + g.sa, c->sRun, g.decoded);
    else
      rc = launch_units(family, N, bits, c->dIn, inLo, c->dOut, outLo, c->dBlocks + g.ua, (uint32_t)(g.ub - g.ua), c->dCounters + 1, c->sRun,
                        c->hStatus, g.decoded);
    if (rc < 0) return fail();
    CU_TRY(cudaEventRecord(c->evRun[gi], c->sRun), return fail());
    g_trace.mark("run end", gi, g.decoded, c->sRun);
  }

  // as each group finishes: copy back the streams that decoded cleanly, merging runs whose outputs are contiguous
  size_t ok = 0;
  for (size_t gi = 0; gi < groups.size(); gi++) {
    const Group &g = groups[gi];
    CU_TRY(cudaEventSynchronize(c->evRun[gi]), return fail());
    uint64_t runLo = 0, runHi = 0;
    bool open = false;
    auto flush = [&]() -> bool {
      if (!open) return true;
      CU_TRY(cudaMemcpyAsync(outBase + runLo, c->dOut + (runLo - outLo), (size_t)(runHi - runLo), cudaMemcpyDeviceToHost, c->sOut), return false);
      g_trace.mark("d2h end", gi, runHi - runLo, c->sOut);
      open = false;
      return true;
    };
    for (size_t k = g.sa; k < g.sb; k++) {
      const uint32_t i = order[k];
      const uint32_t st = family == HSR_BLOCK ? c->hStatus[k] : c->hStatus[i];
      if (st) continue;
      decodedLengths[i] = hdr[i].n;
      ok++;
      const uint64_t lo = items[i].outOffset, hi = lo + hdr[i].n;
      if (open && lo == runHi) { runHi = hi; continue; }
      if (!flush()) return fail();
      runLo = lo; runHi = hi; open = true;
    }
    if (!flush()) return fail();
  }
  CU_TRY(cudaStreamSynchronize(c->sOut), return fail());
  fl.join();
  if (fl.failed) { set_err("%s", fl.error.c_str()); return fail(); }
  CU_TRY(cudaStreamSynchronize(c->sIn), return fail());
  g_trace.dump();
  if (ok != count) set_err("%zu of %zu streams were malformed", count - ok, count);
  return ok;
}

extern "C" size_t hsr_decode_batch(int family, int N, int bits, const uint8_t *inBase, uint8_t *outBase, const hsr_batch_item_t *items,
                                   size_t count, uint64_t *decodedLengths)
{
  return guarded<size_t>(0, [&] { return decode_batch_impl(family, N, bits, inBase, outBase, items, count, decodedLengths); });
}

// One process per GPU: this rank decodes only shard `shard` of `shards` of the chain (the same contiguous block ranges
// as hsr_stream_upload), from the caller's host buffers, into the caller's full-size output buffer at the shard's own
// offset. Only the shard's compressed bytes cross PCIe; the header walk covers the whole chain (it is serial).
static size_t decode_mt_shard_impl(int N, int bits, const uint8_t *in, size_t inLength, uint8_t *out, size_t outCapacity, int shard, int shards,
                                   size_t *shardOffset)
{
  g_err.clear();
  if (shardOffset) *shardOffset = 0;
  if (!valid_codec(HSR_MT, N, bits)) { set_err("unsupported codec (N %d, bits %d)", N, bits); return 0; }
  if (shards < 1 || shard < 0 || shard >= shards) { set_err("bad shard %d of %d", shard, shards); return 0; }
  Header h;
  if (!read_header(HSR_MT, N, in, inLength, outCapacity, &h)) return 0;
  if (!out) { set_err("null output"); return 0; }
  int device = 0;
  CU_TRY(cudaGetDevice(&device), return 0);
  std::vector<hsr_block_t> all;
  if (!mt_index_vector(N, in, (size_t)h.compLen, &all)) return 0;
  std::vector<size_t> first((size_t)shards + 1);
  hsr_mt_partition(all.data(), all.size(), shards, first.data());
  const size_t a = first[(size_t)shard], b = first[(size_t)shard + 1];
  if (a >= b) return 0; // this shard owns no block (fewer blocks than shards): nothing to do, no error
  if (shardOffset) *shardOffset = (size_t)all[a].outOffset;
  if (!decode_units_from_host(device, HSR_MT, N, bits, in, out, all.data(), a, b)) return 0;
  return (size_t)(all[b - 1].outOffset + all[b - 1].count - all[a].outOffset);
}

extern "C" size_t hsr_decode_mt_shard(int N, int bits, const uint8_t *in, size_t inLength, uint8_t *out, size_t outCapacity, int shard,
                                      int shards, size_t *shardOffset)
{
  return guarded<size_t>(0, [&] { return decode_mt_shard_impl(N, bits, in, inLength, out, outCapacity, shard, shards, shardOffset); });
}

// ------------------------------------------------------------------------------------------------ one process, many GPUs

// Persistent host threads for hsr_decode_mt_multi (one per device shard in flight; they outlive the call, so a call
// costs a queue hand-off, not a thread creation per device).
class ShardWorkers {
 public:
  void submit(std::function<void()> f)
  {
    std::unique_lock<std::mutex> lock(mu_);
    queue_.push_back(std::move(f));
    if (idle_ == 0 && threads_ < 64) {
      threads_++;
      std::thread([this] { loop(); }).detach();
    }
    lock.unlock();
    cv_.notify_one();
  }

 private:
  void loop()
  {
    std::unique_lock<std::mutex> lock(mu_);
    for (;;) {
      idle_++;
      cv_.wait(lock, [this] { return !queue_.empty(); });
      idle_--;
      std::function<void()> f = std::move(queue_.front());
      queue_.pop_front();
      lock.unlock();
      f();
      lock.lock();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> queue_;
  int idle_ = 0, threads_ = 0;
};
static ShardWorkers &shard_workers()
{
  static ShardWorkers *w = new ShardWorkers; // leaked on purpose: its threads are detached
  return *w;
}

static size_t decode_mt_multi_impl(int N, int bits, const uint8_t *in, size_t inLength, uint8_t *out, size_t outCapacity,
                                   const int *devices, int deviceCount)
{
  g_err.clear();
  if (!valid_codec(HSR_MT, N, bits)) { set_err("unsupported codec (N %d, bits %d)", N, bits); return 0; }
  if (deviceCount < 1) { set_err("deviceCount must be >= 1"); return 0; }
  Header h;
  if (!read_header(HSR_MT, N, in, inLength, outCapacity, &h)) return 0;
  if (!out) { set_err("null output"); return 0; }

  // ONE walk of the header chain (the reference's serial walk, src/mt_rANS32x64_16w_decode.cpp:137-265). Shards are
  // contiguous unit ranges balanced on compressed + decoded bytes; a shard is handed to its device the moment the walk
  // has passed its last unit, so device 0 is copying after 1/deviceCount of the walk instead of after two full walks.
  struct Shard {
    std::vector<hsr_block_t> units;
    std::string error;
    bool ok = true;
  };
  std::vector<Shard> shards((size_t)deviceCount);
  std::mutex mu;
  std::condition_variable cv;
  int pending = 0;
  auto dispatch = [&](int d) {
    if (shards[(size_t)d].units.empty()) return;
    {
      std::lock_guard<std::mutex> lock(mu);
      pending++;
    }
    shard_workers().submit([&, d] {
      Shard &sh = shards[(size_t)d];
      const int dev = devices ? devices[d] : d;
      sh.ok = decode_units_from_host(dev, HSR_MT, N, bits, in, out, sh.units.data(), 0, sh.units.size());
      if (!sh.ok) sh.error = g_err;
      {
        std::lock_guard<std::mutex> lock(mu);
        pending--;
      }
      cv.notify_all();
    });
  };
  const uint64_t total = h.compLen + h.n;
  int cur = 0;
  uint64_t acc = 0;
  const long cnt = mt_walk(N, in, (size_t)h.compLen, [&](const hsr_block_t &b, size_t) {
    const uint64_t w = (b.inEnd - b.inOffset) + b.count;
    while (cur + 1 < deviceCount && acc + w / 2 > total / (uint64_t)deviceCount * (uint64_t)(cur + 1)) {
      dispatch(cur);
      cur++;
    }
    shards[(size_t)cur].units.push_back(b);
    acc += w;
  }, nullptr);
  if (cnt >= 0)
    for (; cur < deviceCount; cur++) dispatch(cur);
  {
    std::unique_lock<std::mutex> lock(mu);
    cv.wait(lock, [&] { return pending == 0; });
  }
  if (cnt < 0) return 0; // shards already dispatched wrote only bytes of well-formed blocks; the call still fails
  for (int d = 0; d < deviceCount; d++)
    if (!shards[(size_t)d].ok) { set_err("device shard %d: %s", d, shards[(size_t)d].error.c_str()); return 0; }
  return (size_t)h.n;
}

extern "C" size_t hsr_decode_mt_multi(int N, int bits, const uint8_t *in, size_t inLength, uint8_t *out, size_t outCapacity,
                                      const int *devices, int deviceCount)
{
  int prev = 0;
  const bool havePrev = cudaGetDevice(&prev) == cudaSuccess;
  const size_t r = guarded<size_t>(0, [&] { return decode_mt_multi_impl(N, bits, in, inLength, out, outCapacity, devices, deviceCount); });
  if (havePrev) cudaSetDevice(prev);
  return r;
}
