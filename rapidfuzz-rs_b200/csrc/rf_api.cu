// rf_api.cu -- the extern "C" boundary declared in include/rfgpu.h: corpus / batch-comparator handles and
// the scoring entry points.  Plain CUDA runtime; no torch, no CPU fallback (every compute entry point needs a
// CUDA device and reports RF_ERR_CUDA otherwise).
#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <algorithm>
#include <unordered_map>
#include "../../include/rfgpu.h"
#include "rf_kernels.cuh"
#include "rf_internal.h"

using namespace rfk;

static thread_local std::string g_last_error;
// tuning knobs (rf_set_option)
static std::atomic<int> g_build_lb{1};   // build the length-bucketed interleaved layout at corpus creation
static std::atomic<int> g_w1_path{0};    // 0: interleaved-layout kernel when available, 1: CSR/TMA-tile kernel
static std::atomic<int> g_mw_path{0};    // queries of 65..512 on a resident corpus: 0 register kernel (scan_lbn), 1 shuffle kernel (scan_mw)
static std::atomic<int> g_band{1};       // multi-word Levenshtein with cutoff <= 63: banded kernel (0: block kernel)
static std::atomic<int> g_epi_table{1};  // integer metrics, interleaved layout: score algebra as a per-launch table (0: per pair)
static std::atomic<int> g_jaro32{1};     // Jaro / Jaro-Winkler, query <= 32: row-wise 32-bit kernel (0: generic per-lane routine)
// The four knobs above are DEFAULTS: a comparator takes a snapshot of them when it is created (rf_batch::opt) and its
// scoring calls read the snapshot only, so threads working with different settings never race on process-wide state.
static std::atomic<int> g_stream_mb{64};      // rf_batch_stream_*: chunk size in candidate bytes (MiB)
static std::atomic<int> g_stream_kcand{2048}; // rf_batch_stream_*: chunk size in candidates (x1024)
static std::atomic<int> g_compact32{1};       // rf_corpus_create_u32: keep corpora with <= 255 distinct symbols as renamed bytes
static std::atomic<int> g_cdist_slices{0};    // rf_cdist_topk_*: corpus slices (0: automatic)
static std::atomic<int> g_cdist_skip{1};      // rf_cdist_topk_*: skip groups by length against the running k-th bound

void rf__set_sharded_collective(int mode);  // rf_sharded.cu
void rf__set_gather_chunks(int k);

static rf_status fail(rf_status s, const std::string& msg) {
  g_last_error = msg;
  return s;
}
static rf_status cuda_fail(cudaError_t e, const char* what) {
  rf_status s = (e == cudaErrorMemoryAllocation) ? RF_ERR_OOM : RF_ERR_CUDA;
  return fail(s, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define RF_CUDA(call)                                  \
  do {                                                 \
    cudaError_t e__ = (call);                          \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static int sm_count_of(int device) {
  static std::mutex mu;
  static std::vector<int> cache;
  std::lock_guard<std::mutex> lk(mu);
  if ((int)cache.size() <= device) cache.resize(device + 1, 0);
  if (cache[device] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
    cache[device] = v;
  }
  return cache[device];
}

// scratch counters for the chunk scheduler of scan_lb_kernel: a ring of slots per device so that concurrent
// launches never share one
static unsigned long long* counter_slot(int device) {
  static std::mutex mu;
  static std::vector<unsigned long long*> pools;
  static std::vector<uint32_t> next;
  constexpr uint32_t kSlots = 4096;
  std::lock_guard<std::mutex> lk(mu);
  if ((int)pools.size() <= device) { pools.resize(device + 1, nullptr); next.resize(device + 1, 0); }
  if (!pools[device]) {
    if (cudaMalloc(&pools[device], kSlots * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
  }
  return pools[device] + (next[device]++ % kSlots);
}

extern "C" {

void rf_args_default(rf_args* a) {
  if (!a) return;
  memset(a, 0, sizeof(*a));
  a->insertion_cost = a->deletion_cost = a->substitution_cost = 1;
  a->prefix_weight = 0.1;
}

const char* rf_status_string(rf_status s) {
  switch (s) {
    case RF_OK: return "ok";
    case RF_ERR_INVALID_ARG: return "invalid argument";
    case RF_ERR_UNSUPPORTED: return "unsupported";
    case RF_ERR_CUDA: return "CUDA error";
    case RF_ERR_OOM: return "out of device memory";
    case RF_ERR_NCCL: return "NCCL error";
  }
  return "unknown";
}
const char* rf_last_error(void) { return g_last_error.c_str(); }
void rf__set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }  // internal (rf_io.cpp)

int rf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

uint64_t rf_kernel_launch_count(void) { return kernel_launch_count(); }

rf_status rf_set_option(const char* name, int value) {
  if (!name) return fail(RF_ERR_INVALID_ARG, "name is NULL");
  if (!strcmp(name, "build_interleaved_layout")) { g_build_lb.store(value ? 1 : 0); return RF_OK; }
  if (!strcmp(name, "single_word_path")) { g_w1_path.store(value); return RF_OK; }
  if (!strcmp(name, "epilogue_table")) { g_epi_table.store(value ? 1 : 0); return RF_OK; }
  if (!strcmp(name, "jaro32")) { g_jaro32.store(value >= 0 && value <= 3 ? value : 1); return RF_OK; }
  if (!strcmp(name, "multi_word_path")) { g_mw_path.store(value ? 1 : 0); return RF_OK; }
  if (!strcmp(name, "banded_levenshtein")) { g_band.store(value ? 1 : 0); return RF_OK; }
  if (!strcmp(name, "stream_chunk_mb")) { if (value < 1) return fail(RF_ERR_INVALID_ARG, "stream_chunk_mb < 1"); g_stream_mb.store(value); return RF_OK; }
  if (!strcmp(name, "stream_chunk_kcand")) { if (value < 1) return fail(RF_ERR_INVALID_ARG, "stream_chunk_kcand < 1"); g_stream_kcand.store(value); return RF_OK; }
  if (!strcmp(name, "cdist_slices")) { if (value < 0 || value > 256) return fail(RF_ERR_INVALID_ARG, "cdist_slices not in 0..256"); g_cdist_slices.store(value); return RF_OK; }
  if (!strcmp(name, "compact_u32_corpus")) { g_compact32.store(value ? 1 : 0); return RF_OK; }
  if (!strcmp(name, "cdist_skip")) { g_cdist_skip.store(value ? 1 : 0); return RF_OK; }
  if (!strcmp(name, "sharded_collective")) { if (value < 0 || value > 2) return fail(RF_ERR_INVALID_ARG, "sharded_collective not in 0..2"); rf__set_sharded_collective(value); return RF_OK; }
  if (!strcmp(name, "allgather_chunks")) { if (value < 0 || value > 16) return fail(RF_ERR_INVALID_ARG, "allgather_chunks not in 0..16"); rf__set_gather_chunks(value); return RF_OK; }
  return fail(RF_ERR_INVALID_ARG, std::string("unknown option: ") + name);
}

int rf_result_is_float(rf_metric metric, rf_kind kind) { return result_is_float((int)metric, (int)kind) ? 1 : 0; }

// ------------------------------------------------------------------------------------------------ corpus
static rf_status corpus_alloc(rf_corpus* c, bool off64, cudaStream_t st) {
  // 64 bytes of zeroed slack behind the chars (TMA copies whole 16-byte lines, readers over-read one word);
  // 16 entries of slack behind the offsets.
  RF_CUDA(dev_alloc(&c->d_chars, c->total + 64, st));
  RF_CUDA(cudaMemsetAsync(c->d_chars + c->total, 0, 64, st));
  if (off64) RF_CUDA(dev_alloc(&c->d_off64, (c->n + 1 + 16) * sizeof(uint64_t), st));
  else RF_CUDA(dev_alloc(&c->d_off32, (c->n + 1 + 16) * sizeof(uint32_t), st));
  return RF_OK;
}

__global__ void fill_tail_u32(uint32_t* p, uint64_t from, uint32_t v) { p[from + threadIdx.x] = v; }
__global__ void fill_tail_u64(uint64_t* p, uint64_t from, uint64_t v) { p[from + threadIdx.x] = v; }
__global__ void narrow_offsets(const uint64_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n1) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = in[i] > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)in[i];  // saturate: check_offsets then rejects it (> total)
}
// offsets of a sub-range of a larger CSR (sharded corpora): out = in - base, as u32 or u64; values below the base wrap to
// huge ones and saturate, which check_offsets rejects
__global__ void rebase_offsets(const uint64_t* __restrict__ in, uint64_t base, uint32_t* __restrict__ out32,
                               uint64_t* __restrict__ out64, uint64_t n1) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t v = in[i] - base;
    if (out64) out64[i] = v;
    else out32[i] = v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v;
  }
}

// CSR sanity: starts must be non-decreasing and end at `total`.  A corrupt / crafted offset array would otherwise send
// the layout builder and the scan kernels out of bounds (sticky CUDA error = the whole process poisoned).
__global__ void check_offsets_kernel(const uint32_t* __restrict__ o32, const uint64_t* __restrict__ o64, uint64_t n,
                                     uint64_t total, uint32_t* __restrict__ bad) {
  uint32_t b = 0;
  unsigned long long mx = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t a = o64 ? o64[i] : (uint64_t)o32[i], z = o64 ? o64[i + 1] : (uint64_t)o32[i + 1];
    if (z < a || z > total) b = 1;
    else if (z - a > mx) mx = z - a;
  }
  if (__any_sync(0xffffffffu, b) && (threadIdx.x & 31u) == 0) atomicOr(bad, 1u);
  for (int d = 16; d >= 1; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, mx, d);
    mx = o > mx ? o : mx;
  }
  if ((threadIdx.x & 31u) == 0 && mx) atomicMax(reinterpret_cast<unsigned long long*>(bad) + 1, mx);  // longest candidate
}

// synchronises `st`; RF_ERR_INVALID_ARG when the offsets are not a valid CSR index of `total` elements
static rf_status check_offsets(const void* d_off, bool is64, uint64_t n, uint64_t total, cudaStream_t st,
                               uint64_t* max_len = nullptr) {
  if (max_len) *max_len = 0;
  if (n == 0) return RF_OK;
  uint32_t* d_bad = nullptr;
  RF_CUDA(dev_alloc(&d_bad, 16, st));
  cudaError_t e = cudaMemsetAsync(d_bad, 0, 16, st);
  uint32_t bad = 0;
  if (e == cudaSuccess) {
    const uint64_t blocks = (n + 255) / 256;
    const uint32_t grid = (uint32_t)(blocks < 148 * 16 ? blocks : 148 * 16);
    check_offsets_kernel<<<grid, 256, 0, st>>>(is64 ? nullptr : (const uint32_t*)d_off, is64 ? (const uint64_t*)d_off : nullptr, n,
                                               total, d_bad);
    rfk::count_launches(1);
    e = cudaGetLastError();
  }
  unsigned long long res[2] = {0, 0};
  if (e == cudaSuccess) e = cudaMemcpyAsync(res, d_bad, 16, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  bad = (uint32_t)res[0];
  if (max_len) *max_len = res[1];
  dev_free(d_bad, st);
  if (e != cudaSuccess) return cuda_fail(e, "offset validation");
  if (bad) return fail(RF_ERR_INVALID_ARG, "offsets are not non-decreasing CSR starts ending at offsets[n]");
  return RF_OK;
}

static rf_status corpus_finish(rf_corpus* c, cudaStream_t st) {
  {
    rf_status vs = check_offsets(c->d_off32 ? (const void*)c->d_off32 : (const void*)c->d_off64, c->d_off32 == nullptr, c->n,
                                 c->total, st, &c->max_len);
    if (vs != RF_OK) return vs;
  }
  if (c->d_off32) fill_tail_u32<<<1, 16, 0, st>>>(c->d_off32, c->n + 1, (uint32_t)c->total);
  else fill_tail_u64<<<1, 16, 0, st>>>(c->d_off64, c->n + 1, c->total);
  RF_CUDA(cudaGetLastError());
  if (g_build_lb.load()) {
    CorpusView v{c->d_chars, c->d_off32, c->d_off64, c->n, c->total};
    RF_CUDA(lb_build(v, st, &c->lb));
  }
  return RF_OK;
}

// base != 0: offsets[0..n] are a sub-range of a larger CSR index (u64 only); chars is the larger array's start
static rf_status corpus_create_host(const uint8_t* chars, const void* offsets, bool in64, uint64_t n, int device,
                                    rf_corpus** out, uint64_t base = 0) {
  if (!out) return fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!offsets) return fail(RF_ERR_INVALID_ARG, "offsets is NULL");
  if (n >= 0xFFFFFFFFull) return fail(RF_ERR_UNSUPPORTED, "more than 2^32-2 candidates in one corpus");
  const uint64_t first = in64 ? ((const uint64_t*)offsets)[0] : ((const uint32_t*)offsets)[0];
  const uint64_t last = in64 ? ((const uint64_t*)offsets)[n] : ((const uint32_t*)offsets)[n];
  if (first != base) return fail(RF_ERR_INVALID_ARG, "offsets[0] must be 0");
  if (last < base) return fail(RF_ERR_INVALID_ARG, "offsets are not non-decreasing CSR starts ending at offsets[n]");
  const uint64_t total = last - base;
  if (total && !chars) return fail(RF_ERR_INVALID_ARG, "chars is NULL");
  chars = chars ? chars + base : chars;
  if (rf_device_count() <= device || device < 0) return fail(RF_ERR_CUDA, "no such CUDA device");
  DeviceGuard g(device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  rf_corpus* c = new (std::nothrow) rf_corpus();
  if (!c) return fail(RF_ERR_OOM, "host allocation failed");
  c->device = device;
  c->n = n;
  c->total = total;
  const bool off64 = total >= 0xFFFFFFF0ull;
  cudaStream_t st;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { rf_corpus_destroy(c); return fail(RF_ERR_CUDA, "stream"); }
  rf_status s = corpus_alloc(c, off64, st);
  if (s != RF_OK) { cudaStreamSynchronize(st); cudaStreamDestroy(st); rf_corpus_destroy(c); return s; }
  cudaError_t e = cudaSuccess;
  if (total) e = cudaMemcpyAsync(c->d_chars, chars, total, cudaMemcpyHostToDevice, st);
  uint64_t* tmp64 = nullptr;
  if (e == cudaSuccess) {
    if (base != 0) {  // sub-range (always u64 on the host): upload, then subtract the base on the GPU
      e = dev_alloc(&tmp64, (n + 1) * 8, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(tmp64, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) {
        rebase_offsets<<<1024, 256, 0, st>>>(tmp64, base, c->d_off32, c->d_off64, n + 1);
        e = cudaGetLastError();
      }
    } else if (off64 == in64) {
      e = cudaMemcpyAsync(off64 ? (void*)c->d_off64 : (void*)c->d_off32, offsets, (n + 1) * (in64 ? 8 : 4),
                          cudaMemcpyHostToDevice, st);
    } else if (in64) {  // u64 on the host, u32 on the device: upload then narrow on the GPU
      e = dev_alloc(&tmp64, (n + 1) * 8, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(tmp64, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) {
        narrow_offsets<<<1024, 256, 0, st>>>(tmp64, c->d_off32, n + 1);
        e = cudaGetLastError();
      }
    } else {
      e = cudaErrorInvalidValue;  // u32 host offsets cannot describe >= 4 GiB
    }
  }
  if (e == cudaSuccess) {
    s = corpus_finish(c, st);
    if (s == RF_OK) e = cudaStreamSynchronize(st);
  }
  dev_free(tmp64, st);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  if (e != cudaSuccess) { rf_corpus_destroy(c); return cuda_fail(e, "corpus upload"); }
  if (s != RF_OK) { rf_corpus_destroy(c); return s; }
  *out = c;
  return RF_OK;
}

rf_status rf_corpus_create_u8(const uint8_t* chars, const uint64_t* offsets, uint64_t n, int device, rf_corpus** out) {
  return corpus_create_host(chars, offsets, true, n, device, out);
}
rf_status rf_corpus_create_u8_off32(const uint8_t* chars, const uint32_t* offsets, uint64_t n, int device, rf_corpus** out) {
  return corpus_create_host(chars, offsets, false, n, device, out);
}

rf_status rf_corpus_create_device_u8(const uint8_t* d_chars, const uint64_t* d_offsets, uint64_t n, uint64_t total_chars,
                                     int device, void* stream, rf_corpus** out) {
  if (!out) return fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!d_offsets || (total_chars && !d_chars)) return fail(RF_ERR_INVALID_ARG, "NULL device buffer");
  if (n >= 0xFFFFFFFFull) return fail(RF_ERR_UNSUPPORTED, "more than 2^32-2 candidates in one corpus");
  if (rf_device_count() <= device || device < 0) return fail(RF_ERR_CUDA, "no such CUDA device");
  DeviceGuard g(device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  rf_corpus* c = new (std::nothrow) rf_corpus();
  if (!c) return fail(RF_ERR_OOM, "host allocation failed");
  c->device = device;
  c->n = n;
  c->total = total_chars;
  const bool off64 = total_chars >= 0xFFFFFFF0ull;
  cudaStream_t st = (cudaStream_t)stream;
  rf_status s = corpus_alloc(c, off64, st);
  if (s != RF_OK) { cudaStreamSynchronize(st); rf_corpus_destroy(c); return s; }
  cudaError_t e = cudaSuccess;
  if (total_chars) e = cudaMemcpyAsync(c->d_chars, d_chars, total_chars, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) {
    if (off64) e = cudaMemcpyAsync(c->d_off64, d_offsets, (n + 1) * 8, cudaMemcpyDeviceToDevice, st);
    else {
      narrow_offsets<<<1024, 256, 0, st>>>(d_offsets, c->d_off32, n + 1);
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess) {
    s = corpus_finish(c, st);
    if (s == RF_OK) e = cudaStreamSynchronize(st);
  }
  if (e != cudaSuccess) { rf_corpus_destroy(c); return cuda_fail(e, "corpus device copy"); }
  if (s != RF_OK) { rf_corpus_destroy(c); return s; }
  *out = c;
  return RF_OK;
}

static rf_status compact_u32_corpus(rf_corpus* c, cudaStream_t st);

rf_status rf_corpus_create_u32(const uint32_t* elems, const uint64_t* offsets, uint64_t n, int device, rf_corpus** out) {
  if (!out) return fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!offsets) return fail(RF_ERR_INVALID_ARG, "offsets is NULL");
  if (n >= 0xFFFFFFFFull) return fail(RF_ERR_UNSUPPORTED, "more than 2^32-2 candidates in one corpus");
  if (offsets[0] != 0) return fail(RF_ERR_INVALID_ARG, "offsets[0] must be 0");
  const uint64_t total = offsets[n];
  if (total && !elems) return fail(RF_ERR_INVALID_ARG, "elems is NULL");
  if (rf_device_count() <= device || device < 0) return fail(RF_ERR_CUDA, "no such CUDA device");
  DeviceGuard g(device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  rf_corpus* c = new (std::nothrow) rf_corpus();
  if (!c) return fail(RF_ERR_OOM, "host allocation failed");
  c->device = device;
  c->n = n;
  c->total = total;
  const bool off64 = total >= 0xFFFFFFF0ull;
  cudaStream_t st = util_stream(device);
  cudaError_t e = dev_alloc(&c->d_elems32, (total + 16) * sizeof(uint32_t), st);
  if (e == cudaSuccess) e = off64 ? dev_alloc(&c->d_off64, (n + 1 + 16) * sizeof(uint64_t), st)
                                  : dev_alloc(&c->d_off32, (n + 1 + 16) * sizeof(uint32_t), st);
  uint64_t* tmp64 = nullptr;
  if (e == cudaSuccess && total) e = cudaMemcpyAsync(c->d_elems32, elems, total * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    if (off64) e = cudaMemcpyAsync(c->d_off64, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st);
    else {
      e = dev_alloc(&tmp64, (n + 1) * 8, st);
      if (e == cudaSuccess) e = cudaMemcpyAsync(tmp64, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) {
        narrow_offsets<<<1024, 256, 0, st>>>(tmp64, c->d_off32, n + 1);
        e = cudaGetLastError();
      }
    }
  }
  if (e == cudaSuccess) {
    if (c->d_off32) fill_tail_u32<<<1, 16, 0, st>>>(c->d_off32, n + 1, (uint32_t)total);
    else fill_tail_u64<<<1, 16, 0, st>>>(c->d_off64, n + 1, total);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  dev_free(tmp64, st);
  if (e != cudaSuccess) { rf_corpus_destroy(c); return cuda_fail(e, "u32 corpus upload"); }
  {
    rf_status vs = check_offsets(off64 ? (const void*)c->d_off64 : (const void*)c->d_off32, off64, n, total, st, &c->max_len);
    if (vs != RF_OK) { rf_corpus_destroy(c); return vs; }
  }
  {  // an unsigned element >= 2^31 shares its bit pattern with a negative signed one (see rf_corpus_create_elems)
    int hug = 0;
#pragma omp parallel for schedule(static) reduction(| : hug)
    for (int64_t i = 0; i < (int64_t)total; ++i) hug |= (int)(elems[i] >> 31);
    c->has_huge = hug != 0;
  }
  if (g_compact32.load() && total) {
    rf_status s = compact_u32_corpus(c, st);
    if (s != RF_OK) { rf_corpus_destroy(c); return s; }
  }
  *out = c;
  return RF_OK;
}

rf_status rf_corpus_destroy(rf_corpus* c) {
  if (!c) return RF_OK;
  DeviceGuard g(c->device);
  // the *_device entry points are asynchronous on the caller's stream: a scan may still be reading this corpus, and
  // the stream-ordered frees below are ordered on an internal stream only -> drain the device first (destroy is rare)
  cudaDeviceSynchronize();
  cudaStream_t st = util_stream(c->device);
  dev_free(c->d_chars, st);
  dev_free(c->d_elems32, st);
  dev_free(c->d_off32, st);
  dev_free(c->d_off64, st);
  lb_free(&c->lb, st);
  delete c;
  return RF_OK;
}
// Frees the CSR copy (characters + starts) of a byte corpus that also holds the interleaved layout: 3.6 + 0.4 of the 8.8 GB
// BASELINE config 2 occupies.  Everything the interleaved layout serves keeps working (see rfgpu.h); the rest is refused.
rf_status rf_corpus_release_csr(rf_corpus* c) {
  if (!c) return fail(RF_ERR_INVALID_ARG, "NULL handle");
  if (c->d_elems32) return fail(RF_ERR_UNSUPPORTED, "a u32 corpus with more than 255 distinct symbols has no interleaved layout to fall back on");
  if (c->n && !c->lb.gdata) return fail(RF_ERR_UNSUPPORTED, "the corpus was created without the interleaved layout (build_interleaved_layout = 0)");
  DeviceGuard g(c->device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  cudaDeviceSynchronize();  // asynchronous *_device scans may still read the copy (as in rf_corpus_destroy)
  cudaStream_t st = util_stream(c->device);
  dev_free(c->d_chars, st);
  dev_free(c->d_off32, st);
  dev_free(c->d_off64, st);
  c->d_chars = nullptr;
  c->d_off32 = nullptr;
  c->d_off64 = nullptr;
  c->csr_released = true;
  return RF_OK;
}
int rf_corpus_has_csr(const rf_corpus* c) { return c && !c->csr_released ? 1 : 0; }
uint64_t rf_corpus_size(const rf_corpus* c) { return c ? c->n : 0; }
uint64_t rf_corpus_total_chars(const rf_corpus* c) { return c ? c->total : 0; }
int rf_corpus_device(const rf_corpus* c) { return c ? c->device : -1; }

// ------------------------------------------------------------------------------------------------ batch
}  // extern "C"

static constexpr uint32_t kAlphaSlots = 1024;
static inline uint32_t alpha_hash(uint32_t x) { return (x * 2654435761u) >> 22; }  // 10 bits

// host -> device copy that has LANDED when it returns (stream-ordered on the device's utility stream + a stream sync)
static cudaError_t upload_sync(void* dst, const void* src, size_t bytes, int device) {
  cudaStream_t st = util_stream(device);
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return e;
}

// nomatch (optional, [query_len]): positions flagged 1 hold a symbol no candidate element can equal (a u32 query symbol outside
// the corpus' byte / dictionary domain): they set no bit in any match table, which is exact for every table-driven metric.
static rf_status batch_create_bytes(rf_metric metric, const uint8_t* query, uint32_t query_len, int device, rf_batch** out,
                                    const uint8_t* nomatch = nullptr) {
  if (!out) return fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if ((int)metric < 0 || (int)metric > (int)RF_DAMERAU_LEVENSHTEIN) return fail(RF_ERR_INVALID_ARG, "unknown metric");
  if (query_len && !query) return fail(RF_ERR_INVALID_ARG, "query is NULL");
  if (query_len > RF_MAX_QUERY_LEN) return fail(RF_ERR_UNSUPPORTED, "query longer than RF_MAX_QUERY_LEN");
  if (rf_device_count() <= device || device < 0) return fail(RF_ERR_CUDA, "no such CUDA device");
  DeviceGuard g(device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  rf_batch* b = new (std::nothrow) rf_batch();
  if (!b) return fail(RF_ERR_OOM, "host allocation failed");
  b->device = device;
  b->metric = metric;
  b->opt = rf_batch_opts{g_w1_path.load(), g_mw_path.load(), g_band.load(), g_jaro32.load(), g_epi_table.load()};
  b->s1.assign(query, query + query_len);
  b->len1 = query_len;
  b->words = (query_len + 63) / 64;
  // Pattern-match tables (pattern_match_vector.rs:213-224: bit i%64 of PM[ch][i/64] set iff q[i]==ch),
  // in the layouts the kernels want.  One blob: [tab32_top | tab32_bot | tab64_top | tab64_bot | pm_words]
  const uint32_t words = b->words ? b->words : 1;
  const size_t sz32 = 256 * sizeof(uint32_t), sz64 = 256 * sizeof(uint64_t);
  const size_t szw = (size_t)256 * words * sizeof(uint64_t);
  const uint32_t bstride = (2 * (words + 2)) | 1u;  // u32 units, odd
  const size_t szp = (size_t)256 * bstride * sizeof(uint32_t);
  const size_t szp8 = (szp + 7) & ~(size_t)7;
  const size_t szq = (size_t)kQuotDim * kQuotDim * sizeof(double);  // exact a/b for a,b <= 64 (Jaro epilogue)
  const size_t szb = ((size_t)query_len + 15) / 16 * 16 + 16;      // the query bytes themselves (hamming / prefix / postfix)
  // queries of 65..512 elements: the match vectors as integers of `limbs` 32-bit limbs (scan_lbn_kernel)
  const uint32_t limbs = (query_len > 64 && query_len <= 512) ? 4 * ((query_len + 127) / 128) : 0;
  const size_t szn = (size_t)256 * limbs * sizeof(uint32_t);
  const size_t off_n = 2 * sz32 + 2 * sz64 + szw + szp8 + szq + szb;
  std::vector<uint8_t> blob(off_n + 2 * szn, 0);
  if (query_len) memcpy(blob.data() + 2 * sz32 + 2 * sz64 + szw + szp8 + szq, query, query_len);
  uint32_t* t32t = (uint32_t*)blob.data();
  uint32_t* t32b = t32t + 256;
  uint64_t* t64t = (uint64_t*)(blob.data() + 2 * sz32);
  uint64_t* t64b = t64t + 256;
  uint64_t* pmw = t64b + 256;
  uint32_t* pmb = (uint32_t*)(pmw + (size_t)256 * words);
  for (uint32_t i = 0; i < query_len; ++i) {
    if (nomatch && nomatch[i]) continue;
    pmw[(size_t)query[i] * words + i / 64] |= 1ull << (i % 64);
    pmb[(size_t)query[i] * bstride + 2 + i / 32] |= 1u << (i % 32);
  }
  if (query_len >= 1 && query_len <= 64) {
    for (int ch = 0; ch < 256; ++ch) {
      const uint64_t m = pmw[(size_t)ch * words];
      t64b[ch] = m;
      t64t[ch] = m << (64 - query_len);
      if (query_len <= 32) {
        t32b[ch] = (uint32_t)m;
        t32t[ch] = (uint32_t)m << (32 - query_len);
      }
    }
  }
  {
    double* quot = (double*)(blob.data() + 2 * sz32 + 2 * sz64 + szw + szp8);
    for (int a = 0; a < kQuotDim; ++a)
      for (int d = 1; d < kQuotDim; ++d) quot[a * kQuotDim + d] = (double)a / (double)d;
  }
  if (limbs) {
    uint32_t* top = (uint32_t*)(blob.data() + off_n);
    uint32_t* bot = top + (size_t)256 * limbs;
    const uint32_t sh = 32 * limbs - query_len, ws = sh / 32, bs = sh % 32;
    for (int ch = 0; ch < 256; ++ch) {
      uint32_t* bt = bot + (size_t)ch * limbs;
      for (uint32_t j = 0; j < limbs && j < 2 * words; ++j) bt[j] = (uint32_t)(pmw[(size_t)ch * words + j / 2] >> (32 * (j % 2)));
      uint32_t* tp = top + (size_t)ch * limbs;
      for (uint32_t i = ws; i < limbs; ++i) {
        const uint32_t a = bt[i - ws], c = (i > ws) ? bt[i - ws - 1] : 0u;
        tp[i] = bs ? ((a << bs) | (c >> (32 - bs))) : a;
      }
    }
  }
  cudaError_t e = cudaMalloc(&b->d_blob, blob.size());
  // NOT cudaMemcpy: from pageable memory it may return once the data is staged, before the DMA has landed, and the
  // scoring kernels run on non-blocking streams that do not order against the legacy stream.
  if (e == cudaSuccess) e = upload_sync(b->d_blob, blob.data(), blob.size(), device);
  if (e != cudaSuccess) { rf_batch_destroy(b); return cuda_fail(e, "query table upload"); }
  b->view.len1 = b->len1;
  b->view.words = b->words;
  b->view.tab32_top = (const uint32_t*)b->d_blob;
  b->view.tab32_bot = b->view.tab32_top + 256;
  b->view.tab64_top = (const uint64_t*)(b->d_blob + 2 * sz32);
  b->view.tab64_bot = b->view.tab64_top + 256;
  b->view.pm_words = b->view.tab64_bot + 256;
  b->view.pm_band = (const uint32_t*)(b->view.pm_words + (size_t)256 * words);
  b->view.band_stride = bstride;
  b->view.quot = (const double*)(b->d_blob + 2 * sz32 + 2 * sz64 + szw + szp8);
  b->view.qbytes = b->d_blob + 2 * sz32 + 2 * sz64 + szw + szp8 + szq;
  b->view.limbs = limbs;
  b->view.pmn_top = limbs ? (const uint32_t*)(b->d_blob + off_n) : nullptr;
  b->view.pmn_bot = limbs ? b->view.pmn_top + (size_t)256 * limbs : nullptr;
  *out = b;
  return RF_OK;
}

extern "C" {

rf_status rf_batch_create_u8(rf_metric metric, const uint8_t* query, uint32_t query_len, int device, rf_batch** out) {
  return batch_create_bytes(metric, query, query_len, device, out);
}

rf_status rf_batch_create_u32(rf_metric metric, const uint32_t* query, uint32_t query_len, int device, rf_batch** out) {
  if (!out) return fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (query_len && !query) return fail(RF_ERR_INVALID_ARG, "query is NULL");
  if (query_len > RF_MAX_QUERY_LEN) return fail(RF_ERR_UNSUPPORTED, "query longer than RF_MAX_QUERY_LEN");
  std::vector<uint32_t> keys(kAlphaSlots, 0);
  std::vector<uint8_t> codes(kAlphaSlots, 0), renamed(query_len);
  uint32_t distinct = 0;
  bool overflow = false;
  for (uint32_t i = 0; i < query_len; ++i) {
    uint32_t slot = alpha_hash(query[i]);
    while (codes[slot] && keys[slot] != query[i]) slot = (slot + 1) & (kAlphaSlots - 1);
    if (!codes[slot]) {
      if (distinct == 255) { overflow = true; break; }
      keys[slot] = query[i];
      codes[slot] = (uint8_t)++distinct;
    }
    renamed[i] = codes[slot];
  }
  rf_batch* b = nullptr;
  rf_status s;
  if (overflow) {
    // more than 255 distinct symbols: the query cannot be renamed to bytes on its own.  It still scores byte corpora and
    // u32 corpora that were renamed to bytes at creation, where it is mapped into the CORPUS' symbol domain (byte_sub /
    // compact_sub: symbols the corpus cannot contain match nothing); its own tables stay empty.
    std::vector<uint8_t> none(query_len, 1);
    std::fill(renamed.begin(), renamed.end(), 0);
    s = batch_create_bytes(metric, renamed.data(), query_len, device, &b, none.data());
  } else {
    s = batch_create_bytes(metric, renamed.data(), query_len, device, &b);
  }
  if (s != RF_OK) return s;
  DeviceGuard g(device);
  b->wide = true;
  b->alpha_overflow = overflow;
  b->s1w.assign(query, query + query_len);
  for (uint32_t i = 0; i < query_len; ++i)
    if (query[i] >> 31) b->has_huge = true;
  cudaError_t e = cudaMalloc(&b->d_alpha_keys, kAlphaSlots * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&b->d_alpha_codes, kAlphaSlots);
  if (e == cudaSuccess) e = upload_sync(b->d_alpha_keys, keys.data(), kAlphaSlots * sizeof(uint32_t), device);
  if (e == cudaSuccess) e = upload_sync(b->d_alpha_codes, codes.data(), kAlphaSlots, device);
  if (e != cudaSuccess) { rf_batch_destroy(b); return cuda_fail(e, "alphabet upload"); }
  *out = b;
  return RF_OK;
}

// Any HashableChar element type (details/common.rs:29-37) -> the u32 domain this library scores in, BY VALUE: unsigned
// types zero-extend, signed types keep the two's-complement 32-bit pattern of the value.  *negative / *huge report whether
// a negative value / an unsigned value >= 2^31 was seen (their 32-bit patterns would collide: the reference never equates
// a negative signed with an unsigned element, Hash::SIGNED vs Hash::UNSIGNED).
static rf_status widen_to_u32(const void* src, rf_elem_type t, uint64_t count, std::vector<uint32_t>* out, bool* negative, bool* huge) {
  out->resize(count);
  uint32_t* d = out->data();
  int neg = 0, hug = 0, bad = 0;
  switch (t) {
    case RF_ELEM_U8: { const uint8_t* p = (const uint8_t*)src;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < (int64_t)count; ++i) d[i] = p[i]; break; }
    case RF_ELEM_U16: { const uint16_t* p = (const uint16_t*)src;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < (int64_t)count; ++i) d[i] = p[i]; break; }
    case RF_ELEM_U32: { const uint32_t* p = (const uint32_t*)src;
#pragma omp parallel for schedule(static) reduction(| : hug)
      for (int64_t i = 0; i < (int64_t)count; ++i) { d[i] = p[i]; hug |= p[i] >> 31; } break; }
    case RF_ELEM_U64: { const uint64_t* p = (const uint64_t*)src;
#pragma omp parallel for schedule(static) reduction(| : hug, bad)
      for (int64_t i = 0; i < (int64_t)count; ++i) { d[i] = (uint32_t)p[i]; hug |= (int)((p[i] >> 31) & 1); bad |= (p[i] >> 32) != 0; } break; }
    case RF_ELEM_I8: { const int8_t* p = (const int8_t*)src;
#pragma omp parallel for schedule(static) reduction(| : neg)
      for (int64_t i = 0; i < (int64_t)count; ++i) { d[i] = (uint32_t)(int32_t)p[i]; neg |= p[i] < 0; } break; }
    case RF_ELEM_I16: { const int16_t* p = (const int16_t*)src;
#pragma omp parallel for schedule(static) reduction(| : neg)
      for (int64_t i = 0; i < (int64_t)count; ++i) { d[i] = (uint32_t)(int32_t)p[i]; neg |= p[i] < 0; } break; }
    case RF_ELEM_I32: { const int32_t* p = (const int32_t*)src;
#pragma omp parallel for schedule(static) reduction(| : neg)
      for (int64_t i = 0; i < (int64_t)count; ++i) { d[i] = (uint32_t)p[i]; neg |= p[i] < 0; } break; }
    case RF_ELEM_I64: { const int64_t* p = (const int64_t*)src;
#pragma omp parallel for schedule(static) reduction(| : neg, bad)
      for (int64_t i = 0; i < (int64_t)count; ++i) {
        d[i] = (uint32_t)(int32_t)p[i]; neg |= p[i] < 0; bad |= (p[i] < -2147483648LL) | (p[i] > 4294967295LL); hug |= p[i] > 2147483647LL;
      } break; }
    default: return fail(RF_ERR_INVALID_ARG, "unknown element type");
  }
  if (bad) return fail(RF_ERR_UNSUPPORTED, "64-bit element values outside [-2^31, 2^32) are not supported");
  *negative = neg != 0;
  *huge = hug != 0;
  return RF_OK;
}

rf_status rf_corpus_create_elems(const void* elems, rf_elem_type type, const uint64_t* offsets, uint64_t n, int device, rf_corpus** out) {
  if (!out) return fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (!offsets) return fail(RF_ERR_INVALID_ARG, "offsets is NULL");
  if (type == RF_ELEM_U8) return rf_corpus_create_u8((const uint8_t*)elems, offsets, n, device, out);
  if (offsets[n] && !elems) return fail(RF_ERR_INVALID_ARG, "elems is NULL");
  std::vector<uint32_t> w;
  bool neg = false, huge = false;
  rf_status s = widen_to_u32(elems, type, offsets[n], &w, &neg, &huge);
  if (s != RF_OK) return s;
  s = rf_corpus_create_u32(w.data(), offsets, n, device, out);
  if (s == RF_OK) { (*out)->has_negative = neg; (*out)->has_huge = huge; }  // (the u32 path saw the negatives' patterns as "huge")
  return s;
}

rf_status rf_batch_create_elems(rf_metric metric, const void* query, rf_elem_type type, uint32_t query_len, int device, rf_batch** out) {
  if (!out) return fail(RF_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  if (query_len && !query) return fail(RF_ERR_INVALID_ARG, "query is NULL");
  std::vector<uint32_t> w;
  bool neg = false, huge = false;
  rf_status s = widen_to_u32(query, type, query_len, &w, &neg, &huge);
  if (s != RF_OK) return s;
  s = rf_batch_create_u32(metric, w.data(), query_len, device, out);
  if (s == RF_OK) { (*out)->has_negative = neg; (*out)->has_huge = huge; }
  return s;
}

rf_status rf_batch_set_option(rf_batch* b, const char* name, int value) {
  if (!b || !name) return fail(RF_ERR_INVALID_ARG, "NULL argument");
  if (!strcmp(name, "single_word_path")) b->opt.w1_path = value;
  else if (!strcmp(name, "multi_word_path")) b->opt.mw_path = value ? 1 : 0;
  else if (!strcmp(name, "banded_levenshtein")) b->opt.band = value ? 1 : 0;
  else if (!strcmp(name, "epilogue_table")) b->opt.epi_table = value ? 1 : 0;
  else if (!strcmp(name, "jaro32")) b->opt.jaro32 = value >= 0 && value <= 3 ? value : 1;
  else return fail(RF_ERR_INVALID_ARG, std::string("not a per-comparator option: ") + name);
  for (auto& kv : b->subs) rf_batch_set_option(kv.second, name, value);
  return RF_OK;
}

rf_status rf_batch_destroy(rf_batch* b) {
  if (!b) return RF_OK;
  DeviceGuard g(b->device);
  if (b->d_blob) cudaFree(b->d_blob);
  if (b->d_alpha_keys) cudaFree(b->d_alpha_keys);
  if (b->d_alpha_codes) cudaFree(b->d_alpha_codes);
  if (b->d_w16_keys) cudaFree(b->d_w16_keys);
  if (b->d_w16_codes) cudaFree(b->d_w16_codes);
  if (b->d_w16_pm) cudaFree(b->d_w16_pm);
  for (auto& kv : b->subs) rf_batch_destroy(kv.second);
  delete b;
  return RF_OK;
}

// ------------------------------------------------------------------------------------------------ scoring
static rf_status make_epi(const rf_batch* b, rf_kind kind, const rf_args* args, Epi* e) {
  rf_args def;
  rf_args_default(&def);
  const rf_args* a = args ? args : &def;
  memset(e, 0, sizeof(*e));
  e->metric = (int)b->metric;
  e->kind = (int)kind;
  e->has_cutoff = a->has_cutoff ? 1 : 0;
  e->cutoff_u = a->cutoff_u;
  e->cutoff_f = a->cutoff_f;
  e->w_ins = a->insertion_cost;
  e->w_del = a->deletion_cost;
  e->w_sub = a->substitution_cost;
  e->prefix_weight = a->prefix_weight;
  e->quirks = a->reference_quirks ? 1 : 0;
  e->pad = a->pad ? 1 : 0;
  e->wclass = WC_UNIFORM;
  if (b->metric == RF_LEVENSHTEIN) {  // weight classes of levenshtein.rs:1301-1330
    if (a->insertion_cost == a->deletion_cost) {
      if (a->insertion_cost == 0) e->wclass = WC_ZERO;
      else if (a->insertion_cost == a->substitution_cost) e->wclass = WC_UNIFORM;
      else if (a->substitution_cost >= a->insertion_cost + a->deletion_cost) e->wclass = WC_INDEL;
      else e->wclass = WC_GENERIC;  // generalized_distance (levenshtein.rs:1330): Wagner-Fischer
    } else {
      e->wclass = WC_GENERIC;
    }
  } else {
    e->w_ins = e->w_del = e->w_sub = 1;
  }
  if (b->metric == RF_RATIO) e->kind = K_NORM_SIMILARITY;  // fuzz.rs:127-149 has one method only
  e->unit32 = ((b->metric == RF_LEVENSHTEIN && e->wclass == WC_UNIFORM && e->w_ins == 1) || b->metric == RF_INDEL ||
               b->metric == RF_LCS_SEQ || b->metric == RF_OSA || b->metric == RF_HAMMING || b->metric == RF_PREFIX ||
               b->metric == RF_POSTFIX || b->metric == RF_DAMERAU_LEVENSHTEIN)
                  ? 1
                  : 0;
  return RF_OK;
}

// Scores the candidates described by `cv` (+ optional interleaved copy `lb`), all resident on `device`.
// [r0, r1): optional candidate sub-range (chunked launches of the overlapped scan + all-gather).  r0 must be a multiple of
// LB_BLOCK; the results still land at out_dev[candidate index].
static rf_status score_view(const rf_batch* b, const CorpusView& cv_in, const LbAlloc* lb, int device, rf_kind kind,
                            const rf_args* args, void* out_dev, bool want_f64, cudaStream_t st, uint32_t* d_err = nullptr,
                            uint64_t r0 = 0, uint64_t r1 = UINT64_MAX) {
  CorpusView cv = cv_in;
  if (r1 > cv.n) r1 = cv.n;
  const bool ranged = r0 != 0 || r1 != cv.n;
  if (ranged) {
    if (r0 % LB_BLOCK != 0 || r0 > r1) return fail(RF_ERR_INVALID_ARG, "candidate range must start on a multiple of 65536");
    if (r0 == r1) return RF_OK;
  }
  if (!b) return fail(RF_ERR_INVALID_ARG, "NULL handle");
  if ((int)kind < 0 || (int)kind > 3) return fail(RF_ERR_INVALID_ARG, "unknown kind");
  if (b->device != device) return fail(RF_ERR_INVALID_ARG, "batch and corpus live on different devices");
  if ((rf_result_is_float(b->metric, kind) != 0) != want_f64)
    return fail(RF_ERR_INVALID_ARG, want_f64 ? "this (metric, kind) yields u32 results; use the _u32 entry point"
                                             : "this (metric, kind) yields f64 results; use the _f64 entry point");
  if (cv.n == 0) return RF_OK;
  if (!out_dev) return fail(RF_ERR_INVALID_ARG, "out is NULL");
  ScanLaunch L{};
  rf_status s = make_epi(b, kind, args, &L.epi);
  if (s != RF_OK) return s;
  DeviceGuard g(device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  L.corpus = cv;
  const Family fam0 = family_of(L.epi.metric, L.epi.wclass);
  const rf_batch_opts opt = b->opt;  // per-comparator snapshot of the kernel-choice knobs
  // rf_corpus_release_csr: only what the interleaved layout serves is left (single_word_path = 1, the CSR kernel, is ignored)
  const bool have_csr = cv.off32 != nullptr || cv.off64 != nullptr;
  const bool use_lb = lb && lb->gdata && (opt.w1_path != 1 || !have_csr) && (fam0 != F_SIMPLE || L.epi.metric == M_HAMMING) &&
                      ((fam0 != F_DL && fam0 != F_WF) || b->len1 <= 64);  // the two DP kernels: shared-memory rows up to 64
  if (!have_csr) {
    const char* why = nullptr;
    if (!use_lb) why = "this metric / query length is served by the CSR copy, which rf_corpus_release_csr freed";
    else if ((fam0 == F_DL || fam0 == F_WF) && cv.max_len > 32000) why = "candidates beyond 32 000 elements take the CSR kernel, and rf_corpus_release_csr freed that copy";
    else if (b->len1 > 64 && (fam0 == F_JARO || !b->view.limbs)) why = "queries beyond 512 elements (Jaro: beyond 64) are served by the CSR copy, which rf_corpus_release_csr freed";
    if (why) return fail(RF_ERR_UNSUPPORTED, why);
  }
  if (ranged) {
    const bool dp = fam0 == F_DL || fam0 == F_WF || fam0 == F_SIMPLE;
    if (use_lb && dp) return fail(RF_ERR_UNSUPPORTED, "candidate sub-ranges are not available for this metric on the interleaved layout");
    if (!use_lb) {  // CSR kernels index candidates from 0: shift the view and the output
      if (cv.off32) cv.off32 += r0; else cv.off64 += r0;
      cv.n = r1 - r0;
      out_dev = (uint8_t*)out_dev + r0 * (want_f64 ? 8 : 4);
      L.corpus = cv;
    }
  }
  if (use_lb) {
    L.lb = LbView{lb->perm, lb->lens, lb->goff, lb->gdata, lb->ngroups};
    if (ranged) {  // whole 65536-candidate blocks = whole groups; results are scattered through perm (absolute indices)
      const uint64_t g0 = r0 / 32, g1 = (r1 + 31) / 32;
      L.lb.perm += g0 * 32;
      L.lb.lens += g0 * 32;
      L.lb.goff += g0;
      L.lb.ngroups = g1 - g0;
    }
    L.lb_counter = counter_slot(device);
    L.lb_flag = counter_slot(device);
    if (!L.lb_counter || !L.lb_flag) return fail(RF_ERR_OOM, "scheduler scratch allocation failed");
  }
  L.query = b->view;
  L.out = out_dev;
  L.out_is_f64 = want_f64 ? 1 : 0;
  L.stream = st;
  L.sm_count = sm_count_of(device);
  L.jaro32 = opt.jaro32;
  L.epi_table = opt.epi_table;
  const Family fam = family_of(L.epi.metric, L.epi.wclass);
  cudaError_t e;
  if (fam == F_SIMPLE) e = launch_simple(L, d_err);
  else if (fam == F_WF) {
    if (b->len1 > kDpMaxQuery) return fail(RF_ERR_UNSUPPORTED, "generic Levenshtein weights: queries longer than 200 000 elements");
    e = launch_wf(L);
  } else if (fam == F_DL) {
    if (b->len1 > kDpMaxQuery) return fail(RF_ERR_UNSUPPORTED, "Damerau-Levenshtein: queries longer than 200 000 elements");
    e = launch_dl(L);
  }
  else if (b->len1 <= 64) {
    const int path = opt.w1_path;
    e = !use_lb ? launch_scan_w1(L) : path == 2 ? launch_scan_lbr(L) : launch_scan_lb(L);
  }
  else if (fam == F_JARO) e = launch_jaro_mw(L);
  else if (use_lb && L.query.limbs && (!have_csr || (opt.mw_path == 0 &&
           !(opt.band && L.epi.metric == M_LEVENSHTEIN && L.epi.wclass == WC_UNIFORM && L.epi.kind == K_DISTANCE &&
             L.epi.has_cutoff && L.epi.cutoff_u / L.epi.w_ins <= 63))))
    e = launch_scan_lbn(L);  // 65..512 elements on a resident corpus: thread per candidate, column in registers
  else if (opt.band && L.epi.metric == M_LEVENSHTEIN && L.epi.wclass == WC_UNIFORM && L.epi.kind == K_DISTANCE &&
           L.epi.has_cutoff && L.epi.cutoff_u / L.epi.w_ins <= 63)
    e = launch_scan_band(L, (uint32_t)(L.epi.cutoff_u / L.epi.w_ins));  // small cutoff: 64-bit Ukkonen band per thread
  else if (L.query.words > 256) e = launch_scan_long(L);  // beyond 16 384 elements: column in stripes, carries through scratch
  else e = launch_scan_mw(L);
  if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
  return RF_OK;
}

}  // extern "C"

// candidates' symbols -> the query's byte alphabet (see rf_batch::wide); 4 elements per thread, table in shared memory
template <class In>
__global__ void __launch_bounds__(256) remap_kernel(const In* __restrict__ in, uint64_t total, const uint32_t* __restrict__ keys,
                                                    const uint8_t* __restrict__ codes, uint32_t* __restrict__ out4) {
  __shared__ uint32_t skeys[1024];
  __shared__ uint8_t scodes[1024];
  for (uint32_t i = threadIdx.x; i < 1024; i += 256) { skeys[i] = keys[i]; scodes[i] = codes[i]; }
  __syncthreads();
  const uint64_t nquads = (total + 3) / 4;
  for (uint64_t qd = (uint64_t)blockIdx.x * 256 + threadIdx.x; qd < nquads; qd += (uint64_t)gridDim.x * 256) {
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t i = qd * 4 + j;
      uint32_t code = 0;
      if (i < total) {
        const uint32_t x = (uint32_t)in[i];
        uint32_t slot = (x * 2654435761u) >> 22;
        while (scodes[slot] && skeys[slot] != x) slot = (slot + 1) & 1023u;
        code = scodes[slot];
      }
      packed |= code << (8 * j);
    }
    out4[qd] = packed;
  }
}

// distinct symbols of a u32 corpus: per-CTA hash set in shared memory, merged into a global one; more than 255 -> overflow
__global__ void __launch_bounds__(256) distinct_kernel(const uint32_t* __restrict__ in, uint64_t total,
                                                       unsigned long long* __restrict__ gset, uint32_t* __restrict__ gcount) {
  __shared__ unsigned long long sset[1024];
  __shared__ uint32_t scount, sover;
  for (uint32_t i = threadIdx.x; i < 1024; i += 256) sset[i] = 0ull;
  if (threadIdx.x == 0) { scount = 0; sover = 0; }
  __syncthreads();
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (uint64_t)gridDim.x * 256) {
    if (*reinterpret_cast<volatile uint32_t*>(&sover)) break;
    const uint32_t x = in[i];
    const unsigned long long key = (1ull << 32) | x;
    uint32_t slot = (x * 2654435761u) >> 22;
    for (;;) {
      const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&sset[slot]);
      if (cur == key) break;
      if (cur == 0ull) {
        const unsigned long long old = atomicCAS(&sset[slot], 0ull, key);
        if (old == 0ull) { if (atomicAdd(&scount, 1u) >= 255u) sover = 1; break; }
        if (old == key) break;
      }
      slot = (slot + 1) & 1023u;
    }
  }
  __syncthreads();
  if (sover) { if (threadIdx.x == 0) gcount[1] = 1; return; }
  for (uint32_t s0 = threadIdx.x; s0 < 1024; s0 += 256) {
    const unsigned long long key = sset[s0];
    if (!key) continue;
    uint32_t slot = ((uint32_t)key * 2654435761u) >> 22;
    for (;;) {
      const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&gset[slot]);
      if (cur == key) break;
      if (cur == 0ull) {
        const unsigned long long old = atomicCAS(&gset[slot], 0ull, key);
        if (old == 0ull) { if (atomicAdd(&gcount[0], 1u) >= 255u) gcount[1] = 1; break; }
        if (old == key) break;
      }
      if (*reinterpret_cast<volatile uint32_t*>(&gcount[1])) break;  // overflowing anyway: the set may fill up
      slot = (slot + 1) & 1023u;
    }
  }
}

static std::atomic<uint64_t> g_dict_serial{1};

// rf_corpus_create_u32, second half: if the whole corpus uses at most 255 distinct symbols, rename them to bytes once
// (codes 1..D in ascending symbol order), keep the corpus as u8 CSR + interleaved layout and drop the u32 copy.
static rf_status compact_u32_corpus(rf_corpus* c, cudaStream_t st) {
  unsigned long long* d_set = nullptr;
  uint32_t* d_cnt = nullptr;
  uint32_t* d_keys = nullptr;
  uint8_t* d_codes = nullptr;
  rf_status s = RF_OK;
  cudaError_t e;
  do {
    if ((e = dev_alloc(&d_set, 1024 * 8 + 16, st)) != cudaSuccess) break;
    d_cnt = reinterpret_cast<uint32_t*>(d_set + 1024);
    if ((e = cudaMemsetAsync(d_set, 0, 1024 * 8 + 16, st)) != cudaSuccess) break;
    const uint64_t blocks = (c->total + 256 * 64 - 1) / (256 * 64);
    const uint32_t grid = (uint32_t)(blocks < (uint64_t)sm_count_of(c->device) * 8 ? (blocks ? blocks : 1) : (uint64_t)sm_count_of(c->device) * 8);
    distinct_kernel<<<grid, 256, 0, st>>>(c->d_elems32, c->total, d_set, d_cnt);
    rfk::count_launches(1);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    std::vector<unsigned long long> hset(1024 + 2);
    if ((e = cudaMemcpyAsync(hset.data(), d_set, 1024 * 8 + 16, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    const uint32_t* cnt = reinterpret_cast<const uint32_t*>(hset.data() + 1024);
    if (cnt[1] || cnt[0] > 255) break;  // too many symbols: stays a u32 corpus (renamed per query)
    std::vector<uint32_t> syms;
    for (int i = 0; i < 1024; ++i)
      if (hset[i]) syms.push_back((uint32_t)hset[i]);
    std::sort(syms.begin(), syms.end());
    c->dict_keys.assign(kAlphaSlots, 0);
    c->dict_codes.assign(kAlphaSlots, 0);
    for (size_t k = 0; k < syms.size(); ++k) {
      uint32_t slot = alpha_hash(syms[k]);
      while (c->dict_codes[slot]) slot = (slot + 1) & (kAlphaSlots - 1);
      c->dict_keys[slot] = syms[k];
      c->dict_codes[slot] = (uint8_t)(k + 1);
    }
    if ((e = dev_alloc(&d_keys, kAlphaSlots * 4, st)) != cudaSuccess) break;
    if ((e = dev_alloc(&d_codes, kAlphaSlots, st)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(d_keys, c->dict_keys.data(), kAlphaSlots * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(d_codes, c->dict_codes.data(), kAlphaSlots, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    const uint64_t padded = (c->total + 3) / 4 * 4;
    if ((e = dev_alloc(&c->d_chars, padded + 64, st)) != cudaSuccess) break;
    if ((e = cudaMemsetAsync(c->d_chars + (padded - 4), 0, 64 + 4, st)) != cudaSuccess) break;
    {
      const uint64_t rb = ((c->total + 3) / 4 + 255) / 256;
      const uint32_t rgrid = (uint32_t)(rb < 148 * 16 ? rb : 148 * 16);
      remap_kernel<uint32_t><<<rgrid, 256, 0, st>>>(c->d_elems32, c->total, d_keys, d_codes, (uint32_t*)c->d_chars);
      rfk::count_launches(1);
      if ((e = cudaGetLastError()) != cudaSuccess) break;
    }
    dev_free(c->d_elems32, st);
    c->d_elems32 = nullptr;
    c->compact32 = true;
    c->dict_serial = g_dict_serial.fetch_add(1);
    s = corpus_finish(c, st);  // interleaved layout for the single-word kernels (the offsets' tail is already filled)
    if (s == RF_OK) e = cudaStreamSynchronize(st);
  } while (0);
  dev_free(d_set, st);
  dev_free(d_keys, st);
  dev_free(d_codes, st);
  cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "u32 corpus alphabet");
  return s;
}

// the byte comparator of a u32 query against one compact corpus (cached per dictionary)
static const rf_batch* compact_sub(const rf_batch* b, const rf_corpus* c) {
  std::lock_guard<std::mutex> lk(b->sub_mu);
  auto it = b->subs.find(c->dict_serial);
  if (it != b->subs.end()) return it->second;
  std::vector<uint8_t> renamed(b->s1w.size());
  for (size_t i = 0; i < b->s1w.size(); ++i) {
    const uint32_t x = b->s1w[i];
    uint32_t slot = alpha_hash(x);
    while (c->dict_codes[slot] && c->dict_keys[slot] != x) slot = (slot + 1) & (kAlphaSlots - 1);
    renamed[i] = c->dict_codes[slot];  // 0 when the corpus never contains x
  }
  rf_batch* sub = nullptr;
  if (batch_create_bytes(b->metric, renamed.data(), (uint32_t)renamed.size(), b->device, &sub) != RF_OK) return nullptr;
  sub->opt = b->opt;
  // Never evicted while the parent lives: a concurrent (or still enqueued, asynchronous) scoring call on the same
  // rf_batch may be using any cached entry.  One entry per distinct compact corpus ever scored (~100 KB of device
  // memory each); all are released by rf_batch_destroy.
  b->subs.emplace(c->dict_serial, sub);
  return sub;
}

// ---- 16-bit codes: a u32 query with more than 255 distinct symbols against a u32 corpus with more than 255 distinct symbols.
// The query's D distinct symbols become the codes 1..D (D <= 65535), the candidates are renamed per call (symbols the query
// does not contain -> 0, whose match row is empty) and scored by the multi-word kernels instantiated for uint16_t elements
// with one match-table row per code -- the role of the reference's per-block hashmap (pattern_match_vector.rs:20-65).
__global__ void __launch_bounds__(256) remap16_kernel(const uint32_t* __restrict__ in, uint64_t total, const uint32_t* __restrict__ keys,
                                                      const uint16_t* __restrict__ codes, uint32_t slot_mask, uint32_t shift,
                                                      uint16_t* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (uint64_t)gridDim.x * 256) {
    const uint32_t x = in[i];
    uint32_t slot = (x * 2654435761u) >> shift;
    uint32_t code;
    while ((code = __ldg(codes + slot)) != 0 && __ldg(keys + slot) != x) slot = (slot + 1) & slot_mask;
    out[i] = (uint16_t)code;
  }
}

static rf_status ensure_w16(const rf_batch* b) {
  std::lock_guard<std::mutex> lk(b->sub_mu);
  if (b->w16_ready) return RF_OK;
  const uint32_t len1 = (uint32_t)b->s1w.size(), words = (len1 + 63) / 64;
  uint32_t slots = 1024, bits = 10;
  while (slots < 4 * (uint64_t)len1 && slots < (1u << 20)) { slots <<= 1; ++bits; }
  std::vector<uint32_t> keys(slots, 0);
  std::vector<uint16_t> codes(slots, 0), code_of(len1);
  uint32_t distinct = 0;
  for (uint32_t i = 0; i < len1; ++i) {
    const uint32_t x = b->s1w[i];
    uint32_t slot = (x * 2654435761u) >> (32 - bits);
    while (codes[slot] && keys[slot] != x) slot = (slot + 1) & (slots - 1);
    if (!codes[slot]) {
      if (distinct == 65535) return fail(RF_ERR_UNSUPPORTED, "u32 query with more than 65535 distinct symbols against a u32 corpus with more than 255");
      keys[slot] = x;
      codes[slot] = (uint16_t)++distinct;
    }
    code_of[i] = codes[slot];
  }
  const uint64_t pm_bytes = (uint64_t)(distinct + 1) * words * 8;
  if (pm_bytes > (4ull << 30)) return fail(RF_ERR_UNSUPPORTED, "match table of this query (distinct symbols x length / 8 bytes) exceeds 4 GiB");
  std::vector<uint64_t> pm((size_t)(distinct + 1) * words, 0);
  for (uint32_t i = 0; i < len1; ++i) pm[(size_t)code_of[i] * words + i / 64] |= 1ull << (i % 64);
  DeviceGuard g(b->device);
  cudaError_t e = cudaMalloc(&b->d_w16_keys, (size_t)slots * 4);
  if (e == cudaSuccess) e = cudaMalloc(&b->d_w16_codes, (size_t)slots * 2);
  if (e == cudaSuccess) e = cudaMalloc(&b->d_w16_pm, pm_bytes);
  if (e == cudaSuccess) e = upload_sync(b->d_w16_keys, keys.data(), (size_t)slots * 4, b->device);
  if (e == cudaSuccess) e = upload_sync(b->d_w16_codes, codes.data(), (size_t)slots * 2, b->device);
  if (e == cudaSuccess) e = upload_sync(b->d_w16_pm, pm.data(), pm_bytes, b->device);
  if (e != cudaSuccess) return cuda_fail(e, "16-bit alphabet tables");
  b->w16_slots = slots;
  b->w16_ready = true;
  return RF_OK;
}

// the byte comparator of a u32 query against BYTE candidates (cached under serial 0): a symbol below 256 is its own byte,
// anything else can equal no candidate element and sets no table bit.  Exact for the table-driven metrics; no pass over
// the candidates, and the interleaved-layout kernels apply.
static const rf_batch* byte_sub(const rf_batch* b) {
  std::lock_guard<std::mutex> lk(b->sub_mu);
  auto it = b->subs.find(0);
  if (it != b->subs.end()) return it->second;
  std::vector<uint8_t> bytes(b->s1w.size()), none(b->s1w.size());
  for (size_t i = 0; i < b->s1w.size(); ++i) {
    bytes[i] = (uint8_t)b->s1w[i];
    none[i] = b->s1w[i] > 255 ? 1 : 0;
  }
  rf_batch* sub = nullptr;
  if (batch_create_bytes(b->metric, bytes.data(), (uint32_t)bytes.size(), b->device, &sub, none.data()) != RF_OK) return nullptr;
  sub->opt = b->opt;
  b->subs.emplace(0, sub);
  return sub;
}
static bool table_driven(rf_metric m, const rf_args* a) {
  switch (m) {
    case RF_LEVENSHTEIN:
      if (!a) return true;
      return a->insertion_cost == a->deletion_cost &&
             (a->insertion_cost == a->substitution_cost || a->substitution_cost >= a->insertion_cost + a->deletion_cost);
    case RF_INDEL: case RF_LCS_SEQ: case RF_OSA: case RF_JARO: case RF_JARO_WINKLER: case RF_RATIO: return true;
    default: return false;  // hamming / prefix / postfix / Damerau-Levenshtein / generic weights compare symbols directly
  }
}

extern "C" {

static rf_status score_device(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, void* out_dev,
                              bool want_f64, cudaStream_t st, uint32_t* d_err = nullptr) {
  if (!b || !c) return fail(RF_ERR_INVALID_ARG, "NULL handle");
  if ((c->d_elems32 || c->compact32) && !b->wide)
    return fail(RF_ERR_INVALID_ARG, "a u32 corpus needs a comparator created with rf_batch_create_u32");
  if ((b->has_negative && c->has_huge) || (b->has_huge && c->has_negative))
    return fail(RF_ERR_UNSUPPORTED, "negative signed elements on one side and unsigned elements >= 2^31 on the other share 32-bit "
                                    "patterns; the reference never equates them (Hash::SIGNED vs Hash::UNSIGNED)");
  if (c->compact32) {  // the corpus already is bytes in ITS alphabet: score with the query renamed through its dictionary
    if (b->device != c->device) return fail(RF_ERR_INVALID_ARG, "batch and corpus live on different devices");
    const rf_batch* sub = compact_sub(b, c);
    if (!sub) return RF_ERR_CUDA;  // message set by the failing call
    return score_view(sub, CorpusView{c->d_chars, c->d_off32, c->d_off64, c->n, c->total, c->max_len}, &c->lb, c->device, kind, args,
                      out_dev, want_f64, st, d_err);
  }
  if (b->wide && !c->d_elems32 && table_driven(b->metric, args)) {  // u32 query, byte corpus: map the QUERY into the byte domain
    if (b->device != c->device) return fail(RF_ERR_INVALID_ARG, "batch and corpus live on different devices");
    const rf_batch* sub = byte_sub(b);
    if (!sub) return RF_ERR_CUDA;
    return score_view(sub, CorpusView{c->d_chars, c->d_off32, c->d_off64, c->n, c->total, c->max_len}, &c->lb, c->device, kind, args,
                      out_dev, want_f64, st, d_err);
  }
  if (b->wide && b->alpha_overflow && c->d_elems32 && table_driven(b->metric, args)) {
    // both sides have large alphabets: candidates renamed to the query's 16-bit codes, multi-word kernels over uint16_t
    if (b->device != c->device) return fail(RF_ERR_INVALID_ARG, "batch and corpus live on different devices");
    if ((int)kind < 0 || (int)kind > 3) return fail(RF_ERR_INVALID_ARG, "unknown kind");
    if ((rf_result_is_float(b->metric, kind) != 0) != want_f64)
      return fail(RF_ERR_INVALID_ARG, want_f64 ? "this (metric, kind) yields u32 results; use the _u32 entry point"
                                               : "this (metric, kind) yields f64 results; use the _f64 entry point");
    if (c->n == 0) return RF_OK;
    if (!out_dev) return fail(RF_ERR_INVALID_ARG, "out is NULL");
    rf_status s = ensure_w16(b);
    if (s != RF_OK) return s;
    DeviceGuard g(c->device);
    if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
    uint16_t* d_codes = nullptr;
    cudaError_t e = dev_alloc(&d_codes, (c->total + 64) * 2, st);
    if (e != cudaSuccess) return cuda_fail(e, "renamed candidates");
    e = cudaMemsetAsync(d_codes + c->total, 0, 128, st);
    if (e == cudaSuccess && c->total) {
      const uint64_t blocks = (c->total + 255) / 256;
      uint32_t bits = 0;
      while ((1u << bits) < b->w16_slots) ++bits;
      remap16_kernel<<<(uint32_t)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(c->d_elems32, c->total, b->d_w16_keys, b->d_w16_codes,
                                                                                        b->w16_slots - 1, 32 - bits, d_codes);
      rfk::count_launches(1);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
      ScanLaunch L{};
      s = make_epi(b, kind, args, &L.epi);
      L.corpus = CorpusView{reinterpret_cast<const uint8_t*>(d_codes), c->d_off32, c->d_off64, c->n, c->total, c->max_len};
      L.query = b->view;
      L.query.pm_words = b->d_w16_pm;
      L.out = out_dev;
      L.out_is_f64 = want_f64 ? 1 : 0;
      L.stream = st;
      L.sm_count = sm_count_of(c->device);
      L.elem16 = 1;
      if (s == RF_OK) {
        const Family fam = family_of(L.epi.metric, L.epi.wclass);
        e = fam == F_JARO ? launch_jaro_mw(L) : (L.query.words > 256 ? launch_scan_long(L) : launch_scan_mw(L));
        if (e != cudaSuccess) s = cuda_fail(e, "kernel launch");
      }
    } else {
      s = cuda_fail(e, "alphabet renaming");
    }
    dev_free(d_codes, st);
    return s;
  }
  if (b->wide) {
    if (b->alpha_overflow)
      return fail(RF_ERR_UNSUPPORTED, "this metric compares symbols directly: u32 queries with more than 255 distinct symbols are not supported for it");
    if (c->csr_released) return fail(RF_ERR_UNSUPPORTED, "renaming the candidates for this u32 query needs the CSR copy, which rf_corpus_release_csr freed");
    if (b->device != c->device) return fail(RF_ERR_INVALID_ARG, "batch and corpus live on different devices");
    if (c->n == 0) return RF_OK;
    DeviceGuard g(c->device);
    if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
    uint8_t* d_bytes = nullptr;
    const uint64_t padded = (c->total + 3) / 4 * 4;
    cudaError_t e = dev_alloc(&d_bytes, padded + 64, st);
    if (e != cudaSuccess) return cuda_fail(e, "renamed candidates");
    e = cudaMemsetAsync(d_bytes + padded, 0, 64, st);
    if (e == cudaSuccess && c->total) {
      const uint64_t blocks = ((c->total + 3) / 4 + 255) / 256;
      const uint32_t grid = (uint32_t)(blocks < 148 * 16 ? blocks : 148 * 16);
      if (c->d_elems32) remap_kernel<uint32_t><<<grid, 256, 0, st>>>(c->d_elems32, c->total, b->d_alpha_keys, b->d_alpha_codes, (uint32_t*)d_bytes);
      else remap_kernel<uint8_t><<<grid, 256, 0, st>>>(c->d_chars, c->total, b->d_alpha_keys, b->d_alpha_codes, (uint32_t*)d_bytes);
      e = cudaGetLastError();
      rfk::count_launches(1);
    }
    rf_status s = (e == cudaSuccess) ? score_view(b, CorpusView{d_bytes, c->d_off32, c->d_off64, c->n, c->total, c->max_len}, nullptr, c->device,
                                                  kind, args, out_dev, want_f64, st, d_err)
                                     : cuda_fail(e, "alphabet renaming");
    dev_free(d_bytes, st);
    return s;
  }
  return score_view(b, CorpusView{c->d_chars, c->d_off32, c->d_off64, c->n, c->total, c->max_len}, &c->lb, c->device, kind, args,
                    out_dev, want_f64, st, d_err);
}

static rf_status score_host(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, void* out_host,
                            bool want_f64) {
  if (!b || !c) return fail(RF_ERR_INVALID_ARG, "NULL handle");
  if (c->n == 0) return score_device(b, c, kind, args, nullptr, want_f64, nullptr);
  if (!out_host) return fail(RF_ERR_INVALID_ARG, "out is NULL");
  DeviceGuard g(c->device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  const size_t bytes = (size_t)c->n * (want_f64 ? 8 : 4);
  uint8_t* d_out = nullptr;  // results, then 4 bytes of error flag (hamming without pad)
  cudaStream_t st;
  cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
  e = dev_alloc(&d_out, bytes + 16, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_out + bytes, 0, 16, st);
  if (e != cudaSuccess) { dev_free(d_out, st); cudaStreamSynchronize(st); cudaStreamDestroy(st); return cuda_fail(e, "result buffer"); }
  rf_status s = score_device(b, c, kind, args, d_out, want_f64, st, (uint32_t*)(d_out + bytes));
  if (s == RF_OK) {
    uint32_t differing = 0;
    e = cudaMemcpyAsync(out_host, d_out, bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&differing, d_out + bytes, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) s = cuda_fail(e, "result download");
    else if (differing) s = fail(RF_ERR_INVALID_ARG, "Differing length arguments provided");  // hamming::Error (hamming.rs:121-136)
  }
  dev_free(d_out, st);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  return s;
}

rf_status rf_batch_score_u32(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint32_t* out) {
  return score_host(b, c, kind, args, out, false);
}
rf_status rf_batch_score_f64(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, double* out) {
  return score_host(b, c, kind, args, out, true);
}
rf_status rf_batch_score_u32_device(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args,
                                    uint32_t* out, void* stream) {
  return score_device(b, c, kind, args, out, false, (cudaStream_t)stream);
}
rf_status rf_batch_score_f64_device(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args,
                                    double* out, void* stream) {
  return score_device(b, c, kind, args, out, true, (cudaStream_t)stream);
}
rf_status rf_batch_distance_u32(const rf_batch* b, const rf_corpus* c, const rf_args* a, uint32_t* out) {
  return score_host(b, c, RF_DISTANCE, a, out, false);
}
rf_status rf_batch_similarity_u32(const rf_batch* b, const rf_corpus* c, const rf_args* a, uint32_t* out) {
  return score_host(b, c, RF_SIMILARITY, a, out, false);
}
rf_status rf_batch_distance_f64(const rf_batch* b, const rf_corpus* c, const rf_args* a, double* out) {
  return score_host(b, c, RF_DISTANCE, a, out, true);
}
rf_status rf_batch_similarity_f64(const rf_batch* b, const rf_corpus* c, const rf_args* a, double* out) {
  return score_host(b, c, RF_SIMILARITY, a, out, true);
}
rf_status rf_batch_normalized_distance_f64(const rf_batch* b, const rf_corpus* c, const rf_args* a, double* out) {
  return score_host(b, c, RF_NORMALIZED_DISTANCE, a, out, true);
}
rf_status rf_batch_normalized_similarity_f64(const rf_batch* b, const rf_corpus* c, const rf_args* a, double* out) {
  return score_host(b, c, RF_NORMALIZED_SIMILARITY, a, out, true);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ extract / filter
// Scoring + on-device post-processing: only k (or the number of hits) index/score pairs cross PCIe.
namespace {
rf_status select_host(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, bool want_f64, bool filter,
                      uint32_t k, uint64_t cap, uint32_t* idx_out, void* score_out, uint32_t* n32_out, uint64_t* n64_out) {
  if (!b || !c) return fail(RF_ERR_INVALID_ARG, "NULL handle");
  if ((int)kind < 0 || (int)kind > 3) return fail(RF_ERR_INVALID_ARG, "unknown kind");
  if ((rf_result_is_float(b->metric, kind) != 0) != want_f64)
    return fail(RF_ERR_INVALID_ARG, want_f64 ? "this (metric, kind) yields u32 results; use the _u32 entry point"
                                             : "this (metric, kind) yields f64 results; use the _f64 entry point");
  if (!filter && (k == 0 || k > 1024)) return fail(RF_ERR_INVALID_ARG, "k must be in 1..1024");
  if (filter ? !n64_out : !n32_out) return fail(RF_ERR_INVALID_ARG, "count output is NULL");
  const uint64_t slots = filter ? cap : (uint64_t)k;
  if (slots && (!idx_out || !score_out)) return fail(RF_ERR_INVALID_ARG, "output buffer is NULL");
  if (filter) *n64_out = 0; else *n32_out = 0;
  if (c->n == 0) return RF_OK;
  DeviceGuard g(c->device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  const size_t esz = want_f64 ? 8 : 4;
  cudaStream_t st;
  cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
  void* d_scores = nullptr;
  uint8_t* d_out = nullptr;  // [idx slots*4][score slots*esz][count 8]
  rf_status s = RF_OK;
  do {
    if ((e = dev_alloc(&d_scores, (size_t)c->n * esz, st)) != cudaSuccess) break;
    const size_t off_score = (slots * 4 + 7) & ~(size_t)7, off_n = off_score + slots * esz;
    if ((e = dev_alloc(&d_out, off_n + 16, st)) != cudaSuccess) break;
    if ((e = cudaMemsetAsync(d_out + off_n, 0, 16, st)) != cudaSuccess) break;
    s = score_device(b, c, kind, args, d_scores, want_f64, st);
    if (s != RF_OK) break;
    rfk::SelectLaunch L{};
    L.scores = d_scores;
    L.n = c->n;
    L.f64 = want_f64 ? 1 : 0;
    const rf_kind ek = (b->metric == RF_RATIO) ? RF_NORMALIZED_SIMILARITY : kind;
    L.desc = (ek == RF_SIMILARITY || ek == RF_NORMALIZED_SIMILARITY) ? 1 : 0;
    L.filter = filter ? 1 : 0;
    L.k = k;
    L.cap = cap;
    L.out_idx = (uint32_t*)d_out;
    L.out_score = d_out + off_score;
    L.out_n32 = (uint32_t*)(d_out + off_n);
    L.out_n64 = (unsigned long long*)(d_out + off_n);
    L.stream = st;
    L.sm_count = sm_count_of(c->device);
    if ((e = rfk::launch_select(L)) != cudaSuccess) break;
    uint64_t cnt = 0;
    if ((e = cudaMemcpyAsync(&cnt, d_out + off_n, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    uint64_t m;
    if (filter) { *n64_out = cnt; m = cnt < cap ? cnt : cap; }
    else { *n32_out = (uint32_t)cnt; m = (uint32_t)cnt; }
    if (m) {
      if ((e = cudaMemcpyAsync(idx_out, d_out, m * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync(score_out, d_out + off_score, m * esz, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
      if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    }
  } while (0);
  if (e != cudaSuccess && s == RF_OK) s = cuda_fail(e, "extract/filter");
  dev_free(d_scores, st);
  dev_free(d_out, st);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  return s;
}
}  // namespace

extern "C" {
rf_status rf_batch_extract_u32(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint32_t k,
                               uint32_t* idx_out, uint32_t* score_out, uint32_t* n_out) {
  return select_host(b, c, kind, args, false, false, k, 0, idx_out, score_out, n_out, nullptr);
}
rf_status rf_batch_extract_f64(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint32_t k,
                               uint32_t* idx_out, double* score_out, uint32_t* n_out) {
  return select_host(b, c, kind, args, true, false, k, 0, idx_out, score_out, n_out, nullptr);
}
rf_status rf_batch_filter_u32(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint64_t capacity,
                              uint32_t* idx_out, uint32_t* score_out, uint64_t* n_hits) {
  return select_host(b, c, kind, args, false, true, 0, capacity, idx_out, score_out, nullptr, n_hits);
}
rf_status rf_batch_filter_f64(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint64_t capacity,
                              uint32_t* idx_out, double* score_out, uint64_t* n_hits) {
  return select_host(b, c, kind, args, true, true, 0, capacity, idx_out, score_out, nullptr, n_hits);
}
}  // extern "C"

// ------------------------------------------------------------------------------------------------ streaming
// One-shot scoring of HOST-resident candidates: the literal shape of the reference's hot loop
// (`for c in candidates { scorer.distance(c) }`, levenshtein.rs:1740-1777) when the candidates are not kept on
// the GPU.  The CSR corpus is cut into chunks; every chunk goes H2D -> scan (CSR kernels, no layout build) ->
// D2H on one of kSlots streams, so the PCIe copies of one chunk overlap the scan and the result download of
// the others.  Steady state is bound by the H2D link (about len+4 bytes per candidate).
namespace {
constexpr int kSlots = 4;
struct StreamSlot {
  cudaStream_t st = nullptr;
  cudaEvent_t done = nullptr;    // recorded behind the slot's last chunk (back-pressure of the _len8 entry points)
  bool busy = false;
  uint8_t* d_chars = nullptr;
  uint8_t* d_renamed = nullptr;  // u32-query comparators: the chunk renamed to the query's byte alphabet (lazily allocated)
  uint8_t* d_packed = nullptr;   // *_packed6 entry points: the chunk's 6-bit packed characters as they crossed the link
  uint8_t* d_lens = nullptr;     // *_len8 entry points: the chunk's u8 lengths, the scan's temporary storage, the narrowed results
  void* d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  uint8_t* d_out8 = nullptr;
  void* d_offs = nullptr;
  void* d_out = nullptr;
};
struct StreamCtx {
  std::mutex mu;  // one streaming call per device at a time
  bool ready = false;
  uint64_t cap_bytes = 0, cap_n = 0;
  StreamSlot slot[kSlots];
};

StreamCtx* stream_ctx(int device) {
  static std::mutex mu;
  static std::vector<StreamCtx*> all;
  std::lock_guard<std::mutex> lk(mu);
  if ((int)all.size() <= device) all.resize(device + 1, nullptr);
  if (!all[device]) all[device] = new StreamCtx();
  return all[device];
}

void stream_ctx_release(StreamCtx* x) {
  for (auto& s : x->slot) {
    if (s.d_chars) cudaFree(s.d_chars);
    if (s.d_renamed) cudaFree(s.d_renamed);
    if (s.d_packed) cudaFree(s.d_packed);
    if (s.d_lens) cudaFree(s.d_lens);
    if (s.d_scan_tmp) cudaFree(s.d_scan_tmp);
    if (s.d_out8) cudaFree(s.d_out8);
    if (s.d_offs) cudaFree(s.d_offs);
    if (s.d_out) cudaFree(s.d_out);
    if (s.done) cudaEventDestroy(s.done);
    if (s.st) cudaStreamDestroy(s.st);
    s = StreamSlot{};
  }
  x->ready = false;
}

cudaError_t stream_ctx_prepare(StreamCtx* x, uint64_t cap_bytes, uint64_t cap_n) {
  if (x->ready && x->cap_bytes == cap_bytes && x->cap_n == cap_n) return cudaSuccess;
  stream_ctx_release(x);
  cudaError_t e = cudaSuccess;
  for (auto& s : x->slot) {
    if ((e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking)) != cudaSuccess) break;
    if ((e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming | cudaEventBlockingSync)) != cudaSuccess) break;
    // chars: 16 bytes of alignment lead-in + 64 bytes of over-read slack; offsets: 16 entries of slack
    if ((e = cudaMalloc(&s.d_chars, cap_bytes + 256)) != cudaSuccess) break;
    if ((e = cudaMemset(s.d_chars, 0, cap_bytes + 256)) != cudaSuccess) break;
    if ((e = cudaMalloc(&s.d_offs, (cap_n + 1 + 16) * 8)) != cudaSuccess) break;
    if ((e = cudaMemset(s.d_offs, 0, (cap_n + 1 + 16) * 8)) != cudaSuccess) break;
    if ((e = cudaMalloc(&s.d_out, cap_n * 8)) != cudaSuccess) break;
  }
  if (e != cudaSuccess) { stream_ctx_release(x); return e; }
  x->cap_bytes = cap_bytes;
  x->cap_n = cap_n;
  x->ready = true;
  return cudaSuccess;
}

template <class OffT>
rf_status stream_impl(const rf_batch* b, const uint8_t* chars, const OffT* offsets, uint64_t n, rf_kind kind,
                      const rf_args* args, void* out_host, bool want_f64, bool sub_range = false) {
  if (!b) return fail(RF_ERR_INVALID_ARG, "NULL handle");
  if ((int)kind < 0 || (int)kind > 3) return fail(RF_ERR_INVALID_ARG, "unknown kind");
  if ((rf_result_is_float(b->metric, kind) != 0) != want_f64)
    return fail(RF_ERR_INVALID_ARG, want_f64 ? "this (metric, kind) yields u32 results; use the _u32 entry point"
                                             : "this (metric, kind) yields f64 results; use the _f64 entry point");
  if (n == 0) return RF_OK;
  if (!offsets || !out_host) return fail(RF_ERR_INVALID_ARG, "NULL argument");
  if (offsets[0] != 0 && !sub_range) return fail(RF_ERR_INVALID_ARG, "offsets[0] must be 0");
  if (offsets[n] && !chars) return fail(RF_ERR_INVALID_ARG, "chars is NULL");
  if (rf_device_count() <= b->device) return fail(RF_ERR_CUDA, "no such CUDA device");
  // u32-query comparator on byte candidates: the table-driven metrics score with the query mapped into the byte domain
  // (no renaming pass over the chunk); the others rename the chunk through the query's own alphabet
  const rf_batch* sb = b;
  if (b->wide && table_driven(b->metric, args)) {
    sb = byte_sub(b);
    if (!sb) return RF_ERR_CUDA;
  } else if (b->wide && b->alpha_overflow) {
    return fail(RF_ERR_UNSUPPORTED, "this metric compares symbols directly: u32 queries with more than 255 distinct symbols are not supported for it");
  }

  DeviceGuard g(b->device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  StreamCtx* x = stream_ctx(b->device);
  std::lock_guard<std::mutex> lk(x->mu);
  const uint64_t cap_bytes = (uint64_t)(g_stream_mb.load() > 0 ? g_stream_mb.load() : 1) << 20;
  const uint64_t cap_n = (uint64_t)(g_stream_kcand.load() > 0 ? g_stream_kcand.load() : 1) << 10;
  cudaError_t e = stream_ctx_prepare(x, cap_bytes, cap_n);
  if (e != cudaSuccess) return cuda_fail(e, "streaming buffers");
  const size_t osz = sizeof(OffT), rsz = want_f64 ? 8 : 4;
  rf_status s = RF_OK;
  uint64_t i0 = 0;
  int k = 0;
  while (i0 < n && s == RF_OK) {
    // largest i1 <= i0 + cap_n whose bytes (from the 16-byte aligned start) fit the slot
    const uint64_t B0 = (uint64_t)offsets[i0] & ~15ull;
    uint64_t hi = (n - i0 < cap_n) ? n : i0 + cap_n;
    if ((uint64_t)offsets[hi] - B0 > cap_bytes) {
      uint64_t lo = i0;  // invariant: offsets[lo] - B0 <= cap_bytes < offsets[hi] - B0
      while (hi - lo > 1) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if ((uint64_t)offsets[mid] - B0 <= cap_bytes) lo = mid; else hi = mid;
      }
      hi = lo;
      if (hi == i0) { s = fail(RF_ERR_UNSUPPORTED, "a single candidate exceeds the streaming chunk size (raise stream_chunk_mb)"); break; }
    }
    const uint64_t i1 = hi, cn = i1 - i0;
    const uint64_t B1 = (uint64_t)offsets[i1];
    StreamSlot& sl = x->slot[k];
    k = (k + 1) % kSlots;
    if (B1 > B0) e = cudaMemcpyAsync(sl.d_chars, chars + B0, B1 - B0, cudaMemcpyHostToDevice, sl.st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sl.d_offs, offsets + i0, (cn + 1) * osz, cudaMemcpyHostToDevice, sl.st);
    if (e != cudaSuccess) { s = cuda_fail(e, "chunk upload"); break; }
    const uint8_t* d_src = sl.d_chars;
    if (b->wide && B1 > B0 && sb == b) {
      // comparator made by rf_batch_create_u32 and a metric that compares symbols directly: its tables are over renamed
      // bytes 1..D, so the candidates' bytes go through the same renaming first
      if (!sl.d_renamed) {
        if ((e = cudaMalloc(&sl.d_renamed, cap_bytes + 256)) == cudaSuccess) e = cudaMemsetAsync(sl.d_renamed, 0, cap_bytes + 256, sl.st);
        if (e != cudaSuccess) { s = cuda_fail(e, "streaming buffers"); break; }
      }
      const uint64_t cb = B1 - B0, blocks = ((cb + 3) / 4 + 255) / 256;
      const uint32_t grid = (uint32_t)(blocks < 148 * 16 ? blocks : 148 * 16);
      remap_kernel<uint8_t><<<grid, 256, 0, sl.st>>>(sl.d_chars, cb, b->d_alpha_keys, b->d_alpha_codes, (uint32_t*)sl.d_renamed);
      rfk::count_launches(1);
      if ((e = cudaGetLastError()) != cudaSuccess) { s = cuda_fail(e, "alphabet renaming"); break; }
      d_src = sl.d_renamed;
    }
    // the kernels index chars with the caller's absolute offsets: hand them the slot shifted back by B0
    CorpusView cv{d_src - B0, osz == 4 ? (const uint32_t*)sl.d_offs : nullptr,
                  osz == 8 ? (const uint64_t*)sl.d_offs : nullptr, cn, B1 - (uint64_t)offsets[i0]};
    s = score_view(sb, cv, nullptr, b->device, kind, args, sl.d_out, want_f64, sl.st);
    if (s != RF_OK) break;
    e = cudaMemcpyAsync((uint8_t*)out_host + i0 * rsz, sl.d_out, cn * rsz, cudaMemcpyDeviceToHost, sl.st);
    if (e != cudaSuccess) { s = cuda_fail(e, "chunk download"); break; }
    i0 = i1;
  }
  for (auto& sl : x->slot) {
    e = cudaStreamSynchronize(sl.st);
    if (e != cudaSuccess && s == RF_OK) s = cuda_fail(e, "streaming scan");
  }
  return s;
}

// sum of n length bytes on the host: psadbw adds 16 bytes per instruction (the chunk planner of the _len8 entry points runs
// ahead of the DMA only if this is much faster than the link: 2 M lengths per 64 MB chunk)
static inline uint64_t sum_bytes(const uint8_t* p, uint64_t n) {
  uint64_t sum = 0, i = 0;
#if defined(__SSE2__)
  __m128i acc = _mm_setzero_si128();
  const __m128i zero = _mm_setzero_si128();
  for (; i + 16 <= n; i += 16) acc = _mm_add_epi64(acc, _mm_sad_epu8(_mm_loadu_si128((const __m128i*)(p + i)), zero));
  sum = (uint64_t)_mm_cvtsi128_si64(acc) + (uint64_t)_mm_cvtsi128_si64(_mm_unpackhi_epi64(acc, acc));
#endif
  for (; i < n; ++i) sum += p[i];
  return sum;
}

// u32 results of a chunk -> bytes (None 0xFFFFFFFF -> 0xFF); any other value above 254 raises the overflow flag
__global__ void __launch_bounds__(256) narrow_results_u8(const uint32_t* __restrict__ in, uint64_t n, uint8_t* __restrict__ out,
                                                         uint32_t* __restrict__ overflow) {
  const uint64_t nq = (n + 3) / 4;
  for (uint64_t qd = (uint64_t)blockIdx.x * 256 + threadIdx.x; qd < nq; qd += (uint64_t)gridDim.x * 256) {
    uint32_t packed = 0;
    bool over = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t i = qd * 4 + j;
      uint32_t v = i < n ? in[i] : 0u;
      if (v == 0xFFFFFFFFu) v = 0xFFu;
      else if (v > 254u) { over = true; v = 0xFFu; }
      packed |= v << (8 * j);
    }
    // whole, 4-byte aligned quads as one word; the last partial quad (and a misaligned caller buffer: rf_batch_score_u8_device)
    // byte by byte, so nothing is written past out[n - 1]
    if (qd * 4 + 4 <= n && (reinterpret_cast<uintptr_t>(out) & 3u) == 0) {
      reinterpret_cast<uint32_t*>(out)[qd] = packed;
    } else {
      for (int j = 0; j < 4; ++j)
        if (qd * 4 + j < n) out[qd * 4 + j] = (uint8_t)(packed >> (8 * j));
    }
    if (over) atomicOr(overflow, 1u);
  }
}

// 6-bit packed characters (rf_pack6_u8: 4 characters in 3 bytes) -> bytes through the 64-entry dictionary; 16 characters per
// thread: three aligned 32-bit loads, one 16-byte store
struct Dict64 { uint8_t b[64]; };
__global__ void __launch_bounds__(256) unpack6_kernel(const uint32_t* __restrict__ in, uint64_t nchars16, const Dict64 dict,
                                                      uint4* __restrict__ out) {
  __shared__ uint8_t sd[64];
  if (threadIdx.x < 64) sd[threadIdx.x] = dict.b[threadIdx.x];
  __syncthreads();
  for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < nchars16; t += (uint64_t)gridDim.x * 256) {
    const uint32_t w0 = in[t * 3], w1 = in[t * 3 + 1], w2 = in[t * 3 + 2];
    // 96 bits = 16 codes of 6 bits, little endian
    const uint64_t lo = (uint64_t)w0 | ((uint64_t)w1 << 32);  // codes 0..9 (60 bits) + 4 bits of code 10
    const uint64_t hi = ((uint64_t)w1 >> 28) | ((uint64_t)w2 << 4);  // from bit 60: codes 10..15
    uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t code = (k < 10) ? (uint32_t)(lo >> (6 * k)) & 63u : (uint32_t)(hi >> (6 * (k - 10))) & 63u;
      o[k >> 2] |= (uint32_t)sd[code] << (8 * (k & 3));
    }
    out[t] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Hands out the chunks of a _len8 streaming call.  One device: a private planner.  rf_sharded_stream_*_len8*: ONE planner
// shared by the per-device workers, each of which takes its next chunk only when one of its three pipeline slots is free,
// so a device behind a faster PCIe link simply takes more chunks (the 8-GPU box gives GPUs 0-3 23 GB/s and GPUs 4-7 35 GB/s
// when all copy at once: equal static shares wait for the slowest link).
struct Len8Planner {
  std::mutex mu;
  const uint8_t* lens = nullptr;
  uint64_t n = 0, cap_n = 0, cap_bytes = 0;
  bool packed6 = false;
  std::vector<uint32_t> block_sums;  // shared planner: the length sum of every 4096-candidate block, computed up front by all
                                     // host threads (a serial 4 GB/s summing loop under the lock capped 8 GPUs at 4.2e9 pairs/s)
  uint64_t i0 = 0, pos = 0;  // next candidate, its character position
  struct Chunk { uint64_t i0, i1, pos, B0, bytes; };
  // false: nothing left (or *too_small: one block of 4096 candidates does not fit a chunk)
  bool next(Chunk* c, bool* too_small) {
    constexpr uint64_t kBlock = 4096;  // chunk boundaries fall on multiples of kBlock candidates: lengths are summed block-wise
    std::lock_guard<std::mutex> lk(mu);
    *too_small = false;
    if (i0 >= n) return false;
    const uint64_t B0 = packed6 ? (pos & ~63ull) : (pos & ~15ull);  // packed: 64 characters = 48 bytes keep every piece 16-byte aligned
    uint64_t i1 = i0, bytes = pos - B0;
    while (i1 < n && i1 - i0 < cap_n) {
      const uint64_t j1 = (n - i1 < kBlock) ? n : i1 + kBlock;
      if (j1 - i0 > cap_n) break;
      const uint64_t sum = block_sums.empty() ? sum_bytes(lens + i1, j1 - i1) : block_sums[i1 / kBlock];
      if (bytes + sum > cap_bytes) break;
      bytes += sum;
      i1 = j1;
    }
    if (i1 == i0) { *too_small = true; return false; }
    *c = Chunk{i0, i1, pos, B0, bytes};
    pos = B0 + bytes;
    i0 = i1;
    return true;
  }
};

// rf_batch_stream_*_len8: the candidates' lengths cross PCIe as ONE byte each instead of a 4- or 8-byte CSR start (the
// starts of a chunk are rebuilt on the device by a prefix sum), and the results can come back as one byte each: 36.9 + 1
// instead of 39.9 + 4 bytes per config-2 pair on the host side of the link.
// packed6: `chars` is the 6-bit packed stream of rf_pack6_u8 (character i at bits [6i, 6i+6)), `dict` its 64 symbols.
rf_status stream_len8_impl(const rf_batch* b, const uint8_t* chars, const uint8_t* lens, uint64_t n, rf_kind kind,
                           const rf_args* args, void* out_host, bool out_u8, const uint8_t* dict = nullptr,
                           Len8Planner* shared = nullptr) {
  const bool packed6 = dict != nullptr;
  if (!b) return fail(RF_ERR_INVALID_ARG, "NULL handle");
  if ((int)kind < 0 || (int)kind > 3) return fail(RF_ERR_INVALID_ARG, "unknown kind");
  if (rf_result_is_float(b->metric, kind)) return fail(RF_ERR_INVALID_ARG, "this (metric, kind) yields f64 results; the _len8 entry points return integer scores");
  if (n == 0) return RF_OK;
  if (!lens || !out_host) return fail(RF_ERR_INVALID_ARG, "NULL argument");
  if (rf_device_count() <= b->device) return fail(RF_ERR_CUDA, "no such CUDA device");
  // u32-query comparator on byte candidates: the table-driven metrics score with the query mapped into the byte domain
  // (no renaming pass over the chunk); the others rename the chunk through the query's own alphabet
  const rf_batch* sb = b;
  if (b->wide && table_driven(b->metric, args)) {
    sb = byte_sub(b);
    if (!sb) return RF_ERR_CUDA;
  } else if (b->wide && b->alpha_overflow) {
    return fail(RF_ERR_UNSUPPORTED, "this metric compares symbols directly: u32 queries with more than 255 distinct symbols are not supported for it");
  }

  DeviceGuard g(b->device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  StreamCtx* x = stream_ctx(b->device);
  std::lock_guard<std::mutex> lk(x->mu);
  const uint64_t cap_bytes = (uint64_t)(g_stream_mb.load() > 0 ? g_stream_mb.load() : 1) << 20;
  const uint64_t cap_n = (uint64_t)(g_stream_kcand.load() > 0 ? g_stream_kcand.load() : 1) << 10;
  cudaError_t e = stream_ctx_prepare(x, cap_bytes, cap_n);
  if (e != cudaSuccess) return cuda_fail(e, "streaming buffers");
  if (packed6)
    for (auto& sl : x->slot) {
      if (sl.d_packed) continue;
      if ((e = cudaMalloc(&sl.d_packed, cap_bytes / 4 * 3 + 256)) != cudaSuccess) return cuda_fail(e, "streaming buffers");
    }
  Dict64 dk;
  memset(&dk, 0, sizeof(dk));
  if (packed6) memcpy(dk.b, dict, 64);
  for (auto& sl : x->slot) {
    if (sl.d_lens) continue;
    sl.scan_tmp_bytes = lens_to_offsets_tmp_bytes(cap_n);
    if ((e = cudaMalloc(&sl.d_lens, cap_n + 64)) != cudaSuccess) break;
    if ((e = cudaMalloc(&sl.d_scan_tmp, sl.scan_tmp_bytes ? sl.scan_tmp_bytes : 16)) != cudaSuccess) break;
    if ((e = cudaMalloc(&sl.d_out8, cap_n + 64)) != cudaSuccess) break;
  }
  if (e != cudaSuccess) return cuda_fail(e, "streaming buffers");
  uint32_t* d_over = nullptr;
  if ((e = cudaMalloc(&d_over, 16)) != cudaSuccess) return cuda_fail(e, "streaming buffers");
  if ((e = cudaMemset(d_over, 0, 16)) != cudaSuccess) { cudaFree(d_over); return cuda_fail(e, "streaming buffers"); }
  Len8Planner local;
  Len8Planner* plan = shared ? shared : &local;
  if (!shared) {
    local.lens = lens;
    local.n = n;
    local.cap_n = cap_n;
    local.cap_bytes = cap_bytes;
    local.packed6 = packed6;
  }
  rf_status s = RF_OK;
  int k = 0;
  for (auto& sl : x->slot) sl.busy = false;
  for (;;) {
    StreamSlot& sl = x->slot[k];
    k = (k + 1) % kSlots;
    if (sl.busy) {  // back-pressure: the slot's previous chunk must have left the device before the next one is taken
      if ((e = cudaEventSynchronize(sl.done)) != cudaSuccess) { s = cuda_fail(e, "streaming scan"); break; }
      sl.busy = false;
    }
    Len8Planner::Chunk ch;
    bool too_small = false;
    if (!plan->next(&ch, &too_small)) {
      if (too_small) s = fail(RF_ERR_UNSUPPORTED, "stream_chunk_mb / stream_chunk_kcand too small for one block of 4096 candidates");
      break;
    }
    const uint64_t i0 = ch.i0, i1 = ch.i1, pos = ch.pos, B0 = ch.B0, bytes = ch.bytes;
    const uint64_t cn = i1 - i0, B1 = B0 + bytes;
    // The length bytes go first: a copy queued BEHIND the unpack kernel of its own stream would hold up the copy engine's
    // queue -- and with it the next chunk's upload -- until that kernel has run, which in turn waits for SMs the previous
    // chunk's persistent scan kernel still occupies (measured: 44.5 GB/s on the link with the lengths after the unpack
    // kernel against 54.3 GB/s for the unpacked format, which has no kernel between its two uploads).
    e = cudaMemcpyAsync(sl.d_lens, lens + i0, cn, cudaMemcpyHostToDevice, sl.st);
    if (e != cudaSuccess) { s = cuda_fail(e, "chunk upload"); break; }
    if (B1 > B0) {
      if (!chars) { s = fail(RF_ERR_INVALID_ARG, "chars is NULL"); break; }
      if (packed6) {
        const uint64_t n16 = (B1 - B0 + 15) / 16;  // 16-character groups = 12 packed bytes each
        e = cudaMemcpyAsync(sl.d_packed, chars + B0 / 4 * 3, n16 * 12, cudaMemcpyHostToDevice, sl.st);
        if (e == cudaSuccess) {
          const uint64_t blocks = (n16 + 255) / 256;
          unpack6_kernel<<<(uint32_t)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, sl.st>>>((const uint32_t*)sl.d_packed, n16, dk, (uint4*)sl.d_chars);
          rfk::count_launches(1);
          e = cudaGetLastError();
        }
      } else {
        e = cudaMemcpyAsync(sl.d_chars, chars + B0, B1 - B0, cudaMemcpyHostToDevice, sl.st);
      }
    }
    if (e == cudaSuccess) e = lens_to_offsets(sl.d_lens, cn, (uint32_t)(pos - B0), (uint32_t*)sl.d_offs, sl.d_scan_tmp, sl.scan_tmp_bytes, sl.st);
    if (e != cudaSuccess) { s = cuda_fail(e, "chunk upload"); break; }
    rfk::count_launches(1);
    const uint8_t* d_src = sl.d_chars;
    if (b->wide && B1 > B0 && sb == b) {
      if (!sl.d_renamed) {
        if ((e = cudaMalloc(&sl.d_renamed, cap_bytes + 256)) == cudaSuccess) e = cudaMemsetAsync(sl.d_renamed, 0, cap_bytes + 256, sl.st);
        if (e != cudaSuccess) { s = cuda_fail(e, "streaming buffers"); break; }
      }
      const uint64_t cb = B1 - B0, blocks = ((cb + 3) / 4 + 255) / 256;
      const uint32_t grid = (uint32_t)(blocks < 148 * 16 ? blocks : 148 * 16);
      remap_kernel<uint8_t><<<grid, 256, 0, sl.st>>>(sl.d_chars, cb, b->d_alpha_keys, b->d_alpha_codes, (uint32_t*)sl.d_renamed);
      rfk::count_launches(1);
      if ((e = cudaGetLastError()) != cudaSuccess) { s = cuda_fail(e, "alphabet renaming"); break; }
      d_src = sl.d_renamed;
    }
    CorpusView cv{d_src, (const uint32_t*)sl.d_offs, nullptr, cn, bytes - (pos - B0), 255};
    s = score_view(sb, cv, nullptr, b->device, kind, args, sl.d_out, false, sl.st);
    if (s != RF_OK) break;
    if (out_u8) {
      const uint64_t blocks = ((cn + 3) / 4 + 255) / 256;
      narrow_results_u8<<<(uint32_t)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, sl.st>>>((const uint32_t*)sl.d_out, cn, sl.d_out8, d_over);
      rfk::count_launches(1);
      if ((e = cudaGetLastError()) == cudaSuccess)
        e = cudaMemcpyAsync((uint8_t*)out_host + i0, sl.d_out8, cn, cudaMemcpyDeviceToHost, sl.st);
    } else {
      e = cudaMemcpyAsync((uint8_t*)out_host + i0 * 4, sl.d_out, cn * 4, cudaMemcpyDeviceToHost, sl.st);
    }
    if (e == cudaSuccess) e = cudaEventRecord(sl.done, sl.st);
    if (e != cudaSuccess) { s = cuda_fail(e, "chunk download"); break; }
    sl.busy = true;
  }
  for (auto& sl : x->slot) {
    e = cudaStreamSynchronize(sl.st);
    sl.busy = false;
    if (e != cudaSuccess && s == RF_OK) s = cuda_fail(e, "streaming scan");
  }
  uint32_t over = 0;
  if (s == RF_OK && out_u8) {
    e = cudaMemcpy(&over, d_over, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) s = cuda_fail(e, "streaming scan");
    else if (over) s = fail(RF_ERR_INVALID_ARG, "a score above 254 does not fit the u8 result; use rf_batch_stream_u32_len8");
  }
  cudaFree(d_over);
  return s;
}

// rf_batch_stream_*_elems32: host-resident candidates with u32 elements.  A chunk is uploaded as 4-byte elements, renamed
// to the query's byte alphabet on the device (remap_kernel) and scanned like a byte chunk.
rf_status stream_elems32_impl(const rf_batch* b, const uint32_t* elems, const uint64_t* offsets, uint64_t n, rf_kind kind,
                              const rf_args* args, void* out_host, bool want_f64) {
  if (!b) return fail(RF_ERR_INVALID_ARG, "NULL handle");
  if (!b->wide) return fail(RF_ERR_INVALID_ARG, "u32 candidates need a comparator created with rf_batch_create_u32");
  if (b->alpha_overflow) return fail(RF_ERR_UNSUPPORTED, "u32 candidates against a u32 query with more than 255 distinct symbols");
  if ((int)kind < 0 || (int)kind > 3) return fail(RF_ERR_INVALID_ARG, "unknown kind");
  if ((rf_result_is_float(b->metric, kind) != 0) != want_f64)
    return fail(RF_ERR_INVALID_ARG, want_f64 ? "this (metric, kind) yields u32 results; use the _u32 entry point"
                                             : "this (metric, kind) yields f64 results; use the _f64 entry point");
  if (n == 0) return RF_OK;
  if (!offsets || !out_host) return fail(RF_ERR_INVALID_ARG, "NULL argument");
  if (offsets[0] != 0) return fail(RF_ERR_INVALID_ARG, "offsets[0] must be 0");
  if (offsets[n] && !elems) return fail(RF_ERR_INVALID_ARG, "elems is NULL");
  if (rf_device_count() <= b->device) return fail(RF_ERR_CUDA, "no such CUDA device");
  DeviceGuard g(b->device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  StreamCtx* x = stream_ctx(b->device);
  std::lock_guard<std::mutex> lk(x->mu);
  const uint64_t cap_bytes = (uint64_t)(g_stream_mb.load() > 0 ? g_stream_mb.load() : 1) << 20;
  const uint64_t cap_n = (uint64_t)(g_stream_kcand.load() > 0 ? g_stream_kcand.load() : 1) << 10;
  const uint64_t cap_elems = cap_bytes / 4;  // the slot's char buffer holds the chunk's u32 elements
  cudaError_t e = stream_ctx_prepare(x, cap_bytes, cap_n);
  if (e != cudaSuccess) return cuda_fail(e, "streaming buffers");
  const size_t rsz = want_f64 ? 8 : 4;
  rf_status s = RF_OK;
  uint64_t i0 = 0;
  int k = 0;
  while (i0 < n && s == RF_OK) {
    const uint64_t E0 = offsets[i0] & ~15ull;  // element index, a multiple of 16: the renamed bytes keep the 16-byte alignment the tile kernels (TMA) need
    uint64_t hi = (n - i0 < cap_n) ? n : i0 + cap_n;
    if (offsets[hi] - E0 > cap_elems) {
      uint64_t lo = i0;
      while (hi - lo > 1) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (offsets[mid] - E0 <= cap_elems) lo = mid; else hi = mid;
      }
      hi = lo;
      if (hi == i0) { s = fail(RF_ERR_UNSUPPORTED, "a single candidate exceeds the streaming chunk size (raise stream_chunk_mb)"); break; }
    }
    const uint64_t i1 = hi, cn = i1 - i0, E1 = offsets[i1];
    StreamSlot& sl = x->slot[k];
    k = (k + 1) % kSlots;
    if (!sl.d_renamed) {
      if ((e = cudaMalloc(&sl.d_renamed, cap_bytes + 256)) == cudaSuccess) e = cudaMemsetAsync(sl.d_renamed, 0, cap_bytes + 256, sl.st);
      if (e != cudaSuccess) { s = cuda_fail(e, "streaming buffers"); break; }
    }
    if (E1 > E0) e = cudaMemcpyAsync(sl.d_chars, elems + E0, (E1 - E0) * 4, cudaMemcpyHostToDevice, sl.st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sl.d_offs, offsets + i0, (cn + 1) * 8, cudaMemcpyHostToDevice, sl.st);
    if (e != cudaSuccess) { s = cuda_fail(e, "chunk upload"); break; }
    if (E1 > E0) {
      const uint64_t ce = E1 - E0, blocks = ((ce + 3) / 4 + 255) / 256;
      const uint32_t grid = (uint32_t)(blocks < 148 * 16 ? blocks : 148 * 16);
      remap_kernel<uint32_t><<<grid, 256, 0, sl.st>>>((const uint32_t*)sl.d_chars, ce, b->d_alpha_keys, b->d_alpha_codes, (uint32_t*)sl.d_renamed);
      rfk::count_launches(1);
      if ((e = cudaGetLastError()) != cudaSuccess) { s = cuda_fail(e, "alphabet renaming"); break; }
    }
    CorpusView cv{sl.d_renamed - E0, nullptr, (const uint64_t*)sl.d_offs, cn, E1 - offsets[i0], 0};
    s = score_view(b, cv, nullptr, b->device, kind, args, sl.d_out, want_f64, sl.st);
    if (s != RF_OK) break;
    e = cudaMemcpyAsync((uint8_t*)out_host + i0 * rsz, sl.d_out, cn * rsz, cudaMemcpyDeviceToHost, sl.st);
    if (e != cudaSuccess) { s = cuda_fail(e, "chunk download"); break; }
    i0 = i1;
  }
  for (auto& sl : x->slot) {
    e = cudaStreamSynchronize(sl.st);
    if (e != cudaSuccess && s == RF_OK) s = cuda_fail(e, "streaming scan");
  }
  return s;
}
}  // namespace

extern "C" {
rf_status rf_batch_stream_u32(const rf_batch* b, const uint8_t* chars, const uint64_t* offsets, uint64_t n, rf_kind kind,
                              const rf_args* args, uint32_t* out_host) {
  return stream_impl(b, chars, offsets, n, kind, args, out_host, false);
}
rf_status rf_batch_stream_u32_off32(const rf_batch* b, const uint8_t* chars, const uint32_t* offsets, uint64_t n,
                                    rf_kind kind, const rf_args* args, uint32_t* out_host) {
  return stream_impl(b, chars, offsets, n, kind, args, out_host, false);
}
rf_status rf_batch_stream_f64(const rf_batch* b, const uint8_t* chars, const uint64_t* offsets, uint64_t n, rf_kind kind,
                              const rf_args* args, double* out_host) {
  return stream_impl(b, chars, offsets, n, kind, args, out_host, true);
}
rf_status rf_batch_stream_f64_off32(const rf_batch* b, const uint8_t* chars, const uint32_t* offsets, uint64_t n,
                                    rf_kind kind, const rf_args* args, double* out_host) {
  return stream_impl(b, chars, offsets, n, kind, args, out_host, true);
}
rf_status rf_batch_stream_u32_elems32(const rf_batch* b, const uint32_t* elems, const uint64_t* offsets, uint64_t n, rf_kind kind,
                                      const rf_args* args, uint32_t* out_host) {
  return stream_elems32_impl(b, elems, offsets, n, kind, args, out_host, false);
}
rf_status rf_batch_stream_f64_elems32(const rf_batch* b, const uint32_t* elems, const uint64_t* offsets, uint64_t n, rf_kind kind,
                                      const rf_args* args, double* out_host) {
  return stream_elems32_impl(b, elems, offsets, n, kind, args, out_host, true);
}
rf_status rf_batch_stream_u32_len8(const rf_batch* b, const uint8_t* chars, const uint8_t* lens, uint64_t n, rf_kind kind,
                                   const rf_args* args, uint32_t* out_host) {
  return stream_len8_impl(b, chars, lens, n, kind, args, out_host, false);
}
rf_status rf_batch_stream_u8_len8(const rf_batch* b, const uint8_t* chars, const uint8_t* lens, uint64_t n, rf_kind kind,
                                  const rf_args* args, uint8_t* out_host) {
  return stream_len8_impl(b, chars, lens, n, kind, args, out_host, true);
}
rf_status rf_batch_stream_u32_len8_packed6(const rf_batch* b, const uint8_t* packed, const uint8_t* dict64, const uint8_t* lens,
                                           uint64_t n, rf_kind kind, const rf_args* args, uint32_t* out_host) {
  if (!dict64) return fail(RF_ERR_INVALID_ARG, "dict64 is NULL");
  return stream_len8_impl(b, packed, lens, n, kind, args, out_host, false, dict64);
}
rf_status rf_batch_stream_u8_len8_packed6(const rf_batch* b, const uint8_t* packed, const uint8_t* dict64, const uint8_t* lens,
                                          uint64_t n, rf_kind kind, const rf_args* args, uint8_t* out_host) {
  if (!dict64) return fail(RF_ERR_INVALID_ARG, "dict64 is NULL");
  return stream_len8_impl(b, packed, lens, n, kind, args, out_host, true, dict64);
}

// Resident corpus, integer results as BYTES (None = 0xFF): the scan writes its u32 scores into scratch, narrow_results_u8 packs
// them, and a quarter of the bytes cross PCIe (config 2: the 0.4 GB result download of rf_batch_score_u32 takes 3.5x as long as
// the scan itself).  A score above 254 -> RF_ERR_INVALID_ARG after the output is filled (such entries read 0xFF).
static rf_status score_u8_impl(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint8_t* out, bool out_on_device,
                               cudaStream_t user_st) {
  if (!b || !c) return fail(RF_ERR_INVALID_ARG, "NULL handle");
  if ((int)kind < 0 || (int)kind > 3) return fail(RF_ERR_INVALID_ARG, "unknown kind");
  if (rf_result_is_float(b->metric, kind)) return fail(RF_ERR_INVALID_ARG, "this (metric, kind) yields f64 results; byte results exist for integer scores only");
  if (c->n == 0) return score_device(b, c, kind, args, nullptr, false, nullptr);
  if (!out) return fail(RF_ERR_INVALID_ARG, "out is NULL");
  DeviceGuard g(c->device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = user_st;
  cudaError_t e = cudaSuccess;
  if (!out_on_device && (e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
  const size_t n4 = (size_t)c->n * 4, n1 = ((size_t)c->n + 3) / 4 * 4;
  uint8_t* d_buf = nullptr;  // [u32 scores][16 B: hamming error flag, overflow flag][byte scores (host-output calls)]
  e = dev_alloc(&d_buf, n4 + 16 + (out_on_device ? 0 : n1), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_buf + n4, 0, 16, st);
  rf_status s = e == cudaSuccess ? RF_OK : cuda_fail(e, "result buffer");
  uint32_t flags[2] = {0, 0};
  if (s == RF_OK) s = score_device(b, c, kind, args, d_buf, false, st, (uint32_t*)(d_buf + n4));
  if (s == RF_OK) {
    uint8_t* d_out8 = out_on_device ? out : d_buf + n4 + 16;
    const uint64_t blocks = (((uint64_t)c->n + 3) / 4 + 255) / 256;
    narrow_results_u8<<<(uint32_t)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, st>>>((const uint32_t*)d_buf, c->n, d_out8, (uint32_t*)(d_buf + n4 + 4));
    rfk::count_launches(1);
    e = cudaGetLastError();
    if (e == cudaSuccess && !out_on_device) e = cudaMemcpyAsync(out, d_out8, c->n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && !out_on_device) {
      e = cudaMemcpyAsync(flags, d_buf + n4, 8, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (e != cudaSuccess) s = cuda_fail(e, "byte results");
    else if (flags[0]) s = fail(RF_ERR_INVALID_ARG, "Differing length arguments provided");
    else if (flags[1]) s = fail(RF_ERR_INVALID_ARG, "a score above 254 does not fit the u8 result; use rf_batch_score_u32");
  }
  dev_free(d_buf, st);
  if (!out_on_device) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  return s;
}
rf_status rf_batch_score_u8(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint8_t* out) {
  return score_u8_impl(b, c, kind, args, out, false, nullptr);
}
rf_status rf_batch_score_u8_device(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint8_t* out, void* stream) {
  return score_u8_impl(b, c, kind, args, out, true, (cudaStream_t)stream);
}
}  // extern "C"

extern "C" {

// ------------------------------------------------------------------------------------------------ cdist
static rf_status cdist_impl(const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c,
                            const rf_args* args, uint32_t k, uint32_t* idx_out, uint32_t* dist_out, bool out_on_device,
                            cudaStream_t user_stream, bool queries_in_corpus_codes = false, rf_metric metric = RF_LEVENSHTEIN) {
  if (!c) return fail(RF_ERR_INVALID_ARG, "NULL corpus");
  if (metric != RF_LEVENSHTEIN && metric != RF_OSA && metric != RF_INDEL && metric != RF_LCS_SEQ)
    return fail(RF_ERR_UNSUPPORTED, "many-vs-many top-k supports the Levenshtein, OSA, Indel and LCSseq distances");
  const bool bottom = metric == RF_INDEL || metric == RF_LCS_SEQ;  // LCS family: bottom-aligned match tables
  if (c->d_elems32 || (c->compact32 && !queries_in_corpus_codes))
    return fail(RF_ERR_UNSUPPORTED, "rf_cdist_topk_u8 needs a u8 corpus (this one was made by rf_corpus_create_u32; use rf_cdist_topk_u32)");
  if (nq == 0) return RF_OK;
  if (!q_offsets || !idx_out || !dist_out) return fail(RF_ERR_INVALID_ARG, "NULL argument");
  if (k == 0 || k > 64) return fail(RF_ERR_INVALID_ARG, "k must be in 1..64");
  if (q_offsets[0] != 0) return fail(RF_ERR_INVALID_ARG, "q_offsets[0] must be 0");
  if (q_offsets[nq] && !q_chars) return fail(RF_ERR_INVALID_ARG, "q_chars is NULL");
  rf_args def;
  rf_args_default(&def);
  const rf_args* a = args ? args : &def;
  if (metric == RF_LEVENSHTEIN && (a->insertion_cost != 1 || a->deletion_cost != 1 || a->substitution_cost != 1))
    return fail(RF_ERR_UNSUPPORTED, "cdist top-k supports unit Levenshtein weights only");
  uint32_t max_len = 0;
  for (uint32_t q = 0; q < nq; ++q) {
    const uint64_t l = q_offsets[q + 1] - q_offsets[q];
    if (l > 64) return fail(RF_ERR_UNSUPPORTED, "cdist top-k supports queries of at most 64 elements");
    if (l > max_len) max_len = (uint32_t)l;
  }
  DeviceGuard g(c->device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  const size_t kk = (size_t)nq * k;
  cudaStream_t st = user_stream;
  bool own_stream = false;
  if (!out_on_device) {
    RF_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    own_stream = true;
  }
  uint32_t *d_idx = nullptr, *d_dist = nullptr, *d_qlen = nullptr;
  void* d_tabs = nullptr;
  unsigned long long* d_scratch = nullptr;
  rf_status s = RF_OK;
  cudaError_t e = cudaSuccess;
  const bool wide = max_len > 32;
  if (c->n == 0 || !c->lb.gdata) {
    // empty corpus (or no interleaved layout): every row is padding
    if (c->n != 0) s = fail(RF_ERR_UNSUPPORTED, "cdist top-k needs the interleaved layout (build_interleaved_layout=1)");
    else {
      if (out_on_device) {
        e = cudaMemsetAsync(idx_out, 0xFF, kk * 4, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(dist_out, 0xFF, kk * 4, st);
        if (e != cudaSuccess) s = cuda_fail(e, "memset");
      } else {
        memset(idx_out, 0xFF, kk * 4);
        memset(dist_out, 0xFF, kk * 4);
      }
    }
  } else {
    // per-query top-aligned match tables (pattern_match_vector.rs:213-224), built on the host
    const size_t wsz = wide ? 8 : 4;
    std::vector<uint8_t> tabs((size_t)nq * 256 * wsz, 0);
    std::vector<uint32_t> qlen(nq);
    for (uint32_t q = 0; q < nq; ++q) {
      const uint8_t* s1 = q_chars + q_offsets[q];
      const uint32_t l = (uint32_t)(q_offsets[q + 1] - q_offsets[q]);
      qlen[q] = l;
      if (wide) {
        uint64_t* t = (uint64_t*)tabs.data() + (size_t)q * 256;
        for (uint32_t i = 0; i < l; ++i) t[s1[i]] |= 1ull << (bottom ? i : i + 64 - l);
      } else {
        uint32_t* t = (uint32_t*)tabs.data() + (size_t)q * 256;
        for (uint32_t i = 0; i < l; ++i) t[s1[i]] |= 1u << (bottom ? i : i + 32 - l);
      }
    }
    const int sms = sm_count_of(c->device);
    const uint64_t layout_bytes = c->lb.total_rows * 256ull + c->lb.ngroups * 32ull * 8ull;
    uint32_t nslices = cdist_slices(sms, nq, layout_bytes, c->lb.ngroups);
    if (g_cdist_slices.load() > 0) nslices = (uint32_t)g_cdist_slices.load();
    do {
      if ((e = cudaMalloc(&d_tabs, tabs.size())) != cudaSuccess) break;
      if ((e = cudaMalloc(&d_qlen, nq * 4)) != cudaSuccess) break;
      if ((e = cudaMalloc(&d_scratch, ((size_t)nq * nslices * k + 1) * 8)) != cudaSuccess) break;
      if (!out_on_device) {
        if ((e = cudaMalloc(&d_idx, kk * 4)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_dist, kk * 4)) != cudaSuccess) break;
      }
      if ((e = cudaMemcpyAsync(d_tabs, tabs.data(), tabs.size(), cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync(d_qlen, qlen.data(), nq * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
      CdistLaunch L{};
      L.lb = LbView{c->lb.perm, c->lb.lens, c->lb.goff, c->lb.gdata, c->lb.ngroups};
      L.total_rows = c->lb.total_rows;
      L.q_tabs = d_tabs;
      L.wide = wide ? 1 : 0;
      L.q_len = d_qlen;
      L.nq = nq;
      L.k = k;
      L.has_cutoff = a->has_cutoff ? 1 : 0;
      L.cutoff = (uint32_t)(a->cutoff_u > 0xFFFFFFFEull ? 0xFFFFFFFEull : a->cutoff_u);
      L.out_idx = out_on_device ? idx_out : d_idx;
      L.out_dist = out_on_device ? dist_out : d_dist;
      L.scratch = d_scratch + 1;
      L.counter = d_scratch;
      L.nslices = nslices;
      L.grid = cdist_grid(sms);
      L.skip = g_cdist_skip.load();
      L.metric = (int)metric;
      L.stream = st;
      if ((e = launch_cdist_topk(L)) != cudaSuccess) break;
      if (!out_on_device) {
        if ((e = cudaMemcpyAsync(idx_out, d_idx, kk * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(dist_out, d_dist, kk * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
      }
      // the pageable host tables and the scratch buffers must outlive the asynchronous work
      e = cudaStreamSynchronize(st);
    } while (0);
    if (e != cudaSuccess) s = cuda_fail(e, "cdist top-k");
  }
  if (d_tabs) cudaFree(d_tabs);
  if (d_qlen) cudaFree(d_qlen);
  if (d_scratch) cudaFree(d_scratch);
  if (d_idx) cudaFree(d_idx);
  if (d_dist) cudaFree(d_dist);
  if (own_stream) cudaStreamDestroy(st);
  return s;
}

rf_status rf_cdist_topk_u8_device(const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c,
                                  const rf_args* args, uint32_t k, uint32_t* idx_device, uint32_t* dist_device,
                                  void* stream) {
  return cdist_impl(q_chars, q_offsets, nq, c, args, k, idx_device, dist_device, true, (cudaStream_t)stream);
}
rf_status rf_cdist_topk_u8(const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c,
                           const rf_args* args, uint32_t k, uint32_t* idx_host, uint32_t* dist_host) {
  return cdist_impl(q_chars, q_offsets, nq, c, args, k, idx_host, dist_host, false, nullptr);
}

// the same many-vs-many top-k for the other bit-parallel distances: RF_LEVENSHTEIN, RF_OSA, RF_INDEL, RF_LCS_SEQ
rf_status rf_cdist_topk_metric_u8(rf_metric metric, const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c,
                                  const rf_args* args, uint32_t k, uint32_t* idx_host, uint32_t* dist_host) {
  return cdist_impl(q_chars, q_offsets, nq, c, args, k, idx_host, dist_host, false, nullptr, false, metric);
}
rf_status rf_cdist_topk_metric_u8_device(rf_metric metric, const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq,
                                         const rf_corpus* c, const rf_args* args, uint32_t k, uint32_t* idx_device,
                                         uint32_t* dist_device, void* stream) {
  return cdist_impl(q_chars, q_offsets, nq, c, args, k, idx_device, dist_device, true, (cudaStream_t)stream, false, metric);
}

// u32-element queries against a u32 corpus that was renamed to bytes at creation (at most 255 distinct symbols, see
// rf_corpus_create_u32): the queries are renamed through the corpus' dictionary on the host (symbols the corpus never
// contains become 0, which matches nothing) and take the byte kernel.
static rf_status cdist_u32_impl(const uint32_t* q_elems, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c,
                                const rf_args* args, uint32_t k, uint32_t* idx_out, uint32_t* dist_out, bool on_device,
                                cudaStream_t st) {
  if (!c) return fail(RF_ERR_INVALID_ARG, "NULL corpus");
  if (nq == 0) return RF_OK;
  if (!q_offsets) return fail(RF_ERR_INVALID_ARG, "NULL argument");
  if (q_offsets[nq] && !q_elems) return fail(RF_ERR_INVALID_ARG, "q_elems is NULL");
  if (c->d_elems32)
    return fail(RF_ERR_UNSUPPORTED, "rf_cdist_topk_u32: the corpus holds more than 255 distinct symbols (not renamed to bytes)");
  std::vector<uint8_t> renamed(q_offsets[nq]);
  for (uint64_t i = 0; i < q_offsets[nq]; ++i) {
    const uint32_t x = q_elems[i];
    if (c->compact32) {
      uint32_t slot = alpha_hash(x);
      while (c->dict_codes[slot] && c->dict_keys[slot] != x) slot = (slot + 1) & (kAlphaSlots - 1);
      renamed[i] = c->dict_codes[slot];  // 0 (never a candidate code) when the corpus does not contain x
    } else {
      // a plain byte corpus may hold every byte value, so a wider symbol has no byte that "matches nothing": refuse
      if (x > 255) return fail(RF_ERR_UNSUPPORTED, "rf_cdist_topk_u32 on a u8 corpus: query symbols beyond 255");
      renamed[i] = (uint8_t)x;
    }
  }
  return cdist_impl(renamed.data(), q_offsets, nq, c, args, k, idx_out, dist_out, on_device, st, true);
}
rf_status rf_cdist_topk_u32(const uint32_t* q_elems, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c,
                            const rf_args* args, uint32_t k, uint32_t* idx_host, uint32_t* dist_host) {
  return cdist_u32_impl(q_elems, q_offsets, nq, c, args, k, idx_host, dist_host, false, nullptr);
}
rf_status rf_cdist_topk_u32_device(const uint32_t* q_elems, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c,
                                   const rf_args* args, uint32_t k, uint32_t* idx_device, uint32_t* dist_device, void* stream) {
  return cdist_u32_impl(q_elems, q_offsets, nq, c, args, k, idx_device, dist_device, true, (cudaStream_t)stream);
}

rf_status rf_topk_merge_device(const uint32_t* idx_parts, const uint32_t* dist_parts, uint64_t part_stride,
                               const uint64_t* index_base_device, uint32_t parts, uint32_t nq, uint32_t k,
                               uint64_t* idx_out_device, uint32_t* dist_out_device, int device, void* stream) {
  if (nq == 0) return RF_OK;
  if (!idx_parts || !dist_parts || !index_base_device || !idx_out_device || !dist_out_device)
    return fail(RF_ERR_INVALID_ARG, "NULL argument");
  if (parts == 0 || k == 0) return fail(RF_ERR_INVALID_ARG, "parts and k must be positive");
  if ((uint64_t)parts * k > 25600) return fail(RF_ERR_UNSUPPORTED, "parts * k must not exceed 25600");
  if (part_stride < (uint64_t)nq * k) return fail(RF_ERR_INVALID_ARG, "part_stride < nq * k");
  DeviceGuard g(device);
  if (!g.ok) return fail(RF_ERR_CUDA, "cudaSetDevice failed");
  cudaError_t e = launch_topk_merge(idx_parts, dist_parts, part_stride, (const unsigned long long*)index_base_device, parts,
                                    nq, k, (unsigned long long*)idx_out_device, dist_out_device, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "top-k merge");
  return RF_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ internal (rf_sharded.cu)
namespace rfi {
rf_status fail(rf_status s, const std::string& msg) { return ::fail(s, msg); }
rf_status cuda_fail(cudaError_t e, const char* what) { return ::cuda_fail(e, what); }
const std::string& last_error() { return g_last_error; }
void set_last_error(const std::string& msg) { g_last_error = msg; }
rf_status score_device(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, void* out_dev, bool want_f64,
                       cudaStream_t st, uint32_t* d_err) {
  return ::score_device(b, c, kind, args, out_dev, want_f64, st, d_err);
}
rf_status cdist(const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c, const rf_args* args, uint32_t k,
                uint32_t* idx_out, uint32_t* dist_out, bool out_on_device, cudaStream_t stream) {
  return ::cdist_impl(q_chars, q_offsets, nq, c, args, k, idx_out, dist_out, out_on_device, stream);
}
rf_status select_host(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, bool want_f64, bool filter, uint32_t k,
                      uint64_t cap, uint32_t* idx_out, void* score_out, uint32_t* n32_out, uint64_t* n64_out) {
  return ::select_host(b, c, kind, args, want_f64, filter, k, cap, idx_out, score_out, n32_out, n64_out);
}
rf_status stream_u64(const rf_batch* b, const uint8_t* chars, const uint64_t* offsets, uint64_t n, rf_kind kind, const rf_args* args,
                     void* out_host, bool want_f64) {
  return stream_impl(b, chars, offsets, n, kind, args, out_host, want_f64, true);
}
rf_status stream_u32(const rf_batch* b, const uint8_t* chars, const uint32_t* offsets, uint64_t n, rf_kind kind, const rf_args* args,
                     void* out_host, bool want_f64) {
  return stream_impl(b, chars, offsets, n, kind, args, out_host, want_f64, true);
}
int sm_count_of(int device) { return ::sm_count_of(device); }
rf_status score_device_range(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, void* out_dev, bool want_f64,
                             cudaStream_t st, uint64_t r0, uint64_t r1) {
  if (!b || !c) return ::fail(RF_ERR_INVALID_ARG, "NULL handle");
  if (b->wide || c->d_elems32 || c->compact32) return ::fail(RF_ERR_UNSUPPORTED, "candidate sub-ranges need a byte comparator and a byte corpus");
  return ::score_view(b, CorpusView{c->d_chars, c->d_off32, c->d_off64, c->n, c->total, c->max_len}, &c->lb, c->device, kind, args,
                      out_dev, want_f64, st, nullptr, r0, r1);
}
// rf_sharded_stream_*_len8*: the per-device workers of one call share `plan` (created by stream_len8_plan_create)
void* stream_len8_plan_create(const uint8_t* lens, uint64_t n, bool packed6) {
  Len8Planner* p = new (std::nothrow) Len8Planner();
  if (!p) return nullptr;
  p->lens = lens;
  p->n = n;
  p->cap_bytes = (uint64_t)(g_stream_mb.load() > 0 ? g_stream_mb.load() : 1) << 20;
  p->cap_n = (uint64_t)(g_stream_kcand.load() > 0 ? g_stream_kcand.load() : 1) << 10;
  p->packed6 = packed6;
  const uint64_t nblocks = (n + 4095) / 4096;
  p->block_sums.resize(nblocks);
#pragma omp parallel for schedule(static)
  for (int64_t bl = 0; bl < (int64_t)nblocks; ++bl) {
    const uint64_t a = (uint64_t)bl * 4096, z = a + 4096 < n ? a + 4096 : n;
    p->block_sums[bl] = (uint32_t)sum_bytes(lens + a, z - a);
  }
  return p;
}
void stream_len8_plan_destroy(void* plan) { delete static_cast<Len8Planner*>(plan); }
rf_status stream_len8_shared(const rf_batch* b, const uint8_t* chars, const uint8_t* dict64, const uint8_t* lens, uint64_t n,
                             rf_kind kind, const rf_args* args, void* out_host, bool out_u8, void* plan) {
  return stream_len8_impl(b, chars, lens, n, kind, args, out_host, out_u8, dict64, static_cast<Len8Planner*>(plan));
}
rf_status corpus_create_sub(const uint8_t* chars, const uint64_t* offsets, uint64_t lo, uint64_t hi, int device, rf_corpus** out) {
  return ::corpus_create_host(chars, offsets + lo, true, hi - lo, device, out, offsets[lo]);
}
}  // namespace rfi
