// rf_kernels.cuh -- launch-side view of the sm_100a scoring kernels (rf_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "rf_core.cuh"

namespace rfk {

// Packed corpus resident in HBM (CSR).  `chars` is 16-byte aligned and has >= 64 bytes of zeroed slack
// behind the data; the offset array has >= 16 entries of slack (TMA bulk copies move 16-byte multiples).
struct CorpusView {
  const uint8_t* chars;
  const uint32_t* off32;  // exactly one of off32/off64 is non-null
  const uint64_t* off64;
  uint64_t n;
  uint64_t total;
  uint64_t max_len;  // longest candidate, 0 = unknown (sizes per-warp scratch of the long-query kernels)
};

// One cached query (BatchComparator::new): compact match tables in device memory.
struct QueryView {
  uint32_t len1;
  uint32_t words;            // ceil(len1/64)
  const uint32_t* tab32_top; // [256] PM << (32-len1)   (len1 <= 32)   Levenshtein / OSA
  const uint32_t* tab32_bot; // [256] PM                (len1 <= 32)   LCS family
  const uint64_t* tab64_top; // [256] PM << (64-len1)   (len1 <= 64)
  const uint64_t* tab64_bot; // [256] PM                (len1 <= 64)   LCS family, Jaro
  const uint64_t* pm_words;  // [256][words] row-major (pattern_match_vector.rs layout), any len1
  const uint32_t* pm_band;   // [256][band_stride] 32-bit words: 2 zero words, the match vector, >= 2 zero words (banded kernel)
  uint32_t band_stride;      // (2*(words+2)) | 1
  const uint8_t* qbytes;     // the query itself, zero-padded to a multiple of 16 plus 16 (hamming / prefix / postfix)
  const double* quot;        // [65][65] exactly rounded a/b (b >= 1): the Jaro formula's quotients without divisions
  // queries of 65..512 elements (scan_lbn_kernel): the match vectors as ONE integer of `limbs` = 4*ceil(len1/128)
  // 32-bit limbs, little endian; top-aligned (PM << (32*limbs - len1): Levenshtein / OSA) and bottom-aligned (LCS family)
  const uint32_t* pmn_top;   // [256][limbs] or null
  const uint32_t* pmn_bot;   // [256][limbs] or null
  uint32_t limbs;
};

// Length-bucketed, warp-interleaved copy of the corpus (built once at corpus creation, rf_layout.cu):
// candidates are sorted by length inside blocks of LB_BLOCK candidates and cut into groups of 32; group g
// stores bytes [8k, 8k+8) of its lane-l candidate at ((uint2*)gdata)[(goff[g] + k) * 32 + l]  (rows of 256 bytes).
struct LbView {
  const uint32_t* perm;   // [ngroups*32] original candidate index, 0xFFFFFFFF = padding lane
  const uint32_t* lens;   // [ngroups*32] candidate length
  const uint64_t* goff;   // [ngroups+1]  first row of each group
  const uint32_t* gdata;  // [total_rows*64 (+ slack)] viewed as uint2 rows
  uint64_t ngroups;
};
constexpr uint32_t LB_BLOCK = 65536;
// longest query of the O(len1 * len2) DP kernels (generic Levenshtein weights, Damerau-Levenshtein): their copy of the
// query lives in shared memory
constexpr uint32_t kDpMaxQuery = 200000;

struct LbAlloc {  // owning pointers of an LbView
  uint32_t* perm = nullptr;
  uint32_t* lens = nullptr;
  uint64_t* goff = nullptr;
  uint32_t* gdata = nullptr;
  uint64_t ngroups = 0;
  uint64_t total_rows = 0;
};
// Builds the layout from the CSR corpus on `stream` (synchronises once to size the data array).
cudaError_t lb_build(const CorpusView& c, cudaStream_t stream, LbAlloc* out);
void lb_free(LbAlloc* a, cudaStream_t stream);

// u8 candidate lengths of a streaming chunk -> its u32 CSR starts (cn + 1 entries, first = init); rf_layout.cu
size_t lens_to_offsets_tmp_bytes(uint64_t cap_n);
cudaError_t lens_to_offsets(const uint8_t* d_lens, uint64_t cn, uint32_t init, uint32_t* d_off, void* tmp, size_t tmp_bytes,
                            cudaStream_t st);

struct ScanLaunch {
  CorpusView corpus;
  LbView lb;
  unsigned long long* lb_counter;  // 8 bytes of device scratch for the chunk scheduler (one per in-flight launch)
  unsigned long long* lb_flag;     // 8 more bytes: scan_jaro32_kernel -> jaro32_long_kernel hand-over flag
  QueryView query;
  Epi epi;
  void* out;          // uint32_t[n] or double[n]
  int out_is_f64;
  cudaStream_t stream;
  int sm_count;
  int elem16;         // corpus.chars is an array of uint16_t codes and query.pm_words has one row per code (multi-word kernels only)
  int jaro32;         // Jaro / Jaro-Winkler, query <= 64: 1 = row-wise kernels + table epilogue, 2 = per-pair epilogue, 3 = 48-register build, 0 = generic per-lane routine
  int epi_table = 1;  // integer metrics on the interleaved layout: 1 = score algebra as a per-launch table (candidates <= 255 elements), 0 = per pair
};

// Single-word path (query <= 64), CSR input: thread per candidate over TMA-staged tiles bucketed by length in-kernel.
cudaError_t launch_scan_w1(const ScanLaunch& L);
// Single-word path (query <= 64), pre-bucketed interleaved layout: warp per group of 32 equal-length candidates.
cudaError_t launch_scan_lb(const ScanLaunch& L);
// Same layout, rows streamed by per-warp TMA bulk copies into a shared-memory ring (sequential metrics; Jaro falls back).
cudaError_t launch_scan_lbr(const ScanLaunch& L);
// Multi-word path, query 65..512, interleaved layout: thread per candidate, the whole column in registers.
cudaError_t launch_scan_lbn(const ScanLaunch& L);
// Multi-word path (query > 64): sub-warp per candidate, carries propagated with warp shuffles.
cudaError_t launch_scan_mw(const ScanLaunch& L);
// Queries beyond 16 384 elements: warp per candidate, the column in stripes of 256 blocks with the carries parked in scratch.
cudaError_t launch_scan_long(const ScanLaunch& L);
// Levenshtein distance with a cutoff of at most 63 unit edits, any query length: one 64-bit sliding band per candidate.
cudaError_t launch_scan_band(const ScanLaunch& L, uint32_t cut);
// Hamming / Prefix / Postfix (any query length): thread per candidate over the CSR corpus.
cudaError_t launch_simple(const ScanLaunch& L, uint32_t* err_flag);
// Levenshtein with generic weights: Wagner-Fischer, query <= 2048 (cudaErrorNotSupported beyond).
cudaError_t launch_wf(const ScanLaunch& L);
// Damerau-Levenshtein (Zhao-Sahni), query <= 2048 (cudaErrorNotSupported beyond).
cudaError_t launch_dl(const ScanLaunch& L);
// Jaro / Jaro-Winkler with a multi-word query (65..2048).
cudaError_t launch_jaro_mw(const ScanLaunch& L);

// many-vs-many Levenshtein top-k (queries <= 64) over the interleaved layout.
struct CdistLaunch {
  LbView lb;
  uint64_t total_rows;
  const void* q_tabs;       // [nq][256] top-aligned tables: uint32_t if !wide (all queries <= 32) else uint64_t
  int wide;
  const uint32_t* q_len;    // [nq]
  uint32_t nq;
  uint32_t k;               // 1..64
  int has_cutoff;
  uint32_t cutoff;
  uint32_t* out_idx;        // [nq][k]
  uint32_t* out_dist;       // [nq][k]
  unsigned long long* scratch;  // [nq][nslices][k] packed (dist<<32|idx) per-slice candidates (unused for 1 slice)
  unsigned long long* counter;  // device, one word: the unit dispenser
  uint32_t nslices;         // corpus slices (cdist_slices); work units = nslices * nq
  uint32_t grid;            // persistent CTAs (cdist_grid)
  int skip;                 // 1: skip groups that cannot reach the current k-th distance by length alone
  int metric;               // M_LEVENSHTEIN (default 0) / M_OSA / M_INDEL / M_LCS_SEQ; tables bottom-aligned for the last two
  cudaStream_t stream;
};
cudaError_t launch_cdist_topk(const CdistLaunch& L);
uint32_t cdist_grid(int sm_count);
uint32_t cdist_slices(int sm_count, uint32_t nq, uint64_t layout_bytes, uint64_t ngroups);

// Global top-k of a sharded corpus from the gathered per-shard lists (rf_select.cu topk_merge_kernel).
cudaError_t launch_topk_merge(const uint32_t* idx_parts, const uint32_t* dist_parts, uint64_t part_stride,
                              const unsigned long long* base, uint32_t parts, uint32_t nq, uint32_t k,
                              unsigned long long* out_idx, uint32_t* out_dist, cudaStream_t stream);

// Result post-processing over a device-resident score vector (rf_select.cu): the k best by (score best-first,
// index ascending), or every score that is not None in index order.
struct SelectLaunch {
  const void* scores;   // u32[n] (0xFFFFFFFF = None) or f64[n] (NaN = None)
  uint64_t n;
  int f64;              // element type of scores
  int desc;             // 1: larger score is better (similarity kinds)
  int filter;           // 0: top-k, 1: all non-None
  uint32_t k;           // top-k: 1..1024
  uint64_t cap;         // filter: capacity of the outputs
  uint32_t* out_idx;    // device, [k] or [cap]
  void* out_score;      // device, same element type as scores
  uint32_t* out_n32;    // device, top-k: number of entries written (< k when fewer candidates qualify)
  unsigned long long* out_n64;  // device, filter: total number of hits (may exceed cap)
  cudaStream_t stream;
  int sm_count;
};
cudaError_t launch_select(const SelectLaunch& L);
void count_launches(uint64_t n);

uint64_t kernel_launch_count();

// Stream-ordered device allocations from the device's default memory pool (release threshold = never), so that
// creating / destroying multi-GB corpora repeatedly does not pay cudaMalloc / cudaFree page-table work each time.
cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t st);
template <class T>
inline cudaError_t dev_alloc(T** p, size_t bytes, cudaStream_t st) { return dev_alloc(reinterpret_cast<void**>(p), bytes, st); }
void dev_free(void* p, cudaStream_t st);
cudaStream_t util_stream(int device);  // internal non-blocking stream (frees issued from destroy functions)

}  // namespace rfk
