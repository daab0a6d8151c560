// rf_internal.h -- handle layouts and the internal entry points shared by rf_api.cu and rf_sharded.cu (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/rfgpu.h"
#include "rf_kernels.cuh"

struct rf_corpus {
  int device = 0;
  uint64_t n = 0, total = 0, max_len = 0;  // max_len: longest candidate (elements)
  bool has_negative = false, has_huge = false;  // rf_corpus_create_elems: a negative signed value / an unsigned value >= 2^31 was seen
  bool csr_released = false;      // rf_corpus_release_csr: d_chars / d_off* are gone, only the interleaved layout is left
  uint8_t* d_chars = nullptr;     // u8 elements ...
  uint32_t* d_elems32 = nullptr;  // ... or u32 elements (rf_corpus_create_u32); exactly one of the two is set
  uint32_t* d_off32 = nullptr;
  uint64_t* d_off64 = nullptr;
  rfk::LbAlloc lb;  // length-bucketed interleaved copy for the single-word kernels
  // rf_corpus_create_u32 with at most 255 distinct symbols in the whole corpus: the symbols are renamed to the bytes
  // 1..D ONCE at creation and the corpus is kept (and scored) as a u8 corpus; d_elems32 is released.  The dictionary
  // stays on the host: a u32 comparator renames its query through it (absent symbols -> 0, which matches nothing).
  bool compact32 = false;
  uint64_t dict_serial = 0;
  std::vector<uint32_t> dict_keys;   // [kAlphaSlots] open addressing (alpha_hash)
  std::vector<uint8_t> dict_codes;   // [kAlphaSlots] 0 = empty slot
};

// kernel-choice knobs, copied from the process-wide defaults (rf_set_option) when the comparator is created and changed
// per comparator with rf_batch_set_option
struct rf_batch_opts {
  int w1_path = 0, mw_path = 0, band = 1, jaro32 = 1, epi_table = 1;
};

struct rf_batch {
  rf_batch_opts opt;
  bool has_negative = false, has_huge = false;  // rf_batch_create_elems (see rf_corpus)
  int device = 0;
  rf_metric metric = RF_LEVENSHTEIN;
  std::vector<uint8_t> s1;
  uint32_t len1 = 0, words = 0;
  uint8_t* d_blob = nullptr;  // all tables in one allocation
  rfk::QueryView view{};
  // rf_batch_create_u32: the query's distinct symbols are renamed to the bytes 1..D (D <= 255); candidates are
  // renamed on the device per scoring call (symbols the query does not contain become 0, which matches nothing).
  // Every metric here depends only on which (query, candidate) positions are equal, so the result is exact.
  bool wide = false;
  bool alpha_overflow = false;  // wide query with more than 255 distinct symbols: no byte alphabet of its own (see rf_batch_create_u32)
  std::vector<uint32_t> s1w;         // the u32 query as given
  mutable std::mutex sub_mu;         // byte comparators of this query against compact u32 corpora, by dictionary
  mutable std::unordered_map<uint64_t, rf_batch*> subs;
  // alpha_overflow queries against u32 corpora with more than 255 distinct symbols: 16-bit codes (built on first use)
  mutable bool w16_ready = false;
  mutable uint32_t w16_slots = 0;             // power of two
  mutable uint32_t* d_w16_keys = nullptr;     // [w16_slots] symbol ...
  mutable uint16_t* d_w16_codes = nullptr;    // ... -> code 1..D, 0 = empty slot / symbol not in the query
  mutable uint64_t* d_w16_pm = nullptr;       // [(D + 1)][words] match vectors per code (row 0 = all zero)
  uint32_t* d_alpha_keys = nullptr;  // [kAlphaSlots] open-addressing table: symbol ...
  uint8_t* d_alpha_codes = nullptr;  // ... -> byte code, 0 = empty slot
};


namespace rfi {
// status + thread-local message (rf_last_error)
rf_status fail(rf_status s, const std::string& msg);
rf_status cuda_fail(cudaError_t e, const char* what);
const std::string& last_error();
void set_last_error(const std::string& msg);
// enqueue the scoring of every candidate of `c` on `st` (rf_batch_score_*_device)
rf_status score_device(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, void* out_dev, bool want_f64,
                       cudaStream_t st, uint32_t* d_err);
// the same for the candidates [r0, r1) only (r0 a multiple of 65536; results at out_dev[candidate index]); byte comparator
// + byte corpus, bit-parallel / Jaro metrics
rf_status score_device_range(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, void* out_dev, bool want_f64,
                             cudaStream_t st, uint64_t r0, uint64_t r1);
// rf_cdist_topk_u8[_device]
rf_status cdist(const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c, const rf_args* args, uint32_t k,
                uint32_t* idx_out, uint32_t* dist_out, bool out_on_device, cudaStream_t stream);
// rf_batch_extract_* / rf_batch_filter_*
rf_status select_host(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, bool want_f64, bool filter, uint32_t k,
                      uint64_t cap, uint32_t* idx_out, void* score_out, uint32_t* n32_out, uint64_t* n64_out);
// rf_batch_stream_*: candidates [0, n) described by offsets[0..n] (absolute positions in `chars`; offsets[0] need not be 0)
rf_status stream_u64(const rf_batch* b, const uint8_t* chars, const uint64_t* offsets, uint64_t n, rf_kind kind, const rf_args* args,
                     void* out_host, bool want_f64);
rf_status stream_u32(const rf_batch* b, const uint8_t* chars, const uint32_t* offsets, uint64_t n, rf_kind kind, const rf_args* args,
                     void* out_host, bool want_f64);
// _len8 streaming with the chunks handed out by ONE planner shared between the per-device workers of a sharded call
void* stream_len8_plan_create(const uint8_t* lens, uint64_t n, bool packed6);
void stream_len8_plan_destroy(void* plan);
rf_status stream_len8_shared(const rf_batch* b, const uint8_t* chars, const uint8_t* dict64, const uint8_t* lens, uint64_t n,
                             rf_kind kind, const rf_args* args, void* out_host, bool out_u8, void* plan);
int sm_count_of(int device);
// corpus of the candidates [lo, hi) of a larger host CSR (chars = the larger array's start, offsets = its full index)
rf_status corpus_create_sub(const uint8_t* chars, const uint64_t* offsets, uint64_t lo, uint64_t hi, int device, rf_corpus** out);
}  // namespace rfi
