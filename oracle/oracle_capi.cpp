// oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY: C entry points over rf_oracle.hpp / rf_textbook.hpp so
// that pytest (ctypes) and bench.py's cpu_baseline / --impl reference legs can drive the CPU oracle.
// Build: see oracle/Makefile (g++ -O3 -march=native -fopenmp -shared -fPIC).
#include <cmath>
#include <cstdint>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "rf_oracle.hpp"
#include "rf_textbook.hpp"

extern "C" {

// mirrors rf_args of include/rfgpu.h field by field (kept separate so the oracle has no product dependency)
struct orc_args {
  uint8_t has_cutoff;
  uint64_t cutoff_u;
  double cutoff_f;
  uint8_t has_hint;
  uint64_t hint_u;
  double hint_f;
  uint64_t ins, del, sub;
  double prefix_weight;
  uint8_t reference_quirks;
  uint8_t pad;
};

static rfo::Args to_args(const orc_args* a) {
  rfo::Args r;
  if (!a) return r;
  r.has_cutoff = a->has_cutoff != 0;
  r.cutoff_u = a->cutoff_u;
  r.cutoff_f = a->cutoff_f;
  r.has_hint = a->has_hint != 0;
  r.hint_u = a->hint_u;
  r.hint_f = a->hint_f;
  r.weights = {a->ins, a->del, a->sub};
  r.prefix_weight = a->prefix_weight;
  r.reference_quirks = a->reference_quirks != 0;
  r.pad = a->pad != 0;
  return r;
}

// Is the (metric, kind) result integer-valued (u32 out) or float-valued (f64 out)?
int orc_result_is_float(int metric, int kind) {
  if (metric == rfo::JARO || metric == rfo::JARO_WINKLER || metric == rfo::RATIO) return 1;
  return (kind == rfo::NORM_DISTANCE || kind == rfo::NORM_SIMILARITY) ? 1 : 0;
}

}  // extern "C"

template <class CQ, class CS>
static int batch_impl(int metric, int kind, const CQ* q, uint64_t qlen, const CS* chars, const uint64_t* offsets,
                      uint64_t n, const orc_args* a_, uint32_t* out_u32, double* out_f64, int nthreads) {
  rfo::Args a = to_args(a_);
  rfo::Batch<CQ> b((rfo::Metric)metric, q, (size_t)qlen);
  const bool is_f = orc_result_is_float(metric, kind) != 0;
  if (is_f && !out_f64) return 1;
  if (!is_f && !out_u32) return 1;
  (void)nthreads;
  int differing = 0;  // hamming without pad: Err(DifferentLengthArgs) (hamming.rs:232-234) for that candidate
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads) reduction(|:differing)
#endif
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const CS* s = chars + offsets[i];
    uint64_t len = offsets[i + 1] - offsets[i];
    if (metric == rfo::HAMMING && !a.pad && len != qlen) {
      differing |= 1;
      if (is_f) out_f64[i] = std::numeric_limits<double>::quiet_NaN(); else out_u32[i] = UINT32_MAX;
      continue;
    }
    if (is_f) {
      rfo::OptF r = b.float_score((rfo::Kind)kind, s, len, a);
      out_f64[i] = r.some ? r.v : std::numeric_limits<double>::quiet_NaN();
    } else {
      rfo::OptU r = b.int_score((rfo::Kind)kind, s, len, a);
      out_u32[i] = r.some ? (uint32_t)r.v : UINT32_MAX;
    }
  }
  return differing ? 3 : 0;
}

extern "C" {

int orc_batch_u8(int metric, int kind, const uint8_t* q, uint64_t qlen, const uint8_t* chars, const uint64_t* offsets,
                 uint64_t n, const orc_args* a, uint32_t* out_u32, double* out_f64, int nthreads) {
  return batch_impl(metric, kind, q, qlen, chars, offsets, n, a, out_u32, out_f64, nthreads);
}
int orc_batch_u32(int metric, int kind, const uint32_t* q, uint64_t qlen, const uint32_t* chars, const uint64_t* offsets,
                  uint64_t n, const orc_args* a, uint32_t* out_u32, double* out_f64, int nthreads) {
  return batch_impl(metric, kind, q, qlen, chars, offsets, n, a, out_u32, out_f64, nthreads);
}

// Raw (un-truncated) single-pair results for the known-answer tests: *some = 0 means None.
int orc_pair_u8(int metric, int kind, const uint8_t* q, uint64_t qlen, const uint8_t* s, uint64_t slen,
                const orc_args* a_, uint64_t* out_u, double* out_f, int* some) {
  rfo::Args a = to_args(a_);
  if (metric == rfo::HAMMING && !a.pad && qlen != slen) { *some = 0; return 3; }  // Err(DifferentLengthArgs)
  rfo::Batch<uint8_t> b((rfo::Metric)metric, q, (size_t)qlen);
  if (orc_result_is_float(metric, kind)) {
    rfo::OptF r = b.float_score((rfo::Kind)kind, s, slen, a);
    *out_f = r.v; *some = r.some;
  } else {
    rfo::OptU r = b.int_score((rfo::Kind)kind, s, slen, a);
    *out_u = r.v; *some = r.some;
  }
  return 0;
}
int orc_pair_u32(int metric, int kind, const uint32_t* q, uint64_t qlen, const uint32_t* s, uint64_t slen,
                 const orc_args* a_, uint64_t* out_u, double* out_f, int* some) {
  rfo::Args a = to_args(a_);
  if (metric == rfo::HAMMING && !a.pad && qlen != slen) { *some = 0; return 3; }  // Err(DifferentLengthArgs)
  rfo::Batch<uint32_t> b((rfo::Metric)metric, q, (size_t)qlen);
  if (orc_result_is_float(metric, kind)) {
    rfo::OptF r = b.float_score((rfo::Kind)kind, s, slen, a);
    *out_f = r.v; *some = r.some;
  } else {
    rfo::OptU r = b.int_score((rfo::Kind)kind, s, slen, a);
    *out_u = r.v; *some = r.some;
  }
  return 0;
}

// Direct access to individual reference kernels, so tests can force a specific code path
// (e.g. the block kernel on short inputs) irrespective of the dispatcher.
uint64_t orc_lev_block_u8(const uint8_t* q, uint64_t qlen, const uint8_t* s, uint64_t slen, uint64_t cutoff) {
  rfo::BlockPM pm(q, (size_t)qlen);
  if (qlen == 0) return slen;
  return rfo::lev_hyrroe2003_block(pm, qlen, s, slen, cutoff);
}
uint64_t orc_lev_small_band_u8(const uint8_t* q, uint64_t qlen, const uint8_t* s, uint64_t slen, uint64_t cutoff) {
  rfo::BlockPM pm(q, (size_t)qlen);
  return rfo::lev_small_band_with_pm(pm, qlen, s, slen, cutoff);
}

// ---- textbook cross-checks
uint64_t orc_tb_levenshtein_u8(const uint8_t* a, uint64_t n, const uint8_t* b, uint64_t m, uint64_t ins, uint64_t del, uint64_t sub) {
  return rftb::levenshtein(a, n, b, m, ins, del, sub);
}
uint64_t orc_tb_lcs_u8(const uint8_t* a, uint64_t n, const uint8_t* b, uint64_t m) { return rftb::lcs(a, n, b, m); }
uint64_t orc_tb_osa_u8(const uint8_t* a, uint64_t n, const uint8_t* b, uint64_t m) { return rftb::osa(a, n, b, m); }
uint64_t orc_tb_damerau_levenshtein_u8(const uint8_t* a, uint64_t n, const uint8_t* b, uint64_t m) {
  return rftb::damerau_levenshtein(a, n, b, m);
}
double orc_tb_jaro_u8(const uint8_t* a, uint64_t n, const uint8_t* b, uint64_t m) { return rftb::jaro(a, n, b, m); }
double orc_tb_jaro_winkler_u8(const uint8_t* a, uint64_t n, const uint8_t* b, uint64_t m, double w) {
  return rftb::jaro_winkler(a, n, b, m, w);
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
