// rf_synth.cpp -- deterministic synthetic workloads (BASELINE.md section 2 / SURVEY.md section 8d):
// counter-based SplitMix64, the 62 ASCII alphanumerics of the reference's benches
// (rapidfuzz-benches/benches/bench_levenshtein.rs:8-14), candidate lengths uniform in [min_len, max_len],
// and 1/64 of the candidates planted as the query with k ~ U[0,kmax] random edits so that cutoff
// configurations have non-trivial hits.  Host code only (OpenMP); used by bench.py and the tests.
// NOT part of the product: its own tiny library (synth/librfsynth.so), so that the CPU reference arm of bench.py
// never maps librfgpu.so.
#include <stdint.h>
#include <string.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "rfsynth.h"

static const char ALPHA[63] = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789";

static inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t rnd(uint64_t seed, uint64_t i, uint64_t k) { return splitmix64(splitmix64(seed + i) + k); }
static inline uint8_t alpha_of(uint64_t byte) { return (uint8_t)ALPHA[(byte * 62) >> 8]; }

// planted candidate: query with k edits, clamped into [min_len, max_len]; returns the length
static uint32_t planted(uint64_t seed, uint64_t i, const uint8_t* q, uint32_t qlen, uint32_t min_len, uint32_t max_len,
                        uint32_t kmax, uint8_t* out /* may be NULL; capacity max(qlen+kmax, max_len) */,
                        std::vector<uint8_t>& buf) {
  buf.assign(q, q + qlen);
  const uint32_t k = (uint32_t)(rnd(seed, i, 1) % (uint64_t)(kmax + 1));
  for (uint32_t e = 0; e < k; ++e) {
    const uint64_t r = rnd(seed, i, 2 + e);
    const uint32_t op = (uint32_t)(r % 3);
    const uint8_t ch = alpha_of((r >> 8) & 0xff);
    const uint64_t pr = r >> 16;
    if (op == 0 && !buf.empty()) buf[pr % buf.size()] = ch;                       // substitution
    else if (op == 1) buf.insert(buf.begin() + (pr % (buf.size() + 1)), ch);     // insertion
    else if (!buf.empty()) buf.erase(buf.begin() + (pr % buf.size()));           // deletion
  }
  uint32_t len = (uint32_t)buf.size();
  if (len > max_len) { buf.resize(max_len); len = max_len; }
  uint32_t j = 0;
  while (len < min_len) { buf.push_back(alpha_of(rnd(seed, i, 1000 + j++) & 0xff)); ++len; }
  if (out) memcpy(out, buf.data(), len);
  return len;
}

extern "C" {

int rf_synth_query_u8(uint64_t seed, uint32_t len, uint8_t* out) {
  if (len && !out) return 1;
  for (uint32_t j = 0; j < len; ++j) out[j] = alpha_of(rnd(seed ^ 0x51554552ull /* "QUER" */, 0, j) & 0xff);
  return 0;
}

int rf_synth_corpus_u8(uint64_t seed, const uint8_t* query, uint32_t query_len, uint64_t n, uint32_t min_len,
                             uint32_t max_len, uint32_t kmax, uint64_t* offsets, uint8_t* chars, int nthreads) {
  if (!offsets || max_len < min_len || (query_len && !query)) return 1;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  const uint64_t span = (uint64_t)max_len - min_len + 1;
  if (!chars) {
    // pass 1: lengths -> offsets[i+1] = len_i, then prefix sum
    offsets[0] = 0;
#pragma omp parallel num_threads(nthreads)
    {
      std::vector<uint8_t> buf;
#pragma omp for schedule(static)
      for (int64_t i = 0; i < (int64_t)n; ++i) {
        const uint64_t r0 = rnd(seed, (uint64_t)i, 0);
        uint32_t len = min_len + (uint32_t)(r0 % span);
        if (query_len && ((r0 >> 40) & 63) == 0) len = planted(seed, (uint64_t)i, query, query_len, min_len, max_len, kmax, nullptr, buf);
        offsets[i + 1] = len;
      }
    }
    uint64_t acc = 0;
    for (uint64_t i = 0; i < n; ++i) { acc += offsets[i + 1]; offsets[i + 1] = acc; }
    return 0;
  }
  // pass 2: bytes (offsets must come from pass 1)
#pragma omp parallel num_threads(nthreads)
  {
    std::vector<uint8_t> buf;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      const uint64_t r0 = rnd(seed, (uint64_t)i, 0);
      uint8_t* dst = chars + offsets[i];
      const uint32_t len = (uint32_t)(offsets[i + 1] - offsets[i]);
      if (query_len && ((r0 >> 40) & 63) == 0) {
        planted(seed, (uint64_t)i, query, query_len, min_len, max_len, kmax, dst, buf);
      } else {
        for (uint32_t j = 0; j < len; j += 8) {
          uint64_t r = rnd(seed, (uint64_t)i, 100000 + j / 8);
          for (uint32_t t = 0; t < 8 && j + t < len; ++t) { dst[j + t] = alpha_of(r & 0xff); r >>= 8; }
        }
      }
    }
  }
  return 0;
}

}  // extern "C"
