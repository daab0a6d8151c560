// Host build of rapidfuzz-rs_b200/csrc/rf_core.cuh (the exact arithmetic the CUDA kernels run), exposed
// to pytest so the bit tricks are checked against the oracle on the CPU box before any GPU time is spent.
#include <cstring>
#include <vector>
#if !defined(__CUDACC__)
struct uint2 { unsigned int x, y; };   // CUDA's vector type, for the host build of the row-wise routines
#endif
#include "../rapidfuzz-rs_b200/csrc/rf_core.cuh"

using namespace rfk;

template <class W>
static void build_tab(const uint8_t* q, uint32_t len1, bool top, W* tab) {
  for (int i = 0; i < 256; ++i) tab[i] = 0;
  for (uint32_t i = 0; i < len1; ++i) tab[q[i]] |= (W)1 << i;
  if (top) {
    const int sh = (int)sizeof(W) * 8 - (int)len1;
    for (int i = 0; i < 256; ++i) tab[i] <<= sh;
  }
}

// copies s into a 4B-aligned padded buffer at byte offset `mis` (0..3) to exercise the funnel reader
struct Padded {
  std::vector<uint32_t> buf;
  uint32_t start;
  Padded(const uint8_t* s, uint32_t len, uint32_t mis) : buf((len + mis) / 4 + 4, 0xA5A5A5A5u), start(mis) {
    if (len) memcpy(reinterpret_cast<uint8_t*>(buf.data()) + mis, s, len);
  }
  const uint8_t* base() const { return reinterpret_cast<const uint8_t*>(buf.data()); }
};

extern "C" {

uint32_t core_raw(int family, int bits, const uint8_t* q, uint32_t len1, const uint8_t* s, uint32_t len2, uint32_t mis) {
  Padded p(s, len2, mis);
  ByteReader rd(p.base(), p.start);
  if (bits == 32) {
    uint32_t tab[256];
    build_tab(q, len1, family != F_LCS, tab);
    auto t = [&](uint32_t ch) { return tab[ch]; };
    if (family == F_LEV) return lev_w1<uint32_t>(t, rd, len2, len1);
    if (family == F_OSA) return osa_w1<uint32_t>(t, rd, len2, len1);
    return lcs_w1<uint32_t>(t, rd, len2);
  }
  uint64_t tab[256];
  build_tab(q, len1, family != F_LCS, tab);
  auto t = [&](uint32_t ch) { return tab[ch]; };
  if (family == F_LEV) return lev_w1<uint64_t>(t, rd, len2, len1);
  if (family == F_OSA) return osa_w1<uint64_t>(t, rd, len2, len1);
  return lcs_w1<uint64_t>(t, rd, len2);
}

static Epi make_epi(int metric, int kind, int has_cutoff, uint64_t cu, double cf, uint64_t wi, uint64_t wd, uint64_t ws,
                    double pw, int quirks) {
  Epi e{};
  e.metric = metric; e.kind = kind; e.has_cutoff = has_cutoff; e.cutoff_u = cu; e.cutoff_f = cf;
  e.w_ins = wi; e.w_del = wd; e.w_sub = ws; e.prefix_weight = pw; e.quirks = quirks;
  e.wclass = WC_UNIFORM;
  if (metric == M_LEVENSHTEIN) {
    if (wi == 0 && wd == 0) e.wclass = WC_ZERO;
    else if (wi == wd && wi == ws) e.wclass = WC_UNIFORM;
    else e.wclass = WC_INDEL;
  }
  e.unit32 = ((metric == M_LEVENSHTEIN && e.wclass == WC_UNIFORM && wi == 1) || metric == M_INDEL || metric == M_LCS_SEQ ||
              metric == M_OSA) ? 1 : 0;   // same rule as rf_api.cu make_epi()
  return e;
}

// full per-candidate pipeline for query <= 64: raw kernel + score algebra; out_u (NONE_U32 = None) / out_f (NaN = None)
int core_score(int metric, int kind, const uint8_t* q, uint32_t len1, const uint8_t* s, uint32_t len2, int has_cutoff,
               uint64_t cu, double cf, uint64_t wi, uint64_t wd, uint64_t ws, double pw, int quirks, uint32_t mis,
               uint32_t* out_u, double* out_f) {
  Epi e = make_epi(metric, kind, has_cutoff, cu, cf, wi, wd, ws, pw, quirks);
  Padded p(s, len2, mis);
  const Family fam = family_of(metric, e.wclass);
  if (fam == F_JARO) {
    uint64_t tab[256];
    build_tab(q, len1, false, tab);
    auto t = [&](uint32_t ch) { return tab[ch]; };
    const uint8_t* b = p.base() + p.start;
    auto bytes = [&](uint32_t j) -> uint32_t { return b[j]; };
    auto jaro = [&](double c) { return jaro_similarity_w1(t, bytes, len1, len2, c); };
    uint32_t prefix = 0;
    while (prefix < 4 && prefix < len1 && prefix < len2 && ((t(bytes(prefix)) >> prefix) & 1)) ++prefix;
    auto sim = [&](double c) {
      return metric == M_JARO ? jaro(c) : jaro_winkler_from(jaro, prefix, e.prefix_weight, c);
    };
    *out_f = finish_float(e, sim);
    return 1;
  }
  uint32_t raw;
  if (len1 == 0) raw = (fam == F_LCS) ? 0 : len2;
  else {
    const int bits = len1 <= 32 ? 32 : 64;
    raw = core_raw(fam, bits, q, len1, s, len2, mis);
  }
  if (result_is_float(metric, kind)) { *out_f = finish_norm(e, raw, len1, len2); return 1; }
  *out_u = finish_int(e, raw, len1, len2);
  return 0;
}

// banded Levenshtein (LevBand64): distance if <= k else 0xFFFFFFFF; check_every = early-exit stride (0: only at the end)
uint32_t core_lev_band(const uint8_t* q, uint32_t len1, const uint8_t* s2, uint32_t len2, uint32_t k, uint32_t check_every) {
  const uint32_t diff = len1 > len2 ? len1 - len2 : len2 - len1;
  if (diff > k) return 0xFFFFFFFFu;
  if (len2 == 0) return len1;
  if (len1 == 0) return len2;
  const uint32_t words = (len1 + 63) / 64, stride = (2 * (words + 2)) | 1u;  // u32 units, odd (as rf_api.cu builds it)
  std::vector<uint32_t> pm((size_t)256 * stride, 0);
  for (uint32_t i = 0; i < len1; ++i) pm[(size_t)q[i] * stride + 2 + i / 32] |= 1u << (i % 32);
  LevBand64 b;
  b.init(len1, len2, k);
  // candidate at a random-ish misalignment inside a 16-byte aligned buffer, read with ByteReader16 like the kernel
  const uint32_t mis = (len1 * 7u + len2 * 3u + k) & 15u;
  std::vector<uint32_t> buf((len2 + mis) / 4 + 16, 0xA5A5A5A5u);
  memcpy(reinterpret_cast<uint8_t*>(buf.data()) + mis, s2, len2);
  ByteReader16 rd(reinterpret_cast<const uint8_t*>(buf.data()), mis);
  Bytes16 blk{};
  for (uint32_t j = 0; j < len2; ++j) {
    if ((j & 15u) == 0) blk = rd.next16();
    const uint32_t ch = (blk.w[(j >> 2) & 3u] >> (8 * (j & 3u))) & 0xffu;
    if (ch != s2[j]) return 0xDEADBEEFu;
    const uint32_t sp = (uint32_t)(b.s + 64);
    const uint32_t* row = pm.data() + (size_t)ch * stride + (sp >> 5);
    // two-word window: valid when the band has at most 33 diagonals (k <= 32); odd k keep the three-word form under test
    b.step((k > 32 || (k & 1u)) ? band_window32(row[0], row[1], row[2], sp & 31u) : band_window32_low33(row[0], row[1], sp & 31u));
    if (check_every && (j % check_every) == check_every - 1 && b.score() > (int32_t)k) return 0xFFFFFFFFu;
  }
  const int32_t d = b.score();
  return d <= (int32_t)k ? (uint32_t)d : 0xFFFFFFFFu;
}

// Jaro via the row-wise 32-bit passes (jaro32_rows); returns NaN when the pair is outside that path's domain
double core_jaro32(const uint8_t* q, uint32_t len1, const uint8_t* s2, uint32_t len2, double cutoff, uint32_t extra_rows) {
  if (len1 == 0 || len1 > 32) return NAN;
  uint32_t tab32[256];
  build_tab(q, len1, false, tab32);
  uint32_t l1 = len1, l2 = len2, bound = 0;
  jaro_bounds(l1, l2, bound);
  if (l2 > 64) return NAN;
  std::vector<uint8_t> buf((size_t)len2 + 8 * (extra_rows + 2), 0x5A);  // junk behind the candidate
  if (len2) memcpy(buf.data(), s2, len2);
  auto tab = [&](uint32_t ch) { return tab32[ch]; };
  auto row = [&](uint32_t r) { uint2 v; memcpy(&v, buf.data() + 8 * r, 8); return v; };
  const uint32_t nrows = (l2 + 7) / 8 + extra_rows;
  const Jaro32Result res = jaro32_rows(tab, row, l2, bound, nrows);
  const bool fm = len2 > 0 && (tab32[s2[0]] & 1u);
  static std::vector<double> quot;
  if (quot.empty()) {
    quot.resize(kQuotDim * kQuotDim);
    for (int a = 0; a < kQuotDim; ++a)
      for (int b = 0; b < kQuotDim; ++b) quot[a * kQuotDim + b] = b ? (double)a / (double)b : 0.0;
  }
  const double with_tab = jaro32_finish(len1, len2, res, fm, cutoff, quot.data());
  const double plain = jaro32_finish(len1, len2, res, fm, cutoff);
  if (memcmp(&with_tab, &plain, sizeof(double)) != 0) return -12345.0;  // the table path must be bit-identical
  return with_tab;
}

double core_div3(double x) { return div3_exact(x); }

// Damerau-Levenshtein through the routine the kernel runs: candidate outer, query inner, strided scratch
uint32_t core_dl(const uint8_t* q, uint32_t len1, const uint8_t* s2, uint32_t len2) {
  const size_t T = 3, rowlen = len1 + 2;
  std::vector<int32_t> rows(3 * rowlen * T, 12345), last(256 * T, -1);
  const uint32_t r = damerau_zhao([&](uint32_t i) -> uint32_t { return s2[i]; }, len2, [&](uint32_t j) -> uint32_t { return q[j]; }, len1,
                                  [&](uint32_t k, uint32_t j) -> int32_t& { return rows[(k * rowlen + j) * T]; },
                                  [&](uint32_t ch) -> int32_t& { return last[ch * T]; });
  for (size_t i = 0; i < 256; ++i) if (last[i * T] != -1) return 0xBADBADu;  // the table must be restored
  return r;
}

// generic weighted Levenshtein through the routine the Wagner-Fischer kernel runs (strided row like on the device)
uint64_t core_wf(const uint8_t* q, uint32_t len1, const uint8_t* s2, uint32_t len2, uint64_t wi, uint64_t wd, uint64_t ws) {
  const size_t stride = 3;
  std::vector<uint64_t> row((size_t)(len1 + 1) * stride, 0xDEADBEEFull);
  return weighted_wagner_fischer([&](uint32_t i) -> uint32_t { return q[i]; }, [&](uint32_t j) -> uint32_t { return s2[j]; }, len1,
                                 len2, wi, wd, ws, [&](uint32_t i) -> uint64_t& { return row[(size_t)i * stride]; });
}

// hamming (pad semantics) / prefix / postfix raw values through the word-wise routines the kernel uses
uint32_t core_simple(int which, const uint8_t* q, uint32_t len1, const uint8_t* s2, uint32_t len2) {
  std::vector<uint32_t> qw(len1 / 4 + 8, 0xA5A5A5A5u), tw(len2 / 4 + 8, 0x5A5A5A5Au);   // junk behind the data
  if (len1) memcpy(qw.data(), q, len1);
  if (len2) memcpy(tw.data(), s2, len2);
  auto q4 = [&](uint32_t w) { return qw[w]; };
  auto t4 = [&](uint32_t w) { return tw[w]; };
  if (which == 0) return hamming_raw(q4, t4, len1, len2);
  if (which == 1) return prefix_raw(q4, t4, len1, len2);
  return postfix_raw([&](uint32_t j) -> uint32_t { return q[j]; }, [&](uint32_t j) -> uint32_t { return s2[j]; }, len1, len2);
}

// generic (multi-word) Jaro on the host, query <= 1024
double core_jaro_generic(const uint8_t* q, uint32_t len1, const uint8_t* s, uint32_t len2, double cutoff) {
  const uint32_t words = (len1 + 63) / 64;
  std::vector<uint64_t> pm(256 * (words ? words : 1), 0);
  for (uint32_t i = 0; i < len1; ++i) pm[q[i] * words + i / 64] |= 1ULL << (i % 64);
  auto pmw = [&](uint32_t w, uint32_t ch) -> uint64_t { return pm[ch * words + w]; };
  auto bytes = [&](uint32_t j) -> uint32_t { return s[j]; };
  return jaro_similarity_generic<1024>(pmw, bytes, len1, len2, cutoff);
}

}  // extern "C"
