// rapidfuzz_b200.hpp -- header-only C++17 host mirror of the rapidfuzz-rs API surface for the one-vs-many
// path, over the C ABI of include/rfgpu.h (librfgpu.so).  The reference's host language is Rust, which this
// image cannot compile; C++ is the compiled-language stand-in, and the Rust shim a maintainer would add is in
// INTEGRATION.md / rapidfuzz-rs_b200/rust/src/lib.rs.
//
// Mirrors (reference paths relative to src/):
//   distance::{levenshtein,indel,lcs_seq,osa,jaro,jaro_winkler}::{Args, BatchComparator, distance,
//     similarity, normalized_distance, normalized_similarity}            (e.g. levenshtein.rs:86-126, :1636-1818)
//   fuzz::{ratio, RatioBatchComparator}                                   (fuzz.rs:48-150)
// Differences forced by the batch model: the candidate side is a Corpus (all candidates, uploaded once) and
// every method returns one value per candidate; methods taking a single string are provided for parity with
// the reference's tests.  With a score_cutoff the element type becomes std::optional<T> exactly where the
// reference's return type becomes Option<T> (common.rs:18-86).
#pragma once
#include <cmath>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <type_traits>
#include <vector>

#include "rfgpu.h"

namespace rapidfuzz_b200 {

struct Error : std::runtime_error {
  rf_status status;
  Error(rf_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};
inline void check(rf_status s) {
  if (s != RF_OK) throw Error(s, std::string(rf_status_string(s)) + ": " + rf_last_error());
}

// Packed candidates resident in one GPU's HBM.
// Integer elements of other widths (the reference takes any HashableChar, details/common.rs:29-37, and compares them
// numerically): widened to the ABI's u32 BY VALUE -- unsigned types zero-extended, negative values keep their
// two's-complement 32-bit pattern -- so a signed -1 never meets an unsigned 255 / 65535.  64-bit element types are
// accepted when every value fits [-2^31, 2^32); a sequence must not mix negative values with values >= 2^31.
template <class T>
inline std::vector<uint32_t> widen_elements(const T* elems, size_t count) {
  static_assert(std::is_integral_v<T>, "elements must be integers");
  std::vector<uint32_t> out(count);
  bool neg = false, big = false;
  for (size_t i = 0; i < count; ++i) {
    const T v = elems[i];
    if constexpr (std::is_signed_v<T>) {
      if (v < 0) {
        neg = true;
        if (static_cast<long long>(v) < -(1ll << 31)) throw Error(RF_ERR_UNSUPPORTED, "element below -2^31");
      }
    }
    if constexpr (sizeof(T) > 4) {
      if (v > 0 && static_cast<unsigned long long>(v) >= (1ull << 32)) throw Error(RF_ERR_UNSUPPORTED, "element of 2^32 or more");
    }
    if (v > 0 && static_cast<unsigned long long>(v) >= (1ull << 31)) big = true;
    out[i] = static_cast<uint32_t>(static_cast<long long>(v));
  }
  if (neg && big) throw Error(RF_ERR_UNSUPPORTED, "negative values and values >= 2^31 in one sequence");
  return out;
}

class Corpus {
 public:
  Corpus(const uint8_t* chars, const uint64_t* offsets, uint64_t n, int device = 0) {
    check(rf_corpus_create_u8(chars, offsets, n, device, &h_));
  }
  template <class Strings>
  static Corpus from_strings(const Strings& strings, int device = 0) {
    std::vector<uint8_t> chars;
    std::vector<uint64_t> offsets{0};
    for (const auto& s : strings) {
      chars.insert(chars.end(), std::begin(s), std::end(s));
      offsets.push_back(chars.size());
    }
    return Corpus(chars.data(), offsets.data(), offsets.size() - 1, device);
  }
  // u32 elements (code points); scored by comparators built from std::u32string_view queries
  static Corpus from_u32(const uint32_t* elems, const uint64_t* offsets, uint64_t n, int device = 0) {
    Corpus c;
    check(rf_corpus_create_u32(elems, offsets, n, device, &c.h_));
    return c;
  }
  // integer elements of any other width, widened by value (widen_elements)
  template <class T>
  static Corpus from_elements(const T* elems, const uint64_t* offsets, uint64_t n, int device = 0) {
    const std::vector<uint32_t> w = widen_elements(elems, (size_t)offsets[n]);
    return from_u32(w.data(), offsets, n, device);
  }
  // corpus file written by rf_corpus_file_write (mmap + upload)
  static Corpus from_file(const std::string& path, int device = 0) {
    Corpus c;
    check(rf_corpus_create_from_file(path.c_str(), device, &c.h_));
    return c;
  }
  Corpus(Corpus&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  Corpus(const Corpus&) = delete;
  Corpus& operator=(const Corpus&) = delete;
  ~Corpus() { rf_corpus_destroy(h_); }
  uint64_t size() const { return rf_corpus_size(h_); }
  // frees the CSR copy (45 % of the footprint): afterwards only what the interleaved layout serves works (see rfgpu.h)
  void release_csr() { check(rf_corpus_release_csr(h_)); }
  bool has_csr() const { return rf_corpus_has_csr(h_) != 0; }
  const rf_corpus* handle() const { return h_; }

 private:
  Corpus() = default;
  rf_corpus* h_ = nullptr;
};

// (candidate index, score) of the post-processing entry points
template <class T>
struct Hit {
  uint32_t index;
  T score;
};

struct NoScoreCutoff {};
template <class T>
struct WithScoreCutoff { T value; };

// Args<ResultType, CutoffType> builder (levenshtein.rs:86-126 and the per-metric equivalents)
template <class T, class Cutoff = NoScoreCutoff>
struct Args {
  Cutoff cutoff{};
  std::optional<T> hint{};
  uint64_t ins = 1, del = 1, sub = 1;
  double prefix_weight_ = 0.1;
  bool quirks = false;
  bool pad_ = false;  // hamming::Args::pad (hamming.rs:112-118)
  Args score_hint(T h) const { Args a = *this; a.hint = h; return a; }
  Args<T, WithScoreCutoff<T>> score_cutoff(T c) const {
    Args<T, WithScoreCutoff<T>> a;
    a.cutoff = {c}; a.hint = hint; a.ins = ins; a.del = del; a.sub = sub; a.prefix_weight_ = prefix_weight_; a.quirks = quirks;
    a.pad_ = pad_;
    return a;
  }
  Args weights(uint64_t insertion_cost, uint64_t deletion_cost, uint64_t substitution_cost) const {
    Args a = *this; a.ins = insertion_cost; a.del = deletion_cost; a.sub = substitution_cost; return a;
  }
  Args prefix_weight(double w) const { Args a = *this; a.prefix_weight_ = w; return a; }
  Args reference_quirks(bool on = true) const { Args a = *this; a.quirks = on; return a; }
  Args pad(bool on = true) const { Args a = *this; a.pad_ = on; return a; }
};

namespace detail {
template <class T, class C>
rf_args to_c(const Args<T, C>& a) {
  rf_args r;
  rf_args_default(&r);
  r.insertion_cost = a.ins; r.deletion_cost = a.del; r.substitution_cost = a.sub;
  r.prefix_weight = a.prefix_weight_;
  r.reference_quirks = a.quirks ? 1 : 0;
  r.pad = a.pad_ ? 1 : 0;
  if constexpr (!std::is_same_v<C, NoScoreCutoff>) {
    r.has_cutoff = 1;
    if constexpr (std::is_floating_point_v<T>) r.cutoff_f = (double)a.cutoff.value; else r.cutoff_u = (uint64_t)a.cutoff.value;
  }
  if (a.hint) {
    r.has_hint = 1;
    if constexpr (std::is_floating_point_v<T>) r.hint_f = (double)*a.hint; else r.hint_u = (uint64_t)*a.hint;
  }
  return r;
}
template <class T> struct Raw;
template <> struct Raw<uint32_t> {
  static rf_status extract(const rf_batch* b, const rf_corpus* c, rf_kind k, const rf_args* a, uint32_t kk, uint32_t* i, uint32_t* s, uint32_t* n) { return rf_batch_extract_u32(b, c, k, a, kk, i, s, n); }
  static rf_status filter(const rf_batch* b, const rf_corpus* c, rf_kind k, const rf_args* a, uint64_t cap, uint32_t* i, uint32_t* s, uint64_t* n) { return rf_batch_filter_u32(b, c, k, a, cap, i, s, n); }
  static rf_status stream(const rf_batch* b, const uint8_t* ch, const uint64_t* off, uint64_t n, rf_kind k, const rf_args* a, uint32_t* out) { return rf_batch_stream_u32(b, ch, off, n, k, a, out); }
  static std::vector<uint32_t> run(const rf_batch* b, const Corpus& c, rf_kind k, const rf_args& a) {
    std::vector<uint32_t> out(c.size());
    check(rf_batch_score_u32(b, c.handle(), k, &a, out.data()));
    return out;
  }
  static bool none(uint32_t v) { return v == UINT32_MAX; }
};
template <> struct Raw<double> {
  static rf_status extract(const rf_batch* b, const rf_corpus* c, rf_kind k, const rf_args* a, uint32_t kk, uint32_t* i, double* s, uint32_t* n) { return rf_batch_extract_f64(b, c, k, a, kk, i, s, n); }
  static rf_status filter(const rf_batch* b, const rf_corpus* c, rf_kind k, const rf_args* a, uint64_t cap, uint32_t* i, double* s, uint64_t* n) { return rf_batch_filter_f64(b, c, k, a, cap, i, s, n); }
  static rf_status stream(const rf_batch* b, const uint8_t* ch, const uint64_t* off, uint64_t n, rf_kind k, const rf_args* a, double* out) { return rf_batch_stream_f64(b, ch, off, n, k, a, out); }
  static std::vector<double> run(const rf_batch* b, const Corpus& c, rf_kind k, const rf_args& a) {
    std::vector<double> out(c.size());
    check(rf_batch_score_f64(b, c.handle(), k, &a, out.data()));
    return out;
  }
  static bool none(double v) { return std::isnan(v); }
};
}  // namespace detail

// One metric module.  IntT = element type of distance/similarity (uint32_t for the edit-distance family,
// double for Jaro / Jaro-Winkler).
template <rf_metric M, class IntT>
struct MetricModule {
  using ArgsInt = Args<IntT>;
  using ArgsF64 = Args<double>;

  class BatchComparator {  // BatchComparator::new(query): caches s1 and its pattern-match table on the GPU
   public:
    explicit BatchComparator(std::string_view query, int device = 0) : device_(device) {
      check(rf_batch_create_u8(M, reinterpret_cast<const uint8_t*>(query.data()), (uint32_t)query.size(), device, &h_));
    }
    explicit BatchComparator(std::u32string_view query, int device = 0) : device_(device) {  // char / u32 elements
      check(rf_batch_create_u32(M, reinterpret_cast<const uint32_t*>(query.data()), (uint32_t)query.size(), device, &h_));
    }
    template <class T, class = std::enable_if_t<std::is_integral_v<T> && !std::is_same_v<T, char> && !std::is_same_v<T, char32_t>>>
    BatchComparator(const T* query, size_t len, int device = 0) : device_(device) {  // other integer widths, by value
      const std::vector<uint32_t> w = widen_elements(query, len);
      check(rf_batch_create_u32(M, w.data(), (uint32_t)w.size(), device, &h_));
    }
    BatchComparator(BatchComparator&& o) noexcept : h_(o.h_), device_(o.device_) { o.h_ = nullptr; }
    BatchComparator(const BatchComparator&) = delete;
    ~BatchComparator() { rf_batch_destroy(h_); }

    // one value per candidate, no cutoff: bare values (common.rs:18-31)
    std::vector<IntT> distance(const Corpus& c) const { return score<IntT>(c, RF_DISTANCE, ArgsInt{}); }
    std::vector<IntT> similarity(const Corpus& c) const { return score<IntT>(c, RF_SIMILARITY, ArgsInt{}); }
    std::vector<double> normalized_distance(const Corpus& c) const { return score<double>(c, RF_NORMALIZED_DISTANCE, ArgsF64{}); }
    std::vector<double> normalized_similarity(const Corpus& c) const { return score<double>(c, RF_NORMALIZED_SIMILARITY, ArgsF64{}); }
    // *_with_args: Option-valued when a score_cutoff is set (common.rs:33-46, :73-86)
    template <class C> auto distance_with_args(const Corpus& c, const Args<IntT, C>& a) const { return wrap<IntT, C>(score<IntT>(c, RF_DISTANCE, a)); }
    template <class C> auto similarity_with_args(const Corpus& c, const Args<IntT, C>& a) const { return wrap<IntT, C>(score<IntT>(c, RF_SIMILARITY, a)); }
    template <class C> auto normalized_distance_with_args(const Corpus& c, const Args<double, C>& a) const { return wrap<double, C>(score<double>(c, RF_NORMALIZED_DISTANCE, a)); }
    template <class C> auto normalized_similarity_with_args(const Corpus& c, const Args<double, C>& a) const { return wrap<double, C>(score<double>(c, RF_NORMALIZED_SIMILARITY, a)); }
    // integer scores as bytes (None = 0xFF): a quarter of the result download; throws when a score exceeds 254
    template <class C>
    std::vector<uint8_t> score_u8(const Corpus& c, rf_kind kind, const Args<IntT, C>& a) const {
      std::vector<uint8_t> out(c.size());
      const rf_args ca = detail::to_c(a);
      check(rf_batch_score_u8(h_, c.handle(), kind, &ca, out.data()));
      return out;
    }
    // new on this side: the k best candidates by (score best-first, index), selected on the GPU; every candidate
    // within the score_cutoff in index order; one-shot scoring of host-resident candidates (chunked PCIe pipeline).
    // T = IntT for RF_DISTANCE / RF_SIMILARITY of the edit-distance metrics, double otherwise.
    template <class T, class C>
    std::vector<Hit<T>> extract(const Corpus& c, rf_kind kind, uint32_t k, const Args<T, C>& a) const {
      std::vector<uint32_t> idx(k);
      std::vector<T> sc(k);
      uint32_t m = 0;
      const rf_args ca = detail::to_c(a);
      check(detail::Raw<T>::extract(h_, c.handle(), kind, &ca, k, idx.data(), sc.data(), &m));
      std::vector<Hit<T>> out(m);
      for (uint32_t i = 0; i < m; ++i) out[i] = {idx[i], sc[i]};
      return out;
    }
    template <class T>
    std::vector<Hit<T>> filter(const Corpus& c, rf_kind kind, const Args<T, WithScoreCutoff<T>>& a) const {
      const rf_args ca = detail::to_c(a);
      uint64_t total = 0;
      check(detail::Raw<T>::filter(h_, c.handle(), kind, &ca, 0, nullptr, nullptr, &total));  // count, then fetch
      std::vector<uint32_t> idx(total);
      std::vector<T> sc(total);
      if (total) check(detail::Raw<T>::filter(h_, c.handle(), kind, &ca, total, idx.data(), sc.data(), &total));
      std::vector<Hit<T>> out(idx.size());
      for (size_t i = 0; i < out.size(); ++i) out[i] = {idx[i], sc[i]};
      return out;
    }
    template <class T, class C>
    std::vector<T> stream(const uint8_t* chars, const uint64_t* offsets, uint64_t n, rf_kind kind, const Args<T, C>& a) const {
      std::vector<T> out(n);
      const rf_args ca = detail::to_c(a);
      check(detail::Raw<T>::stream(h_, chars, offsets, n, kind, &ca, out.data()));
      return out;
    }
    // single-candidate forms, as in the reference's signatures
    IntT distance(std::string_view s2) const { return distance(one(s2))[0]; }
    IntT similarity(std::string_view s2) const { return similarity(one(s2))[0]; }
    double normalized_distance(std::string_view s2) const { return normalized_distance(one(s2))[0]; }
    double normalized_similarity(std::string_view s2) const { return normalized_similarity(one(s2))[0]; }
    template <class C> auto distance_with_args(std::string_view s2, const Args<IntT, C>& a) const { return distance_with_args(one(s2), a)[0]; }
    template <class C> auto similarity_with_args(std::string_view s2, const Args<IntT, C>& a) const { return similarity_with_args(one(s2), a)[0]; }
    template <class C> auto normalized_similarity_with_args(std::string_view s2, const Args<double, C>& a) const { return normalized_similarity_with_args(one(s2), a)[0]; }
    template <class C> auto normalized_distance_with_args(std::string_view s2, const Args<double, C>& a) const { return normalized_distance_with_args(one(s2), a)[0]; }

   private:
    Corpus one(std::string_view s) const { return Corpus::from_strings(std::vector<std::string_view>{s}, device_); }
    template <class T, class A>
    std::vector<T> score(const Corpus& c, rf_kind k, const A& a) const {
      return detail::Raw<T>::run(h_, c, k, detail::to_c(a));
    }
    template <class T, class C>
    static auto wrap(std::vector<T> raw) {
      if constexpr (std::is_same_v<C, NoScoreCutoff>) return raw;
      else {
        std::vector<std::optional<T>> out(raw.size());
        for (size_t i = 0; i < raw.size(); ++i)
          if (!detail::Raw<T>::none(raw[i])) out[i] = raw[i];
        return out;
      }
    }
    rf_batch* h_ = nullptr;
    int device_;
  };

  // pairwise free functions of the reference (e.g. levenshtein.rs:1380-1585)
  static IntT distance(std::string_view s1, std::string_view s2) { return BatchComparator(s1).distance(s2); }
  static IntT similarity(std::string_view s1, std::string_view s2) { return BatchComparator(s1).similarity(s2); }
  static double normalized_distance(std::string_view s1, std::string_view s2) { return BatchComparator(s1).normalized_distance(s2); }
  static double normalized_similarity(std::string_view s1, std::string_view s2) { return BatchComparator(s1).normalized_similarity(s2); }
  template <class C> static auto distance_with_args(std::string_view s1, std::string_view s2, const Args<IntT, C>& a) { return BatchComparator(s1).distance_with_args(s2, a); }
};

namespace distance {
using levenshtein = MetricModule<RF_LEVENSHTEIN, uint32_t>;
using indel = MetricModule<RF_INDEL, uint32_t>;
using lcs_seq = MetricModule<RF_LCS_SEQ, uint32_t>;
using osa = MetricModule<RF_OSA, uint32_t>;
using jaro = MetricModule<RF_JARO, double>;
using jaro_winkler = MetricModule<RF_JARO_WINKLER, double>;
using hamming = MetricModule<RF_HAMMING, uint32_t>;  // unequal lengths without Args::pad(): Error (status RF_ERR_INVALID_ARG)
using prefix = MetricModule<RF_PREFIX, uint32_t>;
using postfix = MetricModule<RF_POSTFIX, uint32_t>;
using damerau_levenshtein = MetricModule<RF_DAMERAU_LEVENSHTEIN, uint32_t>;
}  // namespace distance

namespace fuzz {  // fuzz.rs:48-150
class RatioBatchComparator {
 public:
  explicit RatioBatchComparator(std::string_view query, int device = 0) : device_(device) {
    check(rf_batch_create_u8(RF_RATIO, reinterpret_cast<const uint8_t*>(query.data()), (uint32_t)query.size(), device, &h_));
  }
  RatioBatchComparator(const RatioBatchComparator&) = delete;
  ~RatioBatchComparator() { rf_batch_destroy(h_); }
  std::vector<double> similarity(const Corpus& c) const { return detail::Raw<double>::run(h_, c, RF_SIMILARITY, detail::to_c(Args<double>{})); }
  template <class C>
  auto similarity_with_args(const Corpus& c, const Args<double, C>& a) const {
    auto raw = detail::Raw<double>::run(h_, c, RF_SIMILARITY, detail::to_c(a));
    if constexpr (std::is_same_v<C, NoScoreCutoff>) return raw;
    else {
      std::vector<std::optional<double>> out(raw.size());
      for (size_t i = 0; i < raw.size(); ++i)
        if (!std::isnan(raw[i])) out[i] = raw[i];
      return out;
    }
  }
  double similarity(std::string_view s2) const {
    return similarity(Corpus::from_strings(std::vector<std::string_view>{s2}, device_))[0];
  }

 private:
  rf_batch* h_ = nullptr;
  int device_;
};
inline double ratio(std::string_view s1, std::string_view s2) { return RatioBatchComparator(s1).similarity(s2); }
}  // namespace fuzz

namespace process {  // new on this side (the reference has no many-vs-many): rf_cdist_topk_u8
struct TopK {
  uint32_t nq = 0, k = 0;
  std::vector<uint32_t> index, distance;  // [nq][k], UINT32_MAX = fewer than k hits
  std::optional<Hit<uint32_t>> at(uint32_t q, uint32_t i) const {
    const size_t j = (size_t)q * k + i;
    if (index[j] == UINT32_MAX) return std::nullopt;
    return Hit<uint32_t>{index[j], distance[j]};
  }
};
// for every query the k best candidates of `c` by (Levenshtein distance, index); queries of at most 64 bytes
template <class Strings, class C = NoScoreCutoff>
TopK cdist_topk(const Strings& queries, const Corpus& c, uint32_t k, const Args<uint32_t, C>& a = Args<uint32_t>{}) {
  std::vector<uint8_t> chars;
  std::vector<uint64_t> offsets{0};
  for (const auto& q : queries) {
    const std::string_view v(q);
    chars.insert(chars.end(), v.begin(), v.end());
    offsets.push_back(chars.size());
  }
  TopK r;
  r.nq = (uint32_t)(offsets.size() - 1);
  r.k = k;
  r.index.resize((size_t)r.nq * k);
  r.distance.resize((size_t)r.nq * k);
  const rf_args ca = detail::to_c(a);
  check(rf_cdist_topk_u8(chars.data(), offsets.data(), r.nq, c.handle(), &ca, k, r.index.data(), r.distance.data()));
  return r;
}
}  // namespace process

// ONE process, several GPUs (rf_corpus_create_sharded_u8 ...): the corpus split by bytes over `devices`, results identical
// to the single-GPU calls on the whole corpus.  The reference's comparator is Clone + Send + Sync plain data
// (levenshtein.rs:1635-1639); these handles may be shared between threads the same way.
namespace sharded {
class Corpus {
 public:
  template <class Strings>
  Corpus(const Strings& strings, std::vector<int> devices) : devices_(std::move(devices)) {
    std::vector<uint8_t> chars;
    std::vector<uint64_t> offsets{0};
    for (const auto& s : strings) {
      const std::string_view v(s);
      chars.insert(chars.end(), v.begin(), v.end());
      offsets.push_back(chars.size());
    }
    check(rf_corpus_create_sharded_u8(chars.data(), offsets.data(), offsets.size() - 1, devices_.data(), (int)devices_.size(), &h_));
  }
  Corpus(const Corpus&) = delete;
  Corpus& operator=(const Corpus&) = delete;
  ~Corpus() { rf_sharded_corpus_destroy(h_); }
  uint64_t size() const { return rf_sharded_corpus_size(h_); }
  const rf_sharded_corpus* handle() const { return h_; }
  const std::vector<int>& devices() const { return devices_; }

 private:
  rf_sharded_corpus* h_ = nullptr;
  std::vector<int> devices_;
};

// BatchComparator::new(query) of metric M, replicated on every device of the corpus
template <rf_metric M>
class BatchComparator {
 public:
  BatchComparator(std::string_view q, const std::vector<int>& devices) {
    check(rf_sharded_batch_create_u8(M, (const uint8_t*)q.data(), (uint32_t)q.size(), devices.data(), (int)devices.size(), &h_));
  }
  BatchComparator(const BatchComparator&) = delete;
  BatchComparator& operator=(const BatchComparator&) = delete;
  ~BatchComparator() { rf_sharded_batch_destroy(h_); }
  // integer-valued kinds; UINT32_MAX == None
  std::vector<uint32_t> score_u32(const Corpus& c, rf_kind kind, const rf_args* a = nullptr) const {
    std::vector<uint32_t> out(c.size());
    check(rf_sharded_score_u32(h_, c.handle(), kind, a, out.data()));
    return out;
  }
  std::vector<double> score_f64(const Corpus& c, rf_kind kind, const rf_args* a = nullptr) const {
    std::vector<double> out(c.size());
    check(rf_sharded_score_f64(h_, c.handle(), kind, a, out.data()));
    return out;
  }
  std::vector<uint32_t> distance(const Corpus& c) const { return score_u32(c, RF_DISTANCE); }
  // host-resident candidates: every device streams its part of the range over its own PCIe link
  std::vector<uint32_t> stream_u32(const uint8_t* chars, const uint64_t* offsets, uint64_t n, rf_kind kind, const rf_args* a = nullptr) const {
    std::vector<uint32_t> out(n);
    check(rf_sharded_stream_u32(h_, chars, offsets, n, kind, a, out.data()));
    return out;
  }

 private:
  rf_sharded_batch* h_ = nullptr;
};

struct TopK {
  uint32_t nq = 0, k = 0;
  std::vector<uint64_t> index;     // [nq][k] GLOBAL candidate indices, UINT64_MAX = fewer than k hits
  std::vector<uint32_t> distance;  // [nq][k]
};
template <class Strings>
TopK cdist_topk(const Strings& queries, const Corpus& c, uint32_t k, const rf_args* a = nullptr) {
  std::vector<uint8_t> chars;
  std::vector<uint64_t> offsets{0};
  for (const auto& q : queries) {
    const std::string_view v(q);
    chars.insert(chars.end(), v.begin(), v.end());
    offsets.push_back(chars.size());
  }
  TopK r;
  r.nq = (uint32_t)(offsets.size() - 1);
  r.k = k;
  r.index.resize((size_t)r.nq * k);
  r.distance.resize((size_t)r.nq * k);
  check(rf_sharded_cdist_topk_u8(chars.data(), offsets.data(), r.nq, c.handle(), a, k, r.index.data(), r.distance.data()));
  return r;
}
}  // namespace sharded

}  // namespace rapidfuzz_b200
