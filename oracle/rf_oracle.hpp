// rf_oracle.hpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// CPU restatement (C++17, scalar u64 arithmetic) of the one-vs-many `BatchComparator` scoring path of
// rapidfuzz-rs 0.5.0 (reference @151f82c).  Each function cites the reference file:line it follows
// (paths relative to /root/reference/src).  The reference is pure Rust and no Rust toolchain exists in
// this image, so it cannot be executed here; parity is pinned instead by the reference's own
// known-answer vectors (tests/golden/*.json, transcribed from the #[test] blocks of the reference) and by
// an independent textbook-DP cross-check (rf_textbook.hpp).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace rfo {

using usize = uint64_t;
constexpr usize USIZE_MAX = UINT64_MAX;  // Rust usize::MAX on a 64-bit target

// ---------------------------------------------------------------- details/intrinsics.rs
inline usize ceil_div(usize a, usize d) { return a / d + (a % d != 0); }       // intrinsics.rs:1-3
inline uint64_t bit_mask_lsb(usize n) {                                         // intrinsics.rs:31-37
  uint64_t m = ~0ULL;
  if (n < 64) m += (1ULL << n);
  return m;
}
inline uint64_t blsi(uint64_t v) { return v & (0 - v); }                        // intrinsics.rs:38-40
inline uint64_t carrying_add(uint64_t a, uint64_t b, bool cin, bool* cout) {    // intrinsics.rs:25-29
  uint64_t s = a + b;
  bool c1 = s < a;
  uint64_t t = s + (uint64_t)cin;
  bool c2 = t < s;
  *cout = c1 | c2;
  return t;
}
inline usize abs_diff(usize a, usize b) { return a > b ? a - b : b - a; }

// ---------------------------------------------------------------- details/pattern_match_vector.rs
// 128-slot open-addressing map with CPython-style probing (pattern_match_vector.rs:20-65).
struct BitvectorHashmap {
  struct Elem { uint64_t key = 0, value = 0; };
  Elem map[128];
  size_t lookup(uint64_t key) const {
    size_t i = (size_t)(key % 128);
    if (map[i].value == 0 || map[i].key == key) return i;
    uint64_t perturb = key;
    for (;;) {
      i = (i * 5 + (size_t)perturb + 1) % 128;
      if (map[i].value == 0 || map[i].key == key) return i;
      perturb >>= 5;
    }
  }
  uint64_t get(uint64_t key) const { return map[lookup(key)].value; }
  uint64_t& get_mut(uint64_t key) {
    size_t i = lookup(key);
    map[i].key = key;
    return map[i].value;
  }
};

// BlockPatternMatchVector (pattern_match_vector.rs:195-321); extended_ascii is the row-major
// BitMatrix [256][block_count] (matrix.rs:32-36).  Elements are treated as Hash::UNSIGNED
// (u8/u16/u32/char/u64 -- details/common.rs:29-37); values > 255 go to the per-block hashmaps.
struct BlockPM {
  size_t block_count = 0;
  std::vector<uint64_t> ascii;
  std::vector<BitvectorHashmap> map_unsigned;

  template <class C>
  BlockPM(const C* s, size_t len) {                       // new :203-211, insert :213-224
    block_count = (size_t)ceil_div(len, 64);
    ascii.assign(256 * block_count, 0);
    uint64_t mask = 1;
    for (size_t i = 0; i < len; ++i) {
      size_t block = i / 64;
      uint64_t v = (uint64_t)s[i];
      if (v <= 255) {
        ascii[v * block_count + block] |= mask;          // insert_mask :260-264
      } else {
        if (map_unsigned.empty()) map_unsigned.resize(block_count);
        map_unsigned[block].get_mut(v) |= mask;          // :265-277
      }
      mask = (mask << 1) | (mask >> 63);                 // rotate_left(1) :222
    }
  }
  template <class C>
  uint64_t get(size_t block, C ch) const {                // :283-316
    uint64_t v = (uint64_t)ch;
    if (v <= 255) return ascii[v * block_count + block];
    return map_unsigned.empty() ? 0 : map_unsigned[block].get(v);
  }
  size_t size() const { return block_count; }
};

// ---------------------------------------------------------------- details/common.rs
template <class C1, class C2>
inline bool seq_eq(const C1* a, usize la, const C2* b, usize lb) {
  if (la != lb) return false;
  for (usize i = 0; i < la; ++i)
    if ((uint64_t)a[i] != (uint64_t)b[i]) return false;
  return true;
}

template <class C1, class C2>
struct Affix { const C1* s1; usize len1; const C2* s2; usize len2; usize prefix_len, suffix_len; };

template <class C1, class C2>
inline Affix<C1, C2> remove_common_affix(const C1* s1, usize len1, const C2* s2, usize len2) {  // common.rs:79-108
  usize suffix = 0;
  while (suffix < len1 && suffix < len2 &&
         (uint64_t)s1[len1 - 1 - suffix] == (uint64_t)s2[len2 - 1 - suffix]) ++suffix;     // :51-62
  usize l1 = len1 - suffix, l2 = len2 - suffix;
  usize prefix = 0;
  while (prefix < l1 && prefix < l2 && (uint64_t)s1[prefix] == (uint64_t)s2[prefix]) ++prefix;  // :39-49
  return {s1 + prefix, l1 - prefix, s2 + prefix, l2 - prefix, prefix, suffix};
}

inline double norm_sim_to_norm_dist(double c) { return std::min(1.0 - c + 0.00001, 1.0); }  // common.rs:4-7

// ---------------------------------------------------------------- distance/levenshtein.rs
struct Weights { usize ins = 1, del = 1, sub = 1; };  // WeightTable :130-148

template <class C1, class C2>
usize generalized_wagner_fischer(const C1* s1, usize len1, const C2* s2, usize len2, const Weights& w) {  // :212-259
  std::vector<usize> cache(len1 + 1);
  for (usize i = 0; i <= len1; ++i) cache[i] = i * w.del;
  for (usize j = 0; j < len2; ++j) {
    usize temp = cache[0];
    cache[0] += w.ins;
    for (usize i = 0; i < len1; ++i) {
      if ((uint64_t)s1[i] != (uint64_t)s2[j]) {
        temp = std::min(cache[i] + w.del, temp + w.sub);
        temp = std::min(temp, cache[i + 1] + w.ins);
      }
      std::swap(cache[i + 1], temp);
    }
  }
  return cache[len1];
}

inline usize lev_maximum(usize len1, usize len2, const Weights& w) {  // _maximum :263-277
  usize max_dist = len1 * w.del + len2 * w.ins;
  if (len1 >= len2) return std::min(max_dist, len2 * w.sub + (len1 - len2) * w.del);
  return std::min(max_dist, len1 * w.sub + (len2 - len1) * w.ins);
}

inline usize lev_min_distance(usize len1, usize len2, const Weights& w) {  // _min_distance :279-284
  int64_t a = ((int64_t)len1 - (int64_t)len2) * (int64_t)w.del;
  int64_t b = ((int64_t)len2 - (int64_t)len1) * (int64_t)w.ins;
  return (usize)std::max(a, b);
}

template <class C1, class C2>
usize generalized_distance(const C1* s1, usize len1, const C2* s2, usize len2, const Weights& w, usize cutoff) {  // :286-309
  if (lev_min_distance(len1, len2, w) > cutoff) return USIZE_MAX;
  auto a = remove_common_affix(s1, len1, s2, len2);
  return generalized_wagner_fischer(a.s1, a.len1, a.s2, a.len2, w);
}

static const uint8_t LEV_MBLEVEN[9][7] = {  // LEVENSHTEIN_MBLEVEN2018_MATRIX :324-337
    {0x03, 0, 0, 0, 0, 0, 0}, {0x01, 0, 0, 0, 0, 0, 0},
    {0x0F, 0x09, 0x06, 0, 0, 0, 0}, {0x0D, 0x07, 0, 0, 0, 0, 0}, {0x05, 0, 0, 0, 0, 0, 0},
    {0x3F, 0x27, 0x2D, 0x39, 0x36, 0x1E, 0x1B}, {0x3D, 0x37, 0x1F, 0x25, 0x19, 0x16, 0},
    {0x35, 0x1D, 0x17, 0, 0, 0, 0}, {0x15, 0, 0, 0, 0, 0, 0}};

template <class C1, class C2>
usize lev_mbleven2018(const C1* s1, usize len1, const C2* s2, usize len2, usize cutoff) {  // :339-427
  if (len1 < len2) return lev_mbleven2018(s2, len2, s1, len1, cutoff);
  usize len_diff = len1 - len2;
  if (cutoff == 1) return (len_diff == 1 || len1 != 1) ? USIZE_MAX : 1;       // :363-369
  usize ops_index = (cutoff + cutoff * cutoff) / 2 + len_diff - 1;
  const uint8_t* possible = LEV_MBLEVEN[ops_index];
  usize dist = cutoff + 1;
  for (int k = 0; k < 7; ++k) {
    uint8_t ops = possible[k];
    if (ops == 0) break;
    usize i1 = 0, i2 = 0, cur = 0;
    // cur1 = s1[i1] (None if i1>=len1), cur2 likewise; "iter.count()" afterwards counts the elements
    // *after* the current one (:381-422).
    for (;;) {
      bool h1 = i1 < len1, h2 = i2 < len2;
      if (h1 && h2) {
        if ((uint64_t)s1[i1] == (uint64_t)s2[i2]) { ++i1; ++i2; }
        else {
          ++cur;
          if (ops == 0) break;
          if (ops & 1) ++i1;
          if (ops & 2) ++i2;
          ops >>= 2;
        }
      } else if (h1) { ++cur; ++i1; }
      else if (h2) { ++cur; ++i2; }
      else break;
    }
    // remaining elements behind the current positions (:422)
    usize rem1 = (i1 < len1) ? len1 - i1 - 1 : 0;
    usize rem2 = (i2 < len2) ? len2 - i2 - 1 : 0;
    cur += rem1 + rem2;
    dist = std::min(dist, cur);
  }
  return dist;
}

// Hyyro 2003, single word (:435-507)
template <class C2>
usize lev_hyrroe2003(const BlockPM& pm, usize len1, const C2* s2, usize len2, usize cutoff) {
  uint64_t vp = ~0ULL, vn = 0;
  usize dist = len1;
  uint64_t mask = 1ULL << (len1 - 1);
  for (usize j = 0; j < len2; ++j) {
    uint64_t x = pm.get(0, s2[j]);
    uint64_t d0 = (((x & vp) + vp) ^ vp) | x | vn;
    uint64_t hp = vn | ~(d0 | vp);
    uint64_t hn = d0 & vp;
    dist += (hp & mask) != 0;
    dist -= (hn & mask) != 0;
    hp = (hp << 1) | 1;
    hn <<= 1;
    vp = hn | ~(d0 | hp);
    vn = hp & d0;
  }
  return dist <= cutoff ? dist : USIZE_MAX;
}

// diagonal band, 64-bit window (:509-617)
template <class C2>
usize lev_small_band_with_pm(const BlockPM& pm, usize len1, const C2* s2, usize len2, usize cutoff) {
  uint64_t vp = ~0ULL << (64 - cutoff - 1), vn = 0;
  size_t words = pm.size();
  usize curr = cutoff;
  const uint64_t diagonal_mask = 1ULL << 63;
  uint64_t horizontal_mask = 1ULL << 62;
  int64_t start_pos = (int64_t)cutoff + 1 - 64;
  usize break_score = (usize)((int64_t)cutoff + (int64_t)len2 - ((int64_t)len1 - (int64_t)cutoff));
  auto fetch = [&](C2 ch) -> uint64_t {
    if (start_pos < 0) return pm.get(0, ch) << (-start_pos);
    size_t word = (size_t)start_pos / 64, pos = (size_t)start_pos % 64;
    uint64_t v = pm.get(word, ch) >> pos;
    if (word + 1 < words && pos != 0) v |= pm.get(word + 1, ch) << (64 - pos);
    return v;
  };
  usize j = 0;
  if (len1 > cutoff) {
    usize n = std::min<usize>(len1 - cutoff, len2);
    for (; j < n; ++j) {
      uint64_t x = fetch(s2[j]);
      uint64_t d0 = (((x & vp) + vp) ^ vp) | x | vn;
      uint64_t hp = vn | ~(d0 | vp);
      uint64_t hn = d0 & vp;
      curr += (d0 & diagonal_mask) == 0;
      if (curr > break_score) return USIZE_MAX;
      vp = hn | ~((d0 >> 1) | hp);
      vn = (d0 >> 1) & hp;
      ++start_pos;
    }
  }
  for (; j < len2; ++j) {
    uint64_t x = fetch(s2[j]);
    uint64_t d0 = (((x & vp) + vp) ^ vp) | x | vn;
    uint64_t hp = vn | ~(d0 | vp);
    uint64_t hn = d0 & vp;
    curr += (hp & horizontal_mask) != 0;
    curr -= (hn & horizontal_mask) != 0;
    horizontal_mask >>= 1;
    if (curr > break_score) return USIZE_MAX;
    vp = hn | ~((d0 >> 1) | hp);
    vn = (d0 >> 1) & hp;
    ++start_pos;
  }
  return curr;
}

// multi-word + Ukkonen band (:769-1019), RECORD_* = 0
template <class C2>
usize lev_hyrroe2003_block(const BlockPM& pm, usize len1, const C2* s2, usize len2, usize cutoff) {
  if (cutoff < abs_diff(len1, len2)) return USIZE_MAX;
  const int64_t word_size = 64;
  const size_t words = pm.size();
  std::vector<uint64_t> VP(words, ~0ULL), VN(words, 0);
  std::vector<usize> scores(words);
  for (size_t x = 0; x < words; ++x) scores[x] = (x + 1) * 64;
  scores[words - 1] = len1;
  const uint64_t last = 1ULL << ((len1 - 1) % 64);

  cutoff = std::min(cutoff, std::max(len1, len2));
  size_t first_block = 0;
  size_t last_block =
      std::min<usize>(words, ceil_div(std::min(cutoff, (cutoff + len1 - len2) / 2) + 1, 64)) - 1;

  for (usize row = 0; row < len2; ++row) {
    bool hp_carry = true, hn_carry = false;
    auto advance_block = [&](size_t word) {          // :838-875
      uint64_t pm_j = pm.get(word, s2[row]);
      uint64_t vn = VN[word], vp = VP[word];
      uint64_t x = pm_j | (uint64_t)hn_carry;
      uint64_t d0 = (((x & vp) + vp) ^ vp) | x | vn;
      uint64_t hp = vn | ~(d0 | vp);
      uint64_t hn = d0 & vp;
      bool hpc = hp_carry, hnc = hn_carry;
      if (word < words - 1) { hp_carry = (hp >> 63) != 0; hn_carry = (hn >> 63) != 0; }
      else { hp_carry = (hp & last) != 0; hn_carry = (hn & last) != 0; }
      hp = (hp << 1) | (uint64_t)hpc;
      hn = (hn << 1) | (uint64_t)hnc;
      VP[word] = hn | ~(d0 | hp);
      VN[word] = hp & d0;
    };
    auto get_row_num = [&](size_t word) -> int64_t {  // :877-883
      return (word + 1 == words) ? (int64_t)len1 - 1 : (int64_t)(word + 1) * word_size - 1;
    };
    for (size_t word = first_block; word <= last_block; ++word) {   // :885-895
      advance_block(word);
      scores[word] += (usize)hp_carry;
      scores[word] -= (usize)hn_carry;
    }
    cutoff = (usize)std::min<int64_t>(                                // :897-904
        (int64_t)cutoff,
        (int64_t)scores[last_block] +
            std::max<int64_t>((int64_t)len2 - (int64_t)row - 1,
                              (int64_t)len1 - ((int64_t)(1 + last_block) * word_size - 1) - 1));
    if (last_block + 1 < words &&                                     // :912-934
        get_row_num(last_block) <= (int64_t)cutoff + 2 * word_size + (int64_t)row + (int64_t)len1 -
                                       (int64_t)scores[last_block] - 2 - (int64_t)len2) {
      ++last_block;
      VP[last_block] = ~0ULL;
      VN[last_block] = 0;
      usize chars_in_block = (last_block + 1 == words) ? ((len1 - 1) % 64 + 1) : 64;
      scores[last_block] = scores[last_block - 1] + chars_in_block - (usize)hp_carry + (usize)hn_carry;
      advance_block(last_block);
      scores[last_block] += (usize)hp_carry;
      scores[last_block] -= (usize)hn_carry;
    }
    // shrink last_block (:936-960).  last_block is usize in the reference; the loop condition
    // `last_block >= first_block` is checked before the decrement could wrap below 0 only when
    // first_block == 0 and no block is in band -- mirrored here with a signed index.
    int64_t lb = (int64_t)last_block;
    bool band_empty = false;
    for (;;) {
      if (lb < (int64_t)first_block) { band_empty = true; break; }
      bool c1 = scores[lb] < cutoff + 64;
      bool c2 = get_row_num((size_t)lb) <= (int64_t)cutoff + 2 * word_size + (int64_t)row + (int64_t)len1 + 1 -
                                               (int64_t)scores[lb] - 2 - (int64_t)len2;
      if (c1 && c2) break;
      --lb;
    }
    if (band_empty) return USIZE_MAX;   // :982-985 (the first_block loop below cannot revive an empty band)
    last_block = (size_t)lb;
    while (first_block <= last_block) {                                // :963-979
      bool c1 = scores[first_block] < cutoff + 64;
      bool c2 = get_row_num(first_block) >= (int64_t)scores[first_block] + (int64_t)len1 + (int64_t)row -
                                                (int64_t)cutoff - (int64_t)len2;
      if (c1 && c2) break;
      ++first_block;
    }
    if (last_block < first_block) return USIZE_MAX;
  }
  usize dist = scores[words - 1];
  return dist <= cutoff ? dist : USIZE_MAX;
}

// dispatcher (:1021-1102)
template <class C1, class C2>
usize lev_uniform_distance_with_pm(const BlockPM& pm, const C1* s1, usize len1, const C2* s2, usize len2,
                                   usize cutoff, usize hint) {
  cutoff = std::min(cutoff, std::max(len1, len2));
  hint = std::max<usize>(hint, 31);
  if (cutoff == 0) return seq_eq(s1, len1, s2, len2) ? 0 : USIZE_MAX;
  if (cutoff < abs_diff(len1, len2)) return USIZE_MAX;
  if (len1 == 0 || len2 == 0) return len1 + len2;
  if (cutoff >= 4) {
    usize full_band = std::min(len1, 2 * cutoff + 1);
    if (len1 <= 64) return lev_hyrroe2003(pm, len1, s2, len2, cutoff);
    else if (full_band <= 64) return lev_small_band_with_pm(pm, len1, s2, len2, cutoff);
    while (hint < cutoff) {
      full_band = std::min(len1, 2 * hint + 1);
      usize score = (full_band <= 64) ? lev_small_band_with_pm(pm, len1, s2, len2, hint)
                                      : lev_hyrroe2003_block(pm, len1, s2, len2, hint);
      if (score <= hint) return score;
      if (USIZE_MAX / 2 < hint) break;
      hint *= 2;
    }
    return lev_hyrroe2003_block(pm, len1, s2, len2, cutoff);
  }
  auto a = remove_common_affix(s1, len1, s2, len2);
  if (a.len1 == 0 || a.len2 == 0) return a.len1 + a.len2;
  return lev_mbleven2018(a.s1, a.len1, a.s2, a.len2, cutoff);
}

// forward decl (indel.rs:287-310)
template <class C1, class C2>
usize indel_distance_with_pm(const BlockPM& pm, const C1* s1, usize len1, const C2* s2, usize len2, usize cutoff);

// weight-class dispatch (:1285-1331).  `dist *= ins` wraps like Rust release arithmetic (quirk Q4).
template <class C1, class C2>
usize lev_distance_with_pm(const BlockPM& pm, const C1* s1, usize len1, const C2* s2, usize len2,
                           const Weights& w, usize cutoff, usize hint) {
  if (w.ins == w.del) {
    if (w.ins == 0) return 0;
    if (w.ins == w.sub) {
      usize d = lev_uniform_distance_with_pm(pm, s1, len1, s2, len2, ceil_div(cutoff, w.ins), ceil_div(hint, w.ins));
      return d * w.ins;
    } else if (w.sub >= w.ins + w.del) {
      usize d = indel_distance_with_pm(pm, s1, len1, s2, len2, ceil_div(cutoff, w.ins));
      return d * w.ins;
    }
  }
  return generalized_distance(s1, len1, s2, len2, w, cutoff);
}

// ---------------------------------------------------------------- distance/lcs_seq.rs
static const uint8_t LCS_MBLEVEN[14][6] = {  // LCS_SEQ_MBLEVEN2018_MATRIX :113-133
    {0, 0, 0, 0, 0, 0}, {0x01, 0, 0, 0, 0, 0},
    {0x09, 0x06, 0, 0, 0, 0}, {0x01, 0, 0, 0, 0, 0}, {0x05, 0, 0, 0, 0, 0},
    {0x09, 0x06, 0, 0, 0, 0}, {0x25, 0x19, 0x16, 0, 0, 0}, {0x05, 0, 0, 0, 0, 0}, {0x15, 0, 0, 0, 0, 0},
    {0x96, 0x66, 0x5A, 0x99, 0x69, 0xA5}, {0x25, 0x19, 0x16, 0, 0, 0}, {0x65, 0x56, 0x95, 0x59, 0, 0},
    {0x15, 0, 0, 0, 0, 0}, {0x55, 0, 0, 0, 0, 0}};

template <class C1, class C2>
usize lcs_mbleven2018(const C1* s1, usize len1, const C2* s2, usize len2, usize cutoff) {  // :135-197
  if (len1 < len2) return lcs_mbleven2018(s2, len2, s1, len1, cutoff);
  usize len_diff = len1 - len2;
  usize max_misses = len1 + len2 - 2 * cutoff;
  usize ops_index = (max_misses + max_misses * max_misses) / 2 + len_diff - 1;
  const uint8_t* possible = LCS_MBLEVEN[ops_index];
  usize max_len = 0;
  for (int k = 0; k < 6; ++k) {
    uint8_t ops = possible[k];
    if (ops == 0) break;
    usize i1 = 0, i2 = 0, cur = 0;
    while (i1 < len1 && i2 < len2) {
      if ((uint64_t)s1[i1] == (uint64_t)s2[i2]) { ++cur; ++i1; ++i2; }
      else {
        if (ops == 0) break;
        if (ops & 1) ++i1;
        else if (ops & 2) ++i2;
        ops >>= 2;
      }
    }
    max_len = std::max(max_len, cur);
  }
  return max_len;
}

// lcs_unroll<N> (:199-261) -- the unrolling is a scheduling detail; words are visited in order.
template <class C2>
usize lcs_unroll(const BlockPM& pm, size_t N, const C2* s2, usize len2, usize cutoff) {
  uint64_t S[8];
  for (size_t i = 0; i < N; ++i) S[i] = ~0ULL;
  for (usize j = 0; j < len2; ++j) {
    bool carry = false;
    for (size_t w = 0; w < N; ++w) {
      uint64_t matches = pm.get(w, s2[j]);
      uint64_t u = S[w] & matches;
      uint64_t x = carrying_add(S[w], u, carry, &carry);
      S[w] = x | (S[w] - u);
    }
  }
  usize sim = 0;
  for (size_t i = 0; i < N; ++i) sim += (usize)__builtin_popcountll(~S[i]);
  return sim >= cutoff ? sim : 0;
}

template <class C2>
usize lcs_blockwise(const BlockPM& pm, usize len1, const C2* s2, usize len2, usize cutoff) {  // :267-341
  size_t words = pm.size();
  std::vector<uint64_t> S(words, ~0ULL);
  usize band_left = len1 - cutoff, band_right = len2 - cutoff;
  size_t first_block = 0;
  size_t last_block = (size_t)std::min<usize>(words, ceil_div(band_left + 1, 64));
  for (usize row = 0; row < len2; ++row) {
    bool carry = false;
    for (size_t w = first_block; w < last_block; ++w) {
      uint64_t matches = pm.get(w, s2[row]);
      uint64_t u = S[w] & matches;
      uint64_t x = carrying_add(S[w], u, carry, &carry);
      S[w] = x | (S[w] - u);
    }
    if (row > band_right) first_block = (size_t)((row - band_right) / 64);
    if (row + 1 + band_left <= len1) last_block = (size_t)ceil_div(row + 1 + band_left, 64);
  }
  usize sim = 0;
  for (size_t i = 0; i < words; ++i) sim += (usize)__builtin_popcountll(~S[i]);
  return sim >= cutoff ? sim : 0;
}

template <class C2>
usize lcs_with_pm(const BlockPM& pm, usize len1, const C2* s2, usize len2, usize cutoff) {  // :343-409
  size_t words = pm.size();
  usize band_left = len1 - cutoff, band_right = len2 - cutoff;
  usize full_band = band_left + 1 + band_right;
  usize full_band_words = std::min<usize>(words, full_band / 64 + 2);
  if (full_band_words < words) return lcs_blockwise(pm, len1, s2, len2, cutoff);
  usize n = ceil_div(len1, 64);
  if (n == 0) return 0;
  if (n <= 8) return lcs_unroll(pm, (size_t)n, s2, len2, cutoff);
  return lcs_blockwise(pm, len1, s2, len2, cutoff);
}

template <class C1, class C2>
usize lcs_similarity_with_pm(const BlockPM& pm, const C1* s1, usize len1, const C2* s2, usize len2, usize cutoff) {  // :439-486
  if (cutoff > len1 || cutoff > len2) return 0;
  usize max_misses = len1 + len2 - 2 * cutoff;
  if (max_misses == 0 || (max_misses == 1 && len1 == len2)) return seq_eq(s1, len1, s2, len2) ? len1 : 0;
  if (max_misses < abs_diff(len1, len2)) return 0;
  if (max_misses >= 5) return lcs_with_pm(pm, len1, s2, len2, cutoff);
  auto a = remove_common_affix(s1, len1, s2, len2);
  usize sim = a.prefix_len + a.suffix_len;
  if (a.len1 != 0 && a.len2 != 0) {
    usize adj = cutoff >= sim ? cutoff - sim : 0;
    sim += lcs_mbleven2018(a.s1, a.len1, a.s2, a.len2, adj);
  }
  return sim;
}

// ---------------------------------------------------------------- distance/indel.rs
template <class C1, class C2>
usize indel_distance_with_pm(const BlockPM& pm, const C1* s1, usize len1, const C2* s2, usize len2, usize cutoff) {  // :287-310
  usize maximum = len1 + len2;
  usize lcs_cutoff = (maximum / 2 >= cutoff) ? maximum / 2 - cutoff : 0;
  usize lcs = lcs_similarity_with_pm(pm, s1, len1, s2, len2, lcs_cutoff);
  return maximum - 2 * lcs;
}

// ---------------------------------------------------------------- distance/osa.rs
template <class C2>
usize osa_hyrroe2003(const BlockPM& pm, usize len1, const C2* s2, usize len2) {  // :84-135
  uint64_t vp = ~0ULL, vn = 0, d0 = 0, pm_j_old = 0;
  usize curr = len1;
  uint64_t mask = 1ULL << (len1 - 1);
  for (usize j = 0; j < len2; ++j) {
    uint64_t pm_j = pm.get(0, s2[j]);
    uint64_t tr = (((~d0) & pm_j) << 1) & pm_j_old;
    d0 = (((pm_j & vp) + vp) ^ vp) | pm_j | vn;
    d0 |= tr;
    uint64_t hp = vn | ~(d0 | vp);
    uint64_t hn = d0 & vp;
    curr += (hp & mask) != 0;
    curr -= (hn & mask) != 0;
    hp = (hp << 1) | 1;
    hn <<= 1;
    vp = hn | ~(d0 | hp);
    vn = hp & d0;
    pm_j_old = pm_j;
  }
  return curr;
}

template <class C2>
usize osa_hyrroe2003_block(const BlockPM& pm, usize len1, const C2* s2, usize len2) {  // :156-227
  struct Row { uint64_t vp = ~0ULL, vn = 0, d0 = 0, pm = 0; };
  size_t words = pm.size();
  uint64_t last = 1ULL << ((len1 - 1) % 64);
  usize curr = len1;
  std::vector<Row> old_v(words + 1), new_v(words + 1);
  for (usize j = 0; j < len2; ++j) {
    uint64_t hp_carry = 1, hn_carry = 0;
    for (size_t word = 0; word < words; ++word) {
      uint64_t vn = old_v[word + 1].vn, vp = old_v[word + 1].vp, d0 = old_v[word + 1].d0;
      uint64_t d0_last = old_v[word].d0;
      uint64_t pm_j_old = old_v[word + 1].pm;
      uint64_t pm_last = new_v[word].pm;
      uint64_t pm_j = pm.get(word, s2[j]);
      uint64_t x = pm_j;
      uint64_t tr = ((((~d0) & x) << 1) | (((~d0_last) & pm_last) >> 63)) & pm_j_old;
      x |= hn_carry;
      d0 = (((x & vp) + vp) ^ vp) | x | vn | tr;
      uint64_t hp = vn | ~(d0 | vp);
      uint64_t hn = d0 & vp;
      if (word == words - 1) {
        curr += (hp & last) != 0;
        curr -= (hn & last) != 0;
      }
      uint64_t hpc = hp_carry; hp_carry = hp >> 63; hp = (hp << 1) | hpc;
      uint64_t hnc = hn_carry; hn_carry = hn >> 63; hn = (hn << 1) | hnc;
      new_v[word + 1].vp = hn | ~(d0 | hp);
      new_v[word + 1].vn = hp & d0;
      new_v[word + 1].d0 = d0;
      new_v[word + 1].pm = pm_j;
    }
    std::swap(new_v, old_v);
  }
  return curr;
}

template <class C2>
usize osa_batch_distance(const BlockPM& pm, usize len1, const C2* s2, usize len2) {  // osa.rs:431-461
  if (len1 == 0) return len2;
  if (len2 == 0) return len1;
  if (len1 <= 64) return osa_hyrroe2003(pm, len1, s2, len2);
  return osa_hyrroe2003_block(pm, len1, s2, len2);
}

// ---------------------------------------------------------------- distance/jaro.rs
inline double jaro_calculate_similarity(usize p_len, usize t_len, usize cc, usize transpositions) {  // :106-119
  transpositions /= 2;
  double sim = 0.0;
  sim += (double)cc / (double)p_len;
  sim += (double)cc / (double)t_len;
  sim += ((double)cc - (double)transpositions) / (double)cc;
  return sim / 3.0;
}
inline bool jaro_length_filter(usize p_len, usize t_len, double cutoff) {  // :122-131
  if (t_len == 0 || p_len == 0) return false;
  double min_len = (double)std::min(p_len, t_len);
  double sim = min_len / (double)p_len + min_len / (double)t_len + 1.0;
  sim /= 3.0;
  return sim >= cutoff;
}
inline bool jaro_common_char_filter(usize p_len, usize t_len, usize cc, double cutoff) {  // :134-145
  if (cc == 0) return false;
  double sim = 0.0;
  sim += (double)cc / (double)p_len;
  sim += (double)cc / (double)t_len;
  sim += 1.0;
  sim /= 3.0;
  return sim >= cutoff;
}

struct FlaggedWord { uint64_t p_flag = 0, t_flag = 0; };

template <class C2>
FlaggedWord jaro_flag_word(const BlockPM& pm, const C2* s2, usize len2, usize bound) {  // :147-190
  FlaggedWord f;
  uint64_t bound_mask = bit_mask_lsb(bound + 1);
  usize j = 0;
  usize n = std::min(bound, len2);
  for (; j < n; ++j) {
    uint64_t pm_j = pm.get(0, s2[j]) & bound_mask & ~f.p_flag;
    f.p_flag |= blsi(pm_j);
    f.t_flag |= (uint64_t)(pm_j != 0) << j;
    bound_mask = (bound_mask << 1) | 1;
  }
  for (; j < len2; ++j) {
    uint64_t pm_j = pm.get(0, s2[j]) & bound_mask & ~f.p_flag;
    f.p_flag |= blsi(pm_j);
    f.t_flag |= (uint64_t)(pm_j != 0) << j;
    bound_mask <<= 1;
  }
  return f;
}

struct FlaggedMulti { std::vector<uint64_t> p_flag, t_flag; };
struct SearchBoundMask { size_t words, empty_words; uint64_t last_mask, first_mask; };

template <class C2>
void jaro_flag_step(const BlockPM& pm, C2 t_j, FlaggedMulti& f, usize j, const SearchBoundMask& bm) {  // :192-284
  size_t j_word = (size_t)(j / 64), j_pos = (size_t)(j % 64);
  size_t word = bm.empty_words;
  size_t last_word = word + bm.words;
  if (bm.words == 1) {
    uint64_t pm_j = pm.get(word, t_j) & bm.last_mask & bm.first_mask & ~f.p_flag[word];
    f.p_flag[word] |= blsi(pm_j);
    f.t_flag[j_word] |= (uint64_t)(pm_j != 0) << j_pos;
    return;
  }
  if (bm.first_mask != 0) {
    uint64_t pm_j = pm.get(word, t_j) & bm.first_mask & ~f.p_flag[word];
    if (pm_j != 0) {
      f.p_flag[word] |= blsi(pm_j);
      f.t_flag[j_word] |= 1ULL << j_pos;
      return;
    }
    ++word;
  }
  // (:229-265 is a 4x unrolled copy of the loop below; identical semantics)
  while (word + 1 < last_word) {
    uint64_t pm_j = pm.get(word, t_j) & ~f.p_flag[word];
    if (pm_j != 0) {
      f.p_flag[word] |= blsi(pm_j);
      f.t_flag[j_word] |= 1ULL << j_pos;
      return;
    }
    ++word;
  }
  if (bm.last_mask != 0) {
    uint64_t pm_j = pm.get(word, t_j) & bm.last_mask & ~f.p_flag[word];
    f.p_flag[word] |= blsi(pm_j);
    f.t_flag[j_word] |= (uint64_t)(pm_j != 0) << j_pos;
  }
}

template <class C2>
FlaggedMulti jaro_flag_block(const BlockPM& pm, usize len1, const C2* s2, usize len2, usize bound) {  // :286-337
  FlaggedMulti f;
  f.p_flag.assign((size_t)ceil_div(len1, 64), 0);
  f.t_flag.assign((size_t)ceil_div(len2, 64), 0);
  usize start_range = std::min(bound + 1, len1);
  SearchBoundMask bm{(size_t)(1 + start_range / 64), 0, (1ULL << (start_range % 64)) - 1, ~0ULL};
  for (usize j = 0; j < len2; ++j) {
    jaro_flag_step(pm, s2[j], f, j, bm);
    if (j + bound + 1 < len1) {
      bm.last_mask = (bm.last_mask << 1) | 1;
      if (j + bound + 2 < len1 && bm.last_mask == ~0ULL) {
        bm.last_mask = 0;
        bm.words += 1;
      }
    }
    if (j >= bound) {
      bm.first_mask <<= 1;
      if (bm.first_mask == 0) {
        bm.first_mask = ~0ULL;
        bm.words -= 1;
        bm.empty_words += 1;
      }
    }
  }
  return f;
}

template <class C2>
usize jaro_count_transpositions_word(const BlockPM& pm, const C2* s2, const FlaggedWord& f) {  // :339-368
  uint64_t p_flag = f.p_flag, t_flag = f.t_flag;
  usize transpositions = 0, pos = 0;
  while (t_flag != 0) {
    uint64_t pmask = blsi(p_flag);
    usize idx = (usize)__builtin_ctzll(t_flag);
    C2 ch = s2[pos + idx];      // s2.nth(idx) consumes idx+1 elements
    pos += idx + 1;
    transpositions += (pm.get(0, ch) & pmask) == 0;
    t_flag = (t_flag >> 1) >> idx;
    p_flag ^= pmask;
  }
  return transpositions;
}

template <class C2>
usize jaro_count_transpositions_block(const BlockPM& pm, const C2* s2, const FlaggedMulti& f, usize flagged_chars) {  // :370-420
  size_t text_word = 0, pattern_word = 0;
  uint64_t t_flag = f.t_flag[text_word], p_flag = f.p_flag[pattern_word];
  usize transpositions = 0, s2_pos = 0, it = 0;  // `it` = elements consumed from the s2 iterator
  while (flagged_chars != 0) {
    while (t_flag == 0) {
      ++text_word;
      if (s2_pos < 64) it += 64 - s2_pos;      // s2.nth(64-1-s2_pos) consumes 64-s2_pos elements
      t_flag = f.t_flag[text_word];
      s2_pos = 0;
    }
    while (t_flag != 0) {
      while (p_flag == 0) { ++pattern_word; p_flag = f.p_flag[pattern_word]; }
      uint64_t pmask = blsi(p_flag);
      usize idx = (usize)__builtin_ctzll(t_flag);
      C2 ch = s2[it + idx];
      it += idx + 1;
      s2_pos += idx + 1;
      transpositions += (pm.get(pattern_word, ch) & pmask) == 0;
      t_flag = (t_flag >> 1) >> idx;
      p_flag ^= pmask;
      --flagged_chars;
    }
  }
  return transpositions;
}

template <class C1, class C2>
double jaro_similarity_with_pm(const BlockPM& pm, const C1* s1, usize len1, const C2* s2, usize len2, double cutoff) {  // :516-598
  usize len1_orig = len1, len2_orig = len2;
  if (cutoff > 1.0) return 0.0;
  if (len1_orig == 0 && len2_orig == 0) return 1.0;
  if (!jaro_length_filter(len1_orig, len2_orig, cutoff)) return 0.0;
  if (len1_orig == 1 && len2_orig == 1) return ((uint64_t)s1[0] == (uint64_t)s2[0]) ? 1.0 : 0.0;
  usize bound;
  if (len2 > len1) {
    bound = len2 / 2 - 1;
    if (len2 > len1 + bound) len2 = len1 + bound;
  } else {
    bound = len1 / 2 - 1;
    if (len1 > len2 + bound) len1 = len2 + bound;
  }
  usize cc = 0, transpositions = 0;
  if (len1 == 0 || len2 == 0) {
  } else if (len1 <= 64 && len2 <= 64) {
    FlaggedWord f = jaro_flag_word(pm, s2, len2, bound);
    cc += (usize)__builtin_popcountll(f.p_flag);
    if (!jaro_common_char_filter(len1_orig, len2_orig, cc, cutoff)) return 0.0;
    transpositions = jaro_count_transpositions_word(pm, s2, f);
  } else {
    FlaggedMulti f = jaro_flag_block(pm, len1, s2, len2, bound);
    usize flagged = 0;                                            // count_common_chars :78-96
    if (f.p_flag.size() < f.t_flag.size()) for (auto x : f.p_flag) flagged += (usize)__builtin_popcountll(x);
    else for (auto x : f.t_flag) flagged += (usize)__builtin_popcountll(x);
    cc += flagged;
    if (!jaro_common_char_filter(len1_orig, len2_orig, cc, cutoff)) return 0.0;
    transpositions = jaro_count_transpositions_block(pm, s2, f, flagged);
  }
  return jaro_calculate_similarity(len1_orig, len2_orig, cc, transpositions);
}

// ---------------------------------------------------------------- distance/jaro_winkler.rs
template <class C1, class C2>
double jw_similarity_with_pm(const BlockPM& pm, const C1* s1, usize len1, const C2* s2, usize len2,
                             double prefix_weight, double cutoff) {  // :103-141
  usize prefix = 0;
  while (prefix < 4 && prefix < len1 && prefix < len2 && (uint64_t)s1[prefix] == (uint64_t)s2[prefix]) ++prefix;
  double jaro_cutoff = cutoff;
  if (jaro_cutoff > 0.7) {
    double prefix_sim = (double)prefix * prefix_weight;
    jaro_cutoff = (prefix_sim >= 1.0) ? 0.7 : std::max(0.7, (prefix_sim - jaro_cutoff) / (prefix_sim - 1.0));
  }
  double sim = jaro_similarity_with_pm(pm, s1, len1, s2, len2, jaro_cutoff);
  if (sim > 0.7) sim += (double)prefix * prefix_weight * (1.0 - sim);
  return sim;
}

// ---------------------------------------------------------------- score algebra (details/distance.rs)
enum Metric : int { LEVENSHTEIN = 0, INDEL = 1, LCS_SEQ = 2, OSA = 3, JARO = 4, JARO_WINKLER = 5, RATIO = 6,
                    HAMMING = 7, PREFIX = 8, POSTFIX = 9, DAMERAU_LEVENSHTEIN = 10 };
enum Kind : int { DISTANCE = 0, SIMILARITY = 1, NORM_DISTANCE = 2, NORM_SIMILARITY = 3 };

struct Args {
  bool has_cutoff = false;
  usize cutoff_u = 0;
  double cutoff_f = 0.0;
  bool has_hint = false;
  usize hint_u = 0;
  double hint_f = 0.0;
  Weights weights;
  double prefix_weight = 0.1;
  bool reference_quirks = false;   // Q1: literal RatioBatchComparator normalisation (fuzz.rs:141)
  bool pad = false;                // hamming.rs:112-118: unequal lengths count as mismatches instead of being an error
};

// ---------------------------------------------------------------- distance/damerau_levenshtein.rs
// distance_zhao (:111-168): Zhao & Sahni's linear-space Damerau-Levenshtein.  The reference's HybridGrowingHashmap
// (last row in which a character of s1 was seen, default -1) is a std::map here.
template <class C1, class C2>
usize dl_distance_zhao(const C1* s1, usize len1, const C2* s2, usize len2) {
  using isz = long long;
  const isz max_val = (isz)std::max(len1, len2) + 1;
  std::map<uint64_t, isz> last_row_id;
  auto get_row = [&](uint64_t ch) -> isz { auto it = last_row_id.find(ch); return it == last_row_id.end() ? -1 : it->second; };
  const usize size = len2 + 2;
  std::vector<isz> fr(size, max_val), r1(size, max_val), r(size);
  r[0] = max_val;
  for (usize j = 1; j < size; ++j) r[j] = (isz)j - 1;
  for (usize i = 1; i <= len1; ++i) {
    const uint64_t ch1 = (uint64_t)s1[i - 1];
    std::swap(r, r1);
    isz last_col_id = -1;
    isz last_i2l1 = r[1];
    r[1] = (isz)i;
    isz t = max_val;
    for (usize j = 1; j <= len2; ++j) {
      const uint64_t ch2 = (uint64_t)s2[j - 1];
      const isz diag = r1[j] + (ch1 != ch2 ? 1 : 0);
      const isz left = r[j] + 1;
      const isz up = r1[j + 1] + 1;
      isz temp = std::min(diag, std::min(left, up));
      if (ch1 == ch2) {
        last_col_id = (isz)j;
        fr[j + 1] = r1[j - 1];
        t = last_i2l1;
      } else {
        const isz k = get_row(ch2);
        const isz l = last_col_id;
        if ((isz)j - l == 1) temp = std::min(temp, fr[j + 1] + ((isz)i - k));
        else if ((isz)i - k == 1) temp = std::min(temp, t + ((isz)j - l));
      }
      last_i2l1 = r[j + 1];
      r[j + 1] = temp;
    }
    last_row_id[ch1] = (isz)i;
  }
  return (usize)r[len2 + 1];
}
template <class C1, class C2>
usize dl_distance_impl(const C1* s1, usize len1, const C2* s2, usize len2, usize score_cutoff) {  // :170-189
  const usize diff = len1 > len2 ? len1 - len2 : len2 - len1;
  if (score_cutoff < diff) return USIZE_MAX;
  auto a = remove_common_affix(s1, len1, s2, len2);
  return dl_distance_zhao(a.s1, a.len1, a.s2, a.len2);
}

// ---------------------------------------------------------------- distance/hamming.rs, prefix.rs, postfix.rs
template <class C1, class C2>
usize hamming_distance_impl(const C1* s1, usize len1, const C2* s2, usize len2) {  // hamming.rs:136-161
  usize dist = 0, i = 0;
  for (;; ++i) {
    const bool a = i < len1, b = i < len2;
    if (a && b) { if (!(s1[i] == s2[i])) ++dist; }
    else if (!a && !b) return dist;
    else ++dist;
  }
}
template <class C1, class C2>
usize find_common_prefix(const C1* s1, usize len1, const C2* s2, usize len2) {      // details/common.rs:39-49
  usize n = 0;
  while (n < len1 && n < len2 && s1[n] == s2[n]) ++n;
  return n;
}
template <class C1, class C2>
usize find_common_suffix(const C1* s1, usize len1, const C2* s2, usize len2) {      // details/common.rs:51-62
  usize n = 0;
  while (n < len1 && n < len2 && s1[len1 - 1 - n] == s2[len2 - 1 - n]) ++n;
  return n;
}

struct OptU { bool some; usize v; };
struct OptF { bool some; double v; };

// One cached query == the reference's BatchComparator (levenshtein.rs:1636-1657 etc.).
template <class C1>
struct Batch {
  Metric metric;
  std::vector<C1> s1;
  BlockPM pm;
  Batch(Metric m, const C1* q, size_t n) : metric(m), s1(q, q + n), pm(q, n) {}

  // ---- integer metrics: MetricUsize (details/distance.rs:154-275)
  usize maximum(usize len1, usize len2, const Args& a) const {
    switch (metric) {
      case LEVENSHTEIN: return lev_maximum(len1, len2, a.weights);     // levenshtein.rs:1593
      case INDEL: case RATIO: return len1 + len2;                      // indel.rs:331
      default: return std::max(len1, len2);                            // lcs_seq.rs:773, osa.rs:432
    }
  }
  template <class C2>
  usize u_distance(const C2* s2, usize len2, bool has_c, usize c, bool has_h, usize h, const Args& a) const {
    const C1* p = s1.data(); usize len1 = s1.size();
    switch (metric) {
      case LEVENSHTEIN:                                                // levenshtein.rs:1597-1622
        return lev_distance_with_pm(pm, p, len1, s2, len2, a.weights, has_c ? c : USIZE_MAX, has_h ? h : USIZE_MAX);
      case INDEL: case RATIO: {                                        // indel.rs:335-368
        usize cutoff = has_c ? c : USIZE_MAX;
        usize maximum = len1 + len2;
        usize lcs_cutoff = (maximum / 2 >= cutoff) ? maximum / 2 - cutoff : 0;
        usize lcs = lcs_similarity_with_pm(pm, p, len1, s2, len2, lcs_cutoff);
        return maximum - 2 * lcs;
      }
      case OSA: return osa_batch_distance(pm, len1, s2, len2);        // osa.rs:435-460
      case HAMMING: return hamming_distance_impl(p, len1, s2, len2);  // hamming.rs:168-186 (cutoff and hint unused)
      case DAMERAU_LEVENSHTEIN: return dl_distance_impl(p, len1, s2, len2, has_c ? c : USIZE_MAX);  // damerau_levenshtein.rs:198-214
      default: {                                                       // default _distance :157-179
        usize maximum = std::max(len1, len2);
        bool hc = has_c; usize cs = hc ? (maximum >= c ? maximum - c : 0) : 0;
        bool hh = has_h; usize hs = hh ? (maximum >= h ? maximum - h : 0) : 0;
        usize sim = u_similarity(s2, len2, hc, cs, hh, hs, a);
        return maximum - sim;
      }
    }
  }
  template <class C2>
  usize u_similarity(const C2* s2, usize len2, bool has_c, usize c, bool has_h, usize h, const Args& a) const {
    const C1* p = s1.data(); usize len1 = s1.size();
    if (metric == LCS_SEQ) return lcs_similarity_with_pm(pm, p, len1, s2, len2, has_c ? c : 0);  // lcs_seq.rs:777-793
    if (metric == PREFIX) return find_common_prefix(p, len1, s2, len2);    // prefix.rs:47-71
    if (metric == POSTFIX) return find_common_suffix(p, len1, s2, len2);   // postfix.rs:47-71
    usize maximum = this->maximum(len1, len2, a);                      // default _similarity :181-211
    if (has_c) {
      if (c > maximum) return maximum;
      if (has_h) h = std::min(h, c);
    }
    usize dist = u_distance(s2, len2, has_c, has_c ? maximum - c : 0, has_h, has_h ? maximum - h : 0, a);
    return maximum - dist;   // wraps when dist == USIZE_MAX (quirk Q2)
  }
  template <class C2>
  double u_norm_distance(const C2* s2, usize len2, bool has_c, double c, bool has_h, double h, const Args& a) const {  // :213-252
    usize len1 = s1.size();
    usize maximum = this->maximum(len1, len2, a);
    usize cd = 0, hd = 0;
    if (has_c) { double cc = std::clamp(c, 0.0, 1.0); cd = (usize)std::ceil((double)maximum * cc); }
    if (has_h) { double hh = std::clamp(h, 0.0, 1.0); hd = (usize)std::ceil((double)maximum * hh); }
    usize dist = u_distance(s2, len2, has_c, cd, has_h, hd, a);
    return maximum == 0 ? 0.0 : (double)dist / (double)maximum;
  }
  template <class C2>
  double u_norm_similarity(const C2* s2, usize len2, bool has_c, double c, bool has_h, double h, const Args& a) const {  // :254-274
    double cs = has_c ? norm_sim_to_norm_dist(c) : 0.0;
    double hs = has_h ? norm_sim_to_norm_dist(h) : 0.0;
    return 1.0 - u_norm_distance(s2, len2, has_c, cs, has_h, hs, a);
  }

  // ---- float metrics: Metricf64 (details/distance.rs:277-385), maximum == 1.0
  template <class C2>
  double f_similarity(const C2* s2, usize len2, bool has_c, double c, const Args& a) const {
    const C1* p = s1.data(); usize len1 = s1.size();
    double cut = has_c ? c : 0.0;
    if (metric == JARO) return jaro_similarity_with_pm(pm, p, len1, s2, len2, cut);          // jaro.rs:807-823
    return jw_similarity_with_pm(pm, p, len1, s2, len2, a.prefix_weight, cut);              // jaro_winkler.rs:375-399
  }
  template <class C2>
  double f_distance(const C2* s2, usize len2, bool has_c, double c, const Args& a) const {   // :280-302
    double maximum = 1.0;
    double cs = has_c ? (maximum >= c ? maximum - c : 0.0) : 0.0;
    double sim = f_similarity(s2, len2, has_c, cs, a);
    return maximum - sim;
  }
  template <class C2>
  double f_norm_distance(const C2* s2, usize len2, bool has_c, double c, const Args& a) const {  // :336-362
    double maximum = 1.0;
    double dist = f_distance(s2, len2, has_c, has_c ? maximum * c : 0.0, a);
    return maximum > 0.0 ? dist / maximum : 0.0;
  }
  template <class C2>
  double f_norm_similarity(const C2* s2, usize len2, bool has_c, double c, const Args& a) const {  // :364-384
    double cs = has_c ? norm_sim_to_norm_dist(c) : 0.0;
    return 1.0 - f_norm_distance(s2, len2, has_c, cs, a);
  }

  bool is_float_metric() const { return metric == JARO || metric == JARO_WINKLER; }

  // Public wrappers == BatchComparator::{distance,similarity,normalized_*}_with_args incl. the final
  // `score()` filter (common.rs:43-45, :83-85).  Integer-valued results.
  template <class C2>
  OptU int_score(Kind k, const C2* s2, usize len2, const Args& a) const {
    if (k == DISTANCE) {
      usize raw = u_distance(s2, len2, a.has_cutoff, a.cutoff_u, a.has_hint, a.hint_u, a);
      return {!a.has_cutoff || raw <= a.cutoff_u, raw};
    }
    usize raw = u_similarity(s2, len2, a.has_cutoff, a.cutoff_u, a.has_hint, a.hint_u, a);
    return {!a.has_cutoff || raw >= a.cutoff_u, raw};
  }
  // Float-valued results (normalized_* for every metric; distance/similarity for Jaro/JW; fuzz::ratio).
  template <class C2>
  OptF float_score(Kind k, const C2* s2, usize len2, const Args& a) const {
    double raw;
    bool dist_like = (k == DISTANCE || k == NORM_DISTANCE);
    if (metric == RATIO) {
      // fuzz.rs:127-149.  Documented semantics (== fuzz::ratio, fuzz.rs:60-85): Indel normalized
      // similarity.  With reference_quirks the literal code path is taken: the *inner*
      // lcs_seq::BatchComparator's _normalized_similarity (maximum = max(len1,len2)) -- quirk Q1.
      if (a.reference_quirks) {
        Batch<C1> inner(LCS_SEQ, s1.data(), s1.size());
        raw = inner.u_norm_similarity(s2, len2, a.has_cutoff, a.cutoff_f, a.has_hint, a.hint_f, a);
      } else {
        raw = u_norm_similarity(s2, len2, a.has_cutoff, a.cutoff_f, a.has_hint, a.hint_f, a);
      }
      return {!a.has_cutoff || raw >= a.cutoff_f, raw};
    }
    if (is_float_metric()) {
      switch (k) {
        case DISTANCE: raw = f_distance(s2, len2, a.has_cutoff, a.cutoff_f, a); break;
        case SIMILARITY: raw = f_similarity(s2, len2, a.has_cutoff, a.cutoff_f, a); break;
        case NORM_DISTANCE: raw = f_norm_distance(s2, len2, a.has_cutoff, a.cutoff_f, a); break;
        default: raw = f_norm_similarity(s2, len2, a.has_cutoff, a.cutoff_f, a); break;
      }
    } else {
      if (k == NORM_DISTANCE) raw = u_norm_distance(s2, len2, a.has_cutoff, a.cutoff_f, a.has_hint, a.hint_f, a);
      else raw = u_norm_similarity(s2, len2, a.has_cutoff, a.cutoff_f, a.has_hint, a.hint_f, a);
    }
    bool ok = !a.has_cutoff || (dist_like ? raw <= a.cutoff_f : raw >= a.cutoff_f);
    return {ok, raw};
  }
};

}  // namespace rfo
