// rf_core.cuh -- per-candidate arithmetic of the scoring kernels, written __host__ __device__ so that the
// very same code is unit-tested on the CPU (tests/test_core_host.py compiles it with g++) and runs inside
// the sm_100a kernels (rf_kernels.cu).
//
// What is computed follows rapidfuzz-rs 0.5.0 (paths relative to the reference's src/); how it is computed
// is GPU-first:
//  * Levenshtein / OSA (levenshtein.rs:435-507, osa.rs:84-135): Hyyro/Myers bit-vectors, but the pattern is
//    TOP-aligned in the machine word (bit BITS-len1 .. BITS-1) so no per-column score tracking is needed:
//    D[m][n] = n + popc(VP) - popc(VN) from the final vertical delta vectors.  The unused low bits stay at
//    VP=VN=0 and feed the +1 horizontal carry into the pattern's first row by themselves.
//  * LCSseq / Indel / ratio (lcs_seq.rs:199-261): S = (S + (S&M)) | (S & ~(S&M)), lcs = popc(~S).
//  * Jaro / Jaro-Winkler (jaro.rs:147-190, :339-368, :516-598; jaro_winkler.rs:103-141): flag pass +
//    transposition pass on 64-bit flags, f64 epilogue in the reference's operation order.
//  * score algebra (details/distance.rs:154-385, common.rs:43-45, :83-85) in finish_int / finish_float.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define RF_HD __host__ __device__ __forceinline__
#else
#define RF_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define RF_UNROLL _Pragma("unroll")
#else
#define RF_UNROLL  // host pass / plain C++ (the CPU tests of this header): gcc does not know the pragma
#endif

namespace rfk {

enum Metric : int { M_LEVENSHTEIN = 0, M_INDEL = 1, M_LCS_SEQ = 2, M_OSA = 3, M_JARO = 4, M_JARO_WINKLER = 5, M_RATIO = 6,
                    M_HAMMING = 7, M_PREFIX = 8, M_POSTFIX = 9, M_DAMERAU_LEVENSHTEIN = 10 };
enum Kind : int { K_DISTANCE = 0, K_SIMILARITY = 1, K_NORM_DISTANCE = 2, K_NORM_SIMILARITY = 3 };
// which bit-parallel recurrence a metric needs
enum Family : int { F_LEV = 0, F_LCS = 1, F_OSA = 2, F_JARO = 3, F_SIMPLE = 4, F_WF = 5, F_DL = 6 };
// F_SIMPLE: hamming / prefix / postfix; F_WF: generic Levenshtein weights; F_DL: Damerau-Levenshtein
// Levenshtein weight classes (levenshtein.rs:1301-1330)
enum WeightClass : int { WC_UNIFORM = 0, WC_INDEL = 1, WC_ZERO = 2, WC_GENERIC = 3 };

constexpr uint32_t NONE_U32 = 0xFFFFFFFFu;

RF_HD int popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
RF_HD int popc(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}
RF_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}
RF_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {  // (hi:lo) >> sh, sh in {0,8,16,24}
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
#endif
}

// Sequential reader over candidate bytes that start at an arbitrary byte of a 4-byte-aligned buffer.
// Reads whole aligned words (one word of over-read slack is required behind the data).
struct ByteReader {
  const uint32_t* p;
  uint32_t sh;
  uint32_t cur;
  RF_HD ByteReader(const uint8_t* aligned_base, uint32_t start) {
    p = reinterpret_cast<const uint32_t*>(aligned_base + (start & ~3u));
    sh = (start & 3u) * 8u;
    cur = *p++;
  }
  RF_HD uint32_t next4() {
    uint32_t nxt = *p++;
    uint32_t r = funnel_r(cur, nxt, sh);
    cur = nxt;
    return r;
  }
};

// Same for 16 bytes per call from a 16-byte-aligned buffer (one 128-bit load per 16 characters; a thread
// that walks its own candidate issues 4x fewer memory requests than with 4-byte words).  Needs 31 bytes of
// over-read slack behind the data.
struct Bytes16 { uint32_t w[4]; };
struct ByteReader16 {
  const uint32_t* p;  // next aligned 16-byte line, as words
  uint32_t sh, wsel;
  uint32_t cur[4];
  RF_HD ByteReader16(const uint8_t* aligned16_base, uint32_t start) {
    p = reinterpret_cast<const uint32_t*>(aligned16_base + (start & ~15u));
    sh = (start & 3u) * 8u;
    wsel = (start >> 2) & 3u;
    load(cur);
  }
  RF_HD void load(uint32_t* d) {
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
#else
    d[0] = p[0]; d[1] = p[1]; d[2] = p[2]; d[3] = p[3];
#endif
    p += 4;
  }
  RF_HD Bytes16 next16() {
    uint32_t nxt[4];
    load(nxt);
    const uint32_t W[8] = {cur[0], cur[1], cur[2], cur[3], nxt[0], nxt[1], nxt[2], nxt[3]};
    uint32_t A[6], B[5];
    const bool s2 = (wsel & 2u) != 0, s1 = (wsel & 1u) != 0;
RF_UNROLL
    for (int i = 0; i < 6; ++i) A[i] = s2 ? W[i + 2] : W[i];
RF_UNROLL
    for (int i = 0; i < 5; ++i) B[i] = s1 ? A[i + 1] : A[i];
    Bytes16 r;
RF_UNROLL
    for (int i = 0; i < 4; ++i) r.w[i] = funnel_r(B[i], B[i + 1], sh);
RF_UNROLL
    for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
    return r;
  }
};

// ------------------------------------------------------------------------------------------------
// Levenshtein, one machine word (query length 1..BITS).  `tab(ch)` returns the TOP-aligned match mask
// PM[ch] << (BITS - len1).
template <class W, class Tab, class Rd>
RF_HD uint32_t lev_w1(const Tab& tab, Rd rd, uint32_t len2, uint32_t len1) {
  constexpr int BITS = (int)sizeof(W) * 8;
  W VP = (W)(~(W)0) << (BITS - (int)len1);
  W VN = 0;
#define RF_LEV_STEP(CH)                               \
  {                                                   \
    const W X = tab(CH);                              \
    const W D0 = ((((X & VP) + VP) ^ VP) | X) | VN;   \
    W HP = VN | ~(D0 | VP);                           \
    W HN = D0 & VP;                                   \
    HP = (HP << 1) | (W)1;                            \
    HN = HN << 1;                                     \
    VP = HN | ~(D0 | HP);                             \
    VN = HP & D0;                                     \
  }
  const uint32_t nfull = len2 >> 2;
  for (uint32_t i = 0; i < nfull; ++i) {
    const uint32_t w = rd.next4();
    RF_LEV_STEP(w & 0xffu)
    RF_LEV_STEP((w >> 8) & 0xffu)
    RF_LEV_STEP((w >> 16) & 0xffu)
    RF_LEV_STEP(w >> 24)
  }
  const uint32_t rem = len2 & 3u;
  if (rem) {
    const uint32_t w = rd.next4();
    RF_LEV_STEP(w & 0xffu)
    if (rem > 1) RF_LEV_STEP((w >> 8) & 0xffu)
    if (rem > 2) RF_LEV_STEP((w >> 16) & 0xffu)
  }
#undef RF_LEV_STEP
  return len2 + (uint32_t)popc(VP) - (uint32_t)popc(VN);
}

// OSA, one machine word, TOP-aligned like lev_w1 (osa.rs:84-135: D0 |= TR with
// TR = (((~D0_prev) & PM_j) << 1) & PM_{j-1}).
template <class W, class Tab, class Rd>
RF_HD uint32_t osa_w1(const Tab& tab, Rd rd, uint32_t len2, uint32_t len1) {
  constexpr int BITS = (int)sizeof(W) * 8;
  W VP = (W)(~(W)0) << (BITS - (int)len1);
  W VN = 0, D0 = 0, PMold = 0;
#define RF_OSA_STEP(CH)                               \
  {                                                   \
    const W X = tab(CH);                              \
    const W TR = (((~D0) & X) << 1) & PMold;          \
    D0 = (((((X & VP) + VP) ^ VP) | X) | VN) | TR;    \
    W HP = VN | ~(D0 | VP);                           \
    W HN = D0 & VP;                                   \
    HP = (HP << 1) | (W)1;                            \
    HN = HN << 1;                                     \
    VP = HN | ~(D0 | HP);                             \
    VN = HP & D0;                                     \
    PMold = X;                                        \
  }
  const uint32_t nfull = len2 >> 2;
  for (uint32_t i = 0; i < nfull; ++i) {
    const uint32_t w = rd.next4();
    RF_OSA_STEP(w & 0xffu)
    RF_OSA_STEP((w >> 8) & 0xffu)
    RF_OSA_STEP((w >> 16) & 0xffu)
    RF_OSA_STEP(w >> 24)
  }
  const uint32_t rem = len2 & 3u;
  if (rem) {
    const uint32_t w = rd.next4();
    RF_OSA_STEP(w & 0xffu)
    if (rem > 1) RF_OSA_STEP((w >> 8) & 0xffu)
    if (rem > 2) RF_OSA_STEP((w >> 16) & 0xffu)
  }
#undef RF_OSA_STEP
  return len2 + (uint32_t)popc(VP) - (uint32_t)popc(VN);
}

// LCS length, one machine word; `tab(ch)` is the plain (bottom-aligned) PM[ch] (lcs_seq.rs:222-257).
template <class W, class Tab, class Rd>
RF_HD uint32_t lcs_w1(const Tab& tab, Rd rd, uint32_t len2) {
  W S = ~(W)0;
#define RF_LCS_STEP(CH)              \
  {                                  \
    const W U = S & tab(CH);         \
    S = (S + U) | (S & ~U);          \
  }
  const uint32_t nfull = len2 >> 2;
  for (uint32_t i = 0; i < nfull; ++i) {
    const uint32_t w = rd.next4();
    RF_LCS_STEP(w & 0xffu)
    RF_LCS_STEP((w >> 8) & 0xffu)
    RF_LCS_STEP((w >> 16) & 0xffu)
    RF_LCS_STEP(w >> 24)
  }
  const uint32_t rem = len2 & 3u;
  if (rem) {
    const uint32_t w = rd.next4();
    RF_LCS_STEP(w & 0xffu)
    if (rem > 1) RF_LCS_STEP((w >> 8) & 0xffu)
    if (rem > 2) RF_LCS_STEP((w >> 16) & 0xffu)
  }
#undef RF_LCS_STEP
  return (uint32_t)popc((W)~S);
}

// ------------------------------------------------------------------------------------------------
// Banded Levenshtein for a distance cutoff k <= 63 and ANY query length (the job of the reference's
// hyrroe2003_small_band, levenshtein.rs:509-617, and of the Ukkonen band in hyrroe2003_block, :897-985).
// GPU-first formulation: a cell on diagonal h = j - i can lie on a path of cost <= k only if
// |h| + |d - h| <= k (d = len2 - len1), i.e. on one of at most k + 1 <= 64 diagonals, so ONE 64-bit window that
// slides down one pattern row per text column holds the whole band whatever the query length -- one thread per
// candidate, no cross-lane carries.  Window bit b at (1-based) column j is pattern row j - h_hi + b.
//   * rows above the matrix (<= 0) are a flat region (vertical delta 0, no match), which yields D[0][j] = j;
//   * the row entering at the bottom takes the diagonal-move bound (VP = HN | ~HP of its upper neighbour);
//   * the row leaving at the top becomes the boundary with horizontal delta +1.
//   Out-of-band inputs are therefore never under-estimated, in-band cells of any path of cost <= k are exact,
//   and values along a diagonal never decrease: once the cell on the end diagonal d exceeds k the result is
//   None.  `bn` is the value of the window's first row, the score of the end diagonal follows by popcounts.
struct LevBand64 {
  uint64_t VP, VN, mask;
  int32_t bn;  // D'[first window row][j] after column j
  int32_t s;   // 0-based pattern index of window bit 0 at the next column (may be negative)
  // requires |len2 - len1| <= k <= 63
  RF_HD void init(uint32_t len1, uint32_t len2, uint32_t k) {
    const int32_t d = (int32_t)len2 - (int32_t)len1;
    const int32_t ad = d < 0 ? -d : d;
    const int32_t e = ((int32_t)k - ad) / 2;
    const int32_t h_hi = (d > 0 ? d : 0) + e;  // highest diagonal of the band, <= k
    VP = ~0ull << h_hi;                        // rows >= 1 of column 0: vertical delta +1
    VN = 0;
    mask = (1ull << (h_hi - d)) - 1ull;        // window bits above the end diagonal's row (h_hi - d <= k)
    bn = 0;
    s = -h_hi;
  }
  // X = window of the match vector of text char j: bit b set iff query[s + b] == char (0 outside 0..len1-1)
  RF_HD void step(uint64_t X) {
    const uint64_t D0 = (((X & VP) + VP) ^ VP) | X | VN;
    const uint64_t HP = VN | ~(D0 | VP);
    const uint64_t HN = D0 & VP;
    bn += 1 - (int32_t)((uint32_t)D0 & 1u);
    const uint64_t D0s = D0 >> 1;  // the window moves down one row: shift D0 instead of HP/HN
    VP = HN | ~(D0s | HP);
    VN = D0s & HP;
    ++s;
  }
  // value of the cell on the end diagonal in the column processed last (== the distance after column len2)
  RF_HD int32_t score() const { return bn + popc(VP & mask) - popc(VN & mask); }
};

// 64-bit window starting at (possibly negative) bit position s of a match-vector row.  The row is stored as
// 32-bit words with 64 zero bits in front and behind (row32[2 + w] = bits 32w.. of the match vector), so
// position s + 64 >= 1 selects three consecutive words a,b,c at index (s+64)>>5 and two funnel shifts.
RF_HD uint64_t band_window32(uint32_t a, uint32_t b, uint32_t c, uint32_t o) {
  return (uint64_t)funnel_r(a, b, o) | ((uint64_t)funnel_r(b, c, o) << 32);
}
// k <= 32 only.  Only the band's own bits (at most k+1 <= 33 from bit 0) have to be right: rows above them in the window lie
// outside the band, hold over-estimates anyway, and never influence lower bits (carries only travel up).  Two
// words give window bits 0..(63 - o) >= 32, the rest reads as no-match.
RF_HD uint64_t band_window32_low33(uint32_t a, uint32_t b, uint32_t o) {
  return (uint64_t)funnel_r(a, b, o) | ((uint64_t)(b >> o) << 32);
}

// ------------------------------------------------------------------------------------------------
// Generic weighted Levenshtein (levenshtein.rs:212-259, the Wagner-Fischer route of :1330): one cost row of
// len1+1 entries, `cache(i)` is a reference to entry i (strided global scratch on the device).  A match takes the
// diagonal without looking at the other two moves, exactly as the reference does.
template <class QB, class TB, class Cache>
RF_HD uint64_t weighted_wagner_fischer(const QB& qb, const TB& tb, uint32_t len1, uint32_t len2, uint64_t w_ins, uint64_t w_del,
                                       uint64_t w_sub, const Cache& cache) {
  for (uint32_t i = 0; i <= len1; ++i) cache(i) = (uint64_t)i * w_del;
  for (uint32_t j = 0; j < len2; ++j) {
    const uint32_t ch2 = tb(j);
    uint64_t temp = cache(0);
    cache(0) += w_ins;
    for (uint32_t i = 0; i < len1; ++i) {
      uint64_t x = temp;
      const uint64_t old_next = cache(i + 1);
      if (qb(i) != ch2) {
        const uint64_t a = cache(i) + w_del, b = temp + w_sub;
        x = a < b ? a : b;
        const uint64_t c = old_next + w_ins;
        x = x < c ? x : c;
      }
      cache(i + 1) = x;
      temp = old_next;
    }
  }
  return cache(len1);
}

// ------------------------------------------------------------------------------------------------
// Damerau-Levenshtein (unrestricted), Zhao & Sahni's linear-space algorithm as in damerau_levenshtein.rs:111-168.
// The distance is symmetric, so the device runs it with the CANDIDATE as the outer sequence and the QUERY as the
// inner one: the three rows then have len_inner+2 = len1+2 entries whatever the candidate's length.
//   row(k, j)   reference to entry j of row k in {0,1,2}
//   last_row(c) reference to "last outer row in which symbol c was seen", all -1 on entry and restored on exit
template <class Outer, class Inner, class Row, class LastRow>
RF_HD uint32_t damerau_zhao(const Outer& outer, uint32_t len_o, const Inner& inner, uint32_t len_i, const Row& row,
                            const LastRow& last_row) {
  const int32_t max_val = (int32_t)(len_o > len_i ? len_o : len_i) + 1;
  const uint32_t size = len_i + 2;
  uint32_t FR = 0, R1 = 1, R = 2;
  for (uint32_t j = 0; j < size; ++j) {
    row(FR, j) = max_val;
    row(R1, j) = max_val;
    row(R, j) = j == 0 ? max_val : (int32_t)j - 1;
  }
  for (uint32_t i = 1; i <= len_o; ++i) {
    const uint32_t ch1 = outer(i - 1);
    const uint32_t sw = R; R = R1; R1 = sw;
    int32_t last_col_id = -1;
    int32_t last_i2l1 = row(R, 1);
    row(R, 1) = (int32_t)i;
    int32_t t = max_val;
    for (uint32_t j = 1; j <= len_i; ++j) {
      const uint32_t ch2 = inner(j - 1);
      const int32_t diag = row(R1, j) + (ch1 != ch2 ? 1 : 0);
      const int32_t left = row(R, j) + 1;
      const int32_t up = row(R1, j + 1) + 1;
      int32_t temp = diag < left ? diag : left;
      temp = temp < up ? temp : up;
      if (ch1 == ch2) {
        last_col_id = (int32_t)j;        // last occurrence of the outer symbol in the inner sequence
        row(FR, j + 1) = row(R1, j - 1);  // H[i-2][j-2]
        t = last_i2l1;                    // H[i-2][l-1]
      } else {
        const int32_t k = last_row(ch2);
        const int32_t l = last_col_id;
        if ((int32_t)j - l == 1) {
          const int32_t tr = row(FR, j + 1) + ((int32_t)i - k);
          temp = temp < tr ? temp : tr;
        } else if ((int32_t)i - k == 1) {
          const int32_t tr = t + ((int32_t)j - l);
          temp = temp < tr ? temp : tr;
        }
      }
      last_i2l1 = row(R, j + 1);
      row(R, j + 1) = temp;
    }
    last_row(ch1) = (int32_t)i;
  }
  const uint32_t result = (uint32_t)row(R, len_i + 1);
  for (uint32_t i = 0; i < len_o; ++i) last_row(outer(i)) = -1;
  return result;
}

// ------------------------------------------------------------------------------------------------
// Hamming / Prefix / Postfix (hamming.rs:136-161, details/common.rs:39-62): no bit-parallelism needed, the
// candidate is compared with the query 4 bytes at a time.  q4(i) / t4(i) return bytes 4i..4i+3 of the query /
// candidate packed little-endian (garbage beyond the end is masked here); qb(j) / tb(j) single bytes.
RF_HD uint32_t differing_bytes(uint32_t a, uint32_t b) {  // number of byte lanes in which a and b differ
  const uint32_t x = a ^ b;
  // bit 7 of every byte = that byte of x is non-zero (low 7 bits carry into bit 7, or bit 7 itself is set)
  const uint32_t y = (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
  return (uint32_t)popc(y);
}
template <class Q4, class T4>
RF_HD uint32_t hamming_raw(const Q4& q4, const T4& t4, uint32_t len1, uint32_t len2) {
  const uint32_t mn = len1 < len2 ? len1 : len2, mx = len1 < len2 ? len2 : len1;
  uint32_t dist = mx - mn;  // with pad: the excess counts as mismatches (hamming.rs:156-158)
  uint32_t i = 0;
  for (; i + 4 <= mn; i += 4) dist += differing_bytes(q4(i >> 2), t4(i >> 2));
  if (i < mn) {
    const uint32_t mask = (1u << (8 * (mn - i))) - 1u;
    dist += differing_bytes(q4(i >> 2) & mask, t4(i >> 2) & mask);
  }
  return dist;
}
template <class Q4, class T4>
RF_HD uint32_t prefix_raw(const Q4& q4, const T4& t4, uint32_t len1, uint32_t len2) {
  const uint32_t mn = len1 < len2 ? len1 : len2;
  for (uint32_t i = 0; i < mn; i += 4) {
    const uint32_t x = q4(i >> 2) ^ t4(i >> 2);
    if (x) {
      uint32_t k = 0;
      while (!((x >> (8 * k)) & 0xffu)) ++k;  // first differing byte of the word
      const uint32_t n = i + k;
      return n < mn ? n : mn;
    }
  }
  return mn;
}
template <class QB, class TB>
RF_HD uint32_t postfix_raw(const QB& qb, const TB& tb, uint32_t len1, uint32_t len2) {
  const uint32_t mn = len1 < len2 ? len1 : len2;
  uint32_t n = 0;
  while (n < mn && qb(len1 - 1 - n) == tb(len2 - 1 - n)) ++n;
  return n;
}

// ------------------------------------------------------------------------------------------------
// Jaro building blocks (jaro.rs:106-145)
// x / 3.0, correctly rounded, without the general division routine (Markstein: y = RN(1/3), q = RN(x*y),
// r = x - 3q exactly (fma), result RN(q + r*y) is the correctly rounded quotient for normal operands).
RF_HD double div3_exact(double x) {
  const double y = 0.33333333333333331482961625624739;  // RN(1/3)
  const double q = x * y;
  const double r = fma(-3.0, q, x);
  return fma(r, y, q);
}
// Jaro formula (jaro.rs:106-119) with the three quotients of small integers read from a table of exactly
// rounded a/b (a, b <= 64) built on the host: bit-identical to the divisions, ~10x fewer instructions.
constexpr int kQuotDim = 65;
RF_HD double jaro_calculate_similarity_tab(const double* __restrict__ quot, uint32_t p_len, uint32_t t_len, uint32_t cc,
                                           uint32_t transpositions) {
  transpositions /= 2;
  double sim = 0.0;
  sim += quot[cc * kQuotDim + p_len];
  sim += quot[cc * kQuotDim + t_len];
  sim += quot[(cc - transpositions) * kQuotDim + cc];
  return div3_exact(sim);
}
RF_HD double jaro_calculate_similarity(uint32_t p_len, uint32_t t_len, uint32_t cc, uint32_t transpositions) {
  transpositions /= 2;
  double sim = 0.0;
  sim += (double)cc / (double)p_len;
  sim += (double)cc / (double)t_len;
  sim += ((double)cc - (double)transpositions) / (double)cc;
  return sim / 3.0;
}
RF_HD bool jaro_length_filter(uint32_t p_len, uint32_t t_len, double cutoff) {
  if (t_len == 0 || p_len == 0) return false;
  if (cutoff <= 0.0) return true;  // the bound below is positive: same answer without the three divisions
  const double min_len = (double)(p_len < t_len ? p_len : t_len);
  double sim = min_len / (double)p_len + min_len / (double)t_len + 1.0;
  sim /= 3.0;
  return sim >= cutoff;
}
RF_HD bool jaro_common_char_filter(uint32_t p_len, uint32_t t_len, uint32_t cc, double cutoff) {
  if (cc == 0) return false;
  if (cutoff <= 0.0) return true;  // as above
  double sim = 0.0;
  sim += (double)cc / (double)p_len;
  sim += (double)cc / (double)t_len;
  sim += 1.0;
  sim /= 3.0;
  return sim >= cutoff;
}

// Jaro similarity with a cached query (jaro.rs:516-598).
//   pmw(word, ch) -> bottom-aligned PM word; bytes(j) -> candidate byte j.
//   Query length 1..MAXQ (MAXQ multiple of 64); the candidate may have any length: instead of the
//   reference's t_flag bit-vector (which grows with the candidate) the matched text characters are kept
//   in order (at most len1 of them), which yields the identical transposition count.
template <int MAXQ, class MT = uint8_t, class PMW, class Bytes>
RF_HD double jaro_similarity_generic(const PMW& pmw, const Bytes& bytes, uint32_t len1, uint32_t len2, double cutoff) {
  const uint32_t len1_orig = len1, len2_orig = len2;
  if (cutoff > 1.0) return 0.0;
  if (len1_orig == 0 && len2_orig == 0) return 1.0;
  if (!jaro_length_filter(len1_orig, len2_orig, cutoff)) return 0.0;
  if (len1_orig == 1 && len2_orig == 1) return (pmw(0u, bytes(0u)) & 1u) ? 1.0 : 0.0;
  uint32_t bound;
  if (len2 > len1) {
    bound = len2 / 2 - 1;
    if (len2 > len1 + bound) len2 = len1 + bound;
  } else {
    bound = len1 / 2 - 1;
    if (len1 > len2 + bound) len1 = len2 + bound;
  }
  if (len1 == 0 || len2 == 0) return jaro_calculate_similarity(len1_orig, len2_orig, 0, 0);  // NaN like the reference (0/0); unreachable: both >= 1 here
  constexpr int PW = MAXQ / 64;
  uint64_t P[PW];
  MT matched[MAXQ];  // MT = uint16_t for candidates renamed to 16-bit codes
RF_UNROLL
  for (int i = 0; i < PW; ++i) P[i] = 0;
  const uint32_t words = (len1_orig + 63) / 64;
  uint32_t cc = 0;
  for (uint32_t j = 0; j < len2; ++j) {
    // window of pattern positions [lo, hi] (jaro.rs:168-187 / :306-334), clipped to the truncated len1
    const uint32_t lo = j > bound ? j - bound : 0;
    uint32_t hi = j + bound;
    if (hi >= len1) hi = len1 - 1;
    if (lo > hi) continue;
    const uint32_t ch = bytes(j);
    const uint32_t w0 = lo / 64, w1 = hi / 64;
    for (uint32_t w = w0; w <= w1 && w < words; ++w) {
      uint64_t m = pmw(w, ch) & ~P[w];
      if (w == w0) m &= ~0ULL << (lo % 64);
      if (w == w1) m &= ~0ULL >> (63 - (hi % 64));
      if (m) {
        P[w] |= m & (0 - m);
        matched[cc++] = (MT)ch;
        break;
      }
    }
  }
  if (!jaro_common_char_filter(len1_orig, len2_orig, cc, cutoff)) return 0.0;
  // transpositions: k-th flagged pattern position vs k-th matched text character (jaro.rs:339-420)
  uint32_t transpositions = 0, k = 0;
  for (uint32_t w = 0; w < words; ++w) {
    uint64_t p = P[w];
    while (p) {
      const uint64_t bit = p & (0 - p);
      transpositions += (pmw(w, (uint32_t)matched[k]) & bit) == 0;
      ++k;
      p ^= bit;
    }
  }
  return jaro_calculate_similarity(len1_orig, len2_orig, cc, transpositions);
}

// Fast path: query <= 64 and (truncated) candidate <= 64: flags in two registers (jaro.rs:147-190, :339-368).
// tab(ch) -> bottom-aligned 64-bit PM; bytes(j) -> candidate byte j.  Falls back to the generic routine
// when the truncated candidate is longer than 64 (the reference's block path, jaro.rs:584-595).
template <class Tab, class Bytes>
RF_HD double jaro_similarity_w1(const Tab& tab, const Bytes& bytes, uint32_t len1, uint32_t len2, double cutoff) {
  const uint32_t len1_orig = len1, len2_orig = len2;
  if (cutoff > 1.0) return 0.0;
  if (len1_orig == 0 && len2_orig == 0) return 1.0;
  if (!jaro_length_filter(len1_orig, len2_orig, cutoff)) return 0.0;
  if (len1_orig == 1 && len2_orig == 1) return (tab(bytes(0u)) & 1u) ? 1.0 : 0.0;
  uint32_t bound;
  if (len2 > len1) {
    bound = len2 / 2 - 1;
    if (len2 > len1 + bound) len2 = len1 + bound;
  } else {
    bound = len1 / 2 - 1;
    if (len1 > len2 + bound) len1 = len2 + bound;
  }
  if (len2 > 64) {
    auto pmw = [&](uint32_t, uint32_t ch) -> uint64_t { return tab(ch); };
    return jaro_similarity_generic<64>(pmw, bytes, len1_orig, len2_orig, cutoff);
  }
  uint64_t P = 0, T = 0;
  uint64_t bound_mask = (bound + 1 < 64) ? ((1ULL << (bound + 1)) - 1) : ~0ULL;  // bit_mask_lsb_u64(bound+1)
  uint32_t j = 0;
  const uint32_t n0 = bound < len2 ? bound : len2;
  for (; j < n0; ++j) {
    const uint64_t m = tab(bytes(j)) & bound_mask & ~P;
    P |= m & (0 - m);
    T |= (uint64_t)(m != 0) << j;
    bound_mask = (bound_mask << 1) | 1;
  }
  for (; j < len2; ++j) {
    const uint64_t m = tab(bytes(j)) & bound_mask & ~P;
    P |= m & (0 - m);
    T |= (uint64_t)(m != 0) << j;
    bound_mask <<= 1;
  }
  const uint32_t cc = (uint32_t)popc(P);
  if (!jaro_common_char_filter(len1_orig, len2_orig, cc, cutoff)) return 0.0;
  uint32_t transpositions = 0;
  while (T) {
    const uint64_t pbit = P & (0 - P);
    const int idx = ctz64(T);
    transpositions += (tab(bytes((uint32_t)idx)) & pbit) == 0;
    T &= T - 1;
    P ^= pbit;
  }
  return jaro_calculate_similarity(len1_orig, len2_orig, cc, transpositions);
}

// Jaro flag + transposition passes for query <= 32 and (truncated) candidate <= 64, written for a warp whose
// lanes walk 8-byte rows in lock step (jaro.rs:147-190, :339-368): P (pattern flags) and the search window are
// 32-bit, the text flags are collected 8 bits per row.  `tab(ch)` is the bottom-aligned 32-bit PM, `row(r)`
// returns text bytes 8r..8r+7 (anything beyond len2), `nrows` >= ceil(len2/8) may be larger than this lane needs
// (warp-uniform loop bound): characters at j >= len2 are masked out.  No early exits, no data-dependent loops:
// per character ~8 instructions in pass 1 and ~9 in pass 2, identical control flow in all 32 lanes.
//   len1, len2 are the TRUNCATED lengths and bound the window radius (jaro.rs:553-565); requires len2 <= 64.
struct Jaro32Result { uint32_t cc, transpositions; };
template <class Tab, class Row>
RF_HD Jaro32Result jaro32_rows(const Tab& tab, const Row& row, uint32_t len2, uint32_t bound, uint32_t nrows) {
  if (nrows > 8) nrows = 8;  // len2 <= 64
  uint32_t P = 0;
  uint64_t T = 0;
  // window for text position j: pattern bits [j - bound, j + bound]; hi grows by one bit per character,
  // lo drops one bit per character once j > bound
  uint32_t hi = (bound + 1 < 32) ? ((1u << (bound + 1)) - 1u) : 0xFFFFFFFFu;
  uint32_t lo = 0xFFFFFFFFu;
  for (uint32_t r = 0; r < nrows; ++r) {
    const uint2 v = row(r);
    uint32_t t8 = 0;
RF_UNROLL
    for (int t = 0; t < 8; ++t) {
      const uint32_t j = r * 8u + (uint32_t)t;
      const uint32_t ch = ((t < 4 ? v.x : v.y) >> (8 * (t & 3))) & 0xffu;
      uint32_t m = tab(ch) & hi & lo & ~P;
      if (j >= len2) m = 0;
      P |= m & (0u - m);
      t8 |= (uint32_t)(m != 0) << t;
      hi = (hi << 1) | 1u;
      if (j >= bound) lo <<= 1;
    }
    T |= (uint64_t)t8 << (8 * r);
  }
  Jaro32Result res;
  res.cc = (uint32_t)popc(P);
  uint32_t tr = 0;
  for (uint32_t r = 0; r < nrows; ++r) {
    const uint2 v = row(r);
    const uint32_t t8 = (uint32_t)(T >> (8 * r)) & 0xffu;
RF_UNROLL
    for (int t = 0; t < 8; ++t) {
      const uint32_t ch = ((t < 4 ? v.x : v.y) >> (8 * (t & 3))) & 0xffu;
      const uint32_t pbit = P & (0u - P);
      const bool hit = (t8 >> t) & 1u;
      if (hit) {
        tr += (tab(ch) & pbit) == 0;
        P ^= pbit;
      }
    }
  }
  res.transpositions = tr;
  return res;
}

// Jaro similarity from the pass results, with the reference's filters in the reference's order
// (jaro.rs:516-598); len1/len2 are the ORIGINAL lengths, first_match = query[0] == text[0] (1 x 1 case).
//   quot: optional table of exact quotients (see jaro_calculate_similarity_tab), used when both lengths fit.
RF_HD double jaro32_finish(uint32_t len1, uint32_t len2, const Jaro32Result& r, bool first_match, double cutoff,
                           const double* __restrict__ quot = nullptr) {
  if (cutoff > 1.0) return 0.0;
  if (len1 == 0 && len2 == 0) return 1.0;
  if (!jaro_length_filter(len1, len2, cutoff)) return 0.0;
  if (len1 == 1 && len2 == 1) return first_match ? 1.0 : 0.0;
  if (!jaro_common_char_filter(len1, len2, r.cc, cutoff)) return 0.0;
  if (quot && len1 < (uint32_t)kQuotDim && len2 < (uint32_t)kQuotDim)
    return jaro_calculate_similarity_tab(quot, len1, len2, r.cc, r.transpositions);
  return jaro_calculate_similarity(len1, len2, r.cc, r.transpositions);
}
// truncated lengths and window radius (jaro.rs:553-565)
RF_HD void jaro_bounds(uint32_t& len1, uint32_t& len2, uint32_t& bound) {
  if (len2 > len1) {
    bound = len2 / 2 - 1;
    if (len2 > len1 + bound) len2 = len1 + bound;
  } else {
    bound = len1 / 2 - 1;  // wraps for len1 < 2: only reached for 1 x 1 / empty inputs, whose result ignores the passes
    if (len1 > len2 + bound) len1 = len2 + bound;
  }
}

// ------------------------------------------------------------------------------------------------
// Score algebra.  `Epi` is the per-launch image of rf_args + metric + kind.
struct Epi {
  int metric;
  int kind;
  int has_cutoff;
  uint64_t cutoff_u;
  double cutoff_f;
  int wclass;          // Levenshtein only
  uint64_t w_ins, w_del, w_sub;
  double prefix_weight;
  int quirks;
  int unit32;  // set by the launcher: integer metric with unit weights -> 32-bit epilogue
  int pad;     // hamming::Args::pad
};

RF_HD Family family_of(int metric, int wclass) {
  switch (metric) {
    case M_LEVENSHTEIN: return wclass == WC_INDEL ? F_LCS : (wclass == WC_GENERIC ? F_WF : F_LEV);
    case M_INDEL: case M_LCS_SEQ: case M_RATIO: return F_LCS;
    case M_OSA: return F_OSA;
    case M_HAMMING: case M_PREFIX: case M_POSTFIX: return F_SIMPLE;
    case M_DAMERAU_LEVENSHTEIN: return F_DL;
    default: return F_JARO;
  }
}
RF_HD bool result_is_float(int metric, int kind) {
  if (metric == M_JARO || metric == M_JARO_WINKLER || metric == M_RATIO) return true;
  return kind == K_NORM_DISTANCE || kind == K_NORM_SIMILARITY;
}

RF_HD uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }
RF_HD uint64_t umax64(uint64_t a, uint64_t b) { return a > b ? a : b; }

// maximum(len1,len2) per metric (levenshtein.rs:263-277, indel.rs:331, lcs_seq.rs:773, osa.rs:432)
RF_HD uint64_t int_maximum(const Epi& e, uint64_t len1, uint64_t len2) {
  switch (e.metric) {
    case M_LEVENSHTEIN: {
      const uint64_t max_dist = len1 * e.w_del + len2 * e.w_ins;
      if (len1 >= len2) return umin64(max_dist, len2 * e.w_sub + (len1 - len2) * e.w_del);
      return umin64(max_dist, len1 * e.w_sub + (len2 - len1) * e.w_ins);
    }
    case M_INDEL: return len1 + len2;
    case M_RATIO: return e.quirks ? umax64(len1, len2) : len1 + len2;
    default: return umax64(len1, len2);
  }
}

// exact distance of the metric from the raw kernel result
//   F_LEV/F_OSA raw = unit-cost distance, F_LCS raw = LCS length
RF_HD uint64_t int_distance(const Epi& e, uint64_t raw, uint64_t len1, uint64_t len2) {
  switch (e.metric) {
    case M_LEVENSHTEIN:
      if (e.wclass == WC_ZERO) return 0;                                     // levenshtein.rs:1303-1305
      if (e.wclass == WC_GENERIC) return raw;                                // :1330 (raw = weighted Wagner-Fischer result)
      if (e.wclass == WC_INDEL) return (len1 + len2 - 2 * raw) * e.w_ins;    // :1321-1327
      return raw * e.w_ins;                                                  // :1308-1316
    case M_INDEL: return len1 + len2 - 2 * raw;                              // indel.rs:367
    case M_RATIO: return e.quirks ? umax64(len1, len2) - raw : len1 + len2 - 2 * raw;
    case M_LCS_SEQ: case M_PREFIX: case M_POSTFIX: return umax64(len1, len2) - raw;  // details/distance.rs:178 (raw = similarity)
    default: return raw;                                                     // OSA, Hamming
  }
}

// Integer-valued kinds.  Returns NONE_U32 for `None`.
RF_HD uint32_t finish_int(const Epi& e, uint64_t raw, uint64_t len1, uint64_t len2) {
  if (e.unit32) {  // unit weights and 32-bit lengths (the common case): same algebra in 32-bit arithmetic
    const uint32_t r = (uint32_t)raw, l1 = (uint32_t)len1, l2 = (uint32_t)len2;
    const uint32_t mx = l1 > l2 ? l1 : l2;
    uint32_t d, M;
    switch (e.metric) {
      case M_INDEL: d = l1 + l2 - 2u * r; M = l1 + l2; break;
      case M_LCS_SEQ: case M_PREFIX: case M_POSTFIX: d = mx - r; M = mx; break;
      default: d = r; M = mx; break;  // Levenshtein (1,1,1), OSA, Hamming
    }
    const uint32_t v = (e.kind == K_DISTANCE) ? d : M - d;
    if (e.has_cutoff) {
      const bool ok = (e.kind == K_DISTANCE) ? ((uint64_t)v <= e.cutoff_u) : ((uint64_t)v >= e.cutoff_u);
      if (!ok) return NONE_U32;
    }
    return v;
  }
  const uint64_t d = int_distance(e, raw, len1, len2);
  if (e.kind == K_DISTANCE) {
    if (e.has_cutoff && d > e.cutoff_u) return NONE_U32;                     // common.rs:43-45
    return (uint32_t)d;
  }
  const uint64_t M = int_maximum(e, len1, len2);
  const uint64_t s = M - d;
  if (e.has_cutoff && s < e.cutoff_u) return NONE_U32;                       // common.rs:83-85
  return (uint32_t)s;
}

RF_HD double clamp01(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }
RF_HD double qnan() {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(0x7ff8000000000000LL);
#else
  return NAN;
#endif
}

// normalized_distance / normalized_similarity of the integer metrics (+ fuzz::ratio).  NaN for `None`.
// Follows details/distance.rs:213-274 literally: the float cutoff is first turned into an integer
// distance cutoff (internal early-out), the final `score()` filter is applied on the float result.
RF_HD double finish_norm(const Epi& e, uint64_t raw, uint64_t len1, uint64_t len2) {
  const uint64_t d = int_distance(e, raw, len1, len2);
  const uint64_t M = int_maximum(e, len1, len2);
  const bool sim_kind = (e.kind == K_NORM_SIMILARITY) || (e.metric == M_RATIO);
  if (e.has_cutoff) {
    const double c = sim_kind ? fmin(1.0 - e.cutoff_f + 0.00001, 1.0) : e.cutoff_f;   // details/common.rs:4-7
    const uint64_t cd = (uint64_t)ceil((double)M * clamp01(c));                         // :231-236
    if (d > cd) return qnan();
  }
  const double nd = (M == 0) ? 0.0 : (double)d / (double)M;                            // :247-251
  if (!sim_kind) {
    if (e.has_cutoff && !(nd <= e.cutoff_f)) return qnan();
    return nd;
  }
  const double ns = 1.0 - nd;                                                           // :273
  if (e.has_cutoff && !(ns >= e.cutoff_f)) return qnan();
  return ns;
}

// Jaro / Jaro-Winkler: the four kinds through Metricf64 (details/distance.rs:277-385), maximum = 1.0.
//   sim_fn(cutoff) computes the metric's `_similarity` for the given internal cutoff.
template <class SimFn>
RF_HD double finish_float(const Epi& e, const SimFn& sim_fn) {
  const bool has = e.has_cutoff != 0;
  const double c = e.cutoff_f;
  switch (e.kind) {
    case K_SIMILARITY: {
      const double sim = sim_fn(has ? c : 0.0);
      return (has && !(sim >= c)) ? qnan() : sim;
    }
    case K_DISTANCE: {
      const double cs = has ? (1.0 >= c ? 1.0 - c : 0.0) : 0.0;                // :297
      const double dist = 1.0 - sim_fn(cs);                                    // :300-301
      return (has && !(dist <= c)) ? qnan() : dist;
    }
    case K_NORM_DISTANCE: {
      const double cd = has ? 1.0 * c : 0.0;                                   // :353
      const double cs = has ? (1.0 >= cd ? 1.0 - cd : 0.0) : 0.0;
      const double dist = 1.0 - sim_fn(cs);
      const double nd = dist / 1.0;                                            // :357-358
      return (has && !(nd <= c)) ? qnan() : nd;
    }
    default: {
      const double cn = has ? fmin(1.0 - c + 0.00001, 1.0) : 0.0;              // :379
      const double cd = has ? 1.0 * cn : 0.0;
      const double cs = has ? (1.0 >= cd ? 1.0 - cd : 0.0) : 0.0;
      const double dist = 1.0 - sim_fn(cs);
      const double nd = dist / 1.0;
      const double ns = 1.0 - nd;                                              // :382-383
      return (has && !(ns >= c)) ? qnan() : ns;
    }
  }
}

// Jaro-Winkler on top of a Jaro routine (jaro_winkler.rs:103-141). prefix = common prefix length (<= 4).
template <class JaroFn>
RF_HD double jaro_winkler_from(const JaroFn& jaro_fn, uint32_t prefix, double prefix_weight, double cutoff) {
  double jaro_cutoff = cutoff;
  if (jaro_cutoff > 0.7) {
    const double prefix_sim = (double)prefix * prefix_weight;
    if (prefix_sim >= 1.0) jaro_cutoff = 0.7;
    else jaro_cutoff = fmax(0.7, (prefix_sim - jaro_cutoff) / (prefix_sim - 1.0));
  }
  double sim = jaro_fn(jaro_cutoff);
  if (sim > 0.7) sim += (double)prefix * prefix_weight * (1.0 - sim);
  return sim;
}

}  // namespace rfk
