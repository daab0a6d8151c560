// rf_textbook.hpp -- TEST INFRASTRUCTURE ONLY.
// Independent "second opinion" for the oracle: plain O(N*M) dynamic programs straight from the textbook
// definitions (no bit-parallelism, no code shared with rf_oracle.hpp).  Used to cross-check the oracle on
// random inputs, as SURVEY.md section 8(c) asks.
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <vector>

namespace rftb {

template <class C1, class C2>
uint64_t levenshtein(const C1* a, uint64_t n, const C2* b, uint64_t m, uint64_t ins = 1, uint64_t del = 1, uint64_t sub = 1) {
  std::vector<uint64_t> prev(m + 1), cur(m + 1);
  for (uint64_t j = 0; j <= m; ++j) prev[j] = j * ins;
  for (uint64_t i = 1; i <= n; ++i) {
    cur[0] = i * del;
    for (uint64_t j = 1; j <= m; ++j) {
      uint64_t c = prev[j - 1] + (((uint64_t)a[i - 1] == (uint64_t)b[j - 1]) ? 0 : sub);
      c = std::min(c, prev[j] + del);
      c = std::min(c, cur[j - 1] + ins);
      cur[j] = c;
    }
    std::swap(prev, cur);
  }
  return prev[m];
}

template <class C1, class C2>
uint64_t lcs(const C1* a, uint64_t n, const C2* b, uint64_t m) {
  std::vector<uint64_t> prev(m + 1, 0), cur(m + 1, 0);
  for (uint64_t i = 1; i <= n; ++i) {
    cur[0] = 0;
    for (uint64_t j = 1; j <= m; ++j)
      cur[j] = ((uint64_t)a[i - 1] == (uint64_t)b[j - 1]) ? prev[j - 1] + 1 : std::max(prev[j], cur[j - 1]);
    std::swap(prev, cur);
  }
  return prev[m];
}

// optimal string alignment (restricted Damerau-Levenshtein)
template <class C1, class C2>
uint64_t osa(const C1* a, uint64_t n, const C2* b, uint64_t m) {
  std::vector<std::vector<uint64_t>> d(n + 1, std::vector<uint64_t>(m + 1));
  for (uint64_t i = 0; i <= n; ++i) d[i][0] = i;
  for (uint64_t j = 0; j <= m; ++j) d[0][j] = j;
  for (uint64_t i = 1; i <= n; ++i)
    for (uint64_t j = 1; j <= m; ++j) {
      uint64_t cost = ((uint64_t)a[i - 1] == (uint64_t)b[j - 1]) ? 0 : 1;
      uint64_t v = std::min({d[i - 1][j] + 1, d[i][j - 1] + 1, d[i - 1][j - 1] + cost});
      if (i > 1 && j > 1 && (uint64_t)a[i - 1] == (uint64_t)b[j - 2] && (uint64_t)a[i - 2] == (uint64_t)b[j - 1])
        v = std::min(v, d[i - 2][j - 2] + 1);
      d[i][j] = v;
    }
  return d[n][m];
}

// Textbook Jaro: window = max(len)/2 - 1, greedy first-unmatched matching, transpositions/2.
template <class C1, class C2>
double jaro(const C1* a, uint64_t n, const C2* b, uint64_t m) {
  if (n == 0 && m == 0) return 1.0;
  if (n == 0 || m == 0) return 0.0;
  if (n == 1 && m == 1) return ((uint64_t)a[0] == (uint64_t)b[0]) ? 1.0 : 0.0;
  int64_t window = (int64_t)(std::max(n, m) / 2) - 1;
  if (window < 0) window = 0;
  std::vector<char> ma(n, 0), mb(m, 0);
  uint64_t matches = 0;
  for (uint64_t j = 0; j < m; ++j) {
    int64_t lo = std::max<int64_t>(0, (int64_t)j - window);
    int64_t hi = std::min<int64_t>((int64_t)n - 1, (int64_t)j + window);
    for (int64_t i = lo; i <= hi; ++i) {
      if (!ma[i] && (uint64_t)a[i] == (uint64_t)b[j]) { ma[i] = 1; mb[j] = 1; ++matches; break; }
    }
  }
  if (matches == 0) return 0.0;
  uint64_t t = 0, k = 0;
  for (uint64_t j = 0; j < m; ++j) {
    if (!mb[j]) continue;
    while (!ma[k]) ++k;
    if ((uint64_t)a[k] != (uint64_t)b[j]) ++t;
    ++k;
  }
  t /= 2;
  double c = (double)matches;
  return (c / (double)n + c / (double)m + (c - (double)t) / c) / 3.0;
}

template <class C1, class C2>
double jaro_winkler(const C1* a, uint64_t n, const C2* b, uint64_t m, double w = 0.1) {
  double sim = jaro(a, n, b, m);
  uint64_t p = 0;
  while (p < 4 && p < n && p < m && (uint64_t)a[p] == (uint64_t)b[p]) ++p;
  if (sim > 0.7) sim += (double)p * w * (1.0 - sim);
  return sim;
}

// Unrestricted Damerau-Levenshtein, Lowrance & Wagner's full-matrix algorithm (adjacent transpositions with
// arbitrary edits in between): independent of the Zhao/Sahni linear-space version the reference uses.
template <class C1, class C2>
uint64_t damerau_levenshtein(const C1* a, uint64_t n, const C2* b, uint64_t m) {
  const uint64_t INF = n + m;
  std::vector<std::vector<uint64_t>> H(n + 2, std::vector<uint64_t>(m + 2, 0));
  std::map<uint64_t, uint64_t> da;
  H[0][0] = INF;
  for (uint64_t i = 0; i <= n; ++i) { H[i + 1][0] = INF; H[i + 1][1] = i; }
  for (uint64_t j = 0; j <= m; ++j) { H[0][j + 1] = INF; H[1][j + 1] = j; }
  for (uint64_t i = 1; i <= n; ++i) {
    uint64_t db = 0;
    for (uint64_t j = 1; j <= m; ++j) {
      const uint64_t i1 = da.count((uint64_t)b[j - 1]) ? da[(uint64_t)b[j - 1]] : 0, j1 = db;
      uint64_t cost = 1;
      if ((uint64_t)a[i - 1] == (uint64_t)b[j - 1]) { cost = 0; db = j; }
      uint64_t v = H[i][j] + cost;
      v = std::min(v, H[i + 1][j] + 1);
      v = std::min(v, H[i][j + 1] + 1);
      v = std::min(v, H[i1][j1] + (i - i1 - 1) + 1 + (j - j1 - 1));
      H[i + 1][j + 1] = v;
    }
    da[(uint64_t)a[i - 1]] = i;
  }
  return H[n + 1][m + 1];
}

}  // namespace rftb
