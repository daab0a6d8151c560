// Host-only part of the C++ mirror (no GPU needed): element widening by value; the GPU calls are instantiated, not run.
#include "rapidfuzz_b200.hpp"
using namespace rapidfuzz_b200;
int main() {
  const int16_t q[3] = {-1, 7, -1};
  const uint16_t c[3] = {65535, 7, 65535};
  const uint64_t off[2] = {0, 3};
  auto w = widen_elements(q, 3);
  if (w[0] != 0xFFFFFFFFu || w[1] != 7) return 1;
  auto w2 = widen_elements(c, 3);
  if (w2[0] != 65535u) return 2;
  const int64_t big[1] = {1ll << 40};
  bool threw = false;
  try { widen_elements(big, 1); } catch (const Error&) { threw = true; }
  if (!threw) return 3;
  if (false) {  // instantiation only (needs a GPU to run)
    Corpus corp = Corpus::from_elements(c, off, 1);
    distance::levenshtein::BatchComparator b(q, 3);
    (void)b.distance(corp);
  }
  return 0;
}
