// Known-answer tests of the C++ host mirror, written like the reference's own tests / doc-tests
// (levenshtein.rs:1378, :1632-1633, :2024-2066; Readme.md:62-106; jaro.rs:1081-1092; fuzz.rs:94-96).
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>
#include "rapidfuzz_b200.hpp"

using namespace rapidfuzz_b200;
static int fails = 0;
#define EXPECT(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

int main() {
  namespace lev = distance;
  EXPECT(distance::levenshtein::distance("CA", "ABC") == 3);
  distance::levenshtein::BatchComparator scorer("CA");
  EXPECT(scorer.distance("ABC") == 3);
  EXPECT(distance::levenshtein::distance("kitten", "sitting") == 3);
  auto none = distance::levenshtein::distance_with_args("kitten", "sitting", Args<uint32_t>{}.score_cutoff(2));
  EXPECT(!none.has_value());
  auto some = distance::levenshtein::distance_with_args("kitten", "sitting", Args<uint32_t>{}.score_cutoff(3));
  EXPECT(some.has_value() && *some == 3);

  std::vector<std::string> cands = {"North Korea", "South Korea", "", "aabc", "cccd", "Korea"};
  Corpus corpus = Corpus::from_strings(cands);
  distance::levenshtein::BatchComparator sk("South Korea");
  auto d = sk.distance(corpus);
  EXPECT(d.size() == 6 && d[0] == 2 && d[1] == 0 && d[2] == 11 && d[5] == 6);
  auto dc = sk.distance_with_args(corpus, Args<uint32_t>{}.score_cutoff(2));
  EXPECT(dc[0].has_value() && *dc[0] == 2 && dc[1].has_value() && !dc[2].has_value() && !dc[5].has_value());
  auto w = sk.distance_with_args(corpus, Args<uint32_t>{}.weights(1, 1, 2));   // levenshtein.rs:2036-2042
  EXPECT(w[0] == 4);
  EXPECT(distance::indel::distance("lewenstein", "levenshtein") == 3);         // indel.rs:119
  EXPECT(distance::lcs_seq::BatchComparator("lewenstein").similarity("levenshtein") == 9);  // lcs_seq.rs:763-764
  EXPECT(distance::osa::distance("CA", "AC") == 1);                            // osa.rs:678
  EXPECT(std::fabs(distance::jaro::similarity("james", "robert") - 0.455556) < 1e-4);        // jaro.rs:1081-1086
  EXPECT(std::fabs(distance::jaro_winkler::similarity("aaaaaaaa", "aabaaab") - 0.82381) < 1e-4);  // jaro_winkler.rs:694-798
  EXPECT(std::fabs(fuzz::ratio("this is a test", "this is a test!") - 0.9655172) < 1e-6);   // fuzz.rs:94-96
  auto ns = sk.normalized_similarity_with_args(corpus, Args<double>{}.score_cutoff(0.5));
  EXPECT(ns[0].has_value() && std::fabs(*ns[0] - (1.0 - 2.0 / 11.0)) < 1e-12 && !ns[2].has_value());
  // post-processing on the GPU + streaming from host memory
  auto top = sk.extract<uint32_t>(corpus, RF_DISTANCE, 3, Args<uint32_t>{});
  EXPECT(top.size() == 3 && top[0].index == 1 && top[0].score == 0 && top[1].index == 0 && top[1].score == 2 &&
         top[2].index == 5 && top[2].score == 6);
  auto within = sk.filter<uint32_t>(corpus, RF_DISTANCE, Args<uint32_t>{}.score_cutoff(6));
  EXPECT(within.size() == 3 && within[0].index == 0 && within[1].index == 1 && within[2].index == 5 && within[2].score == 6);
  auto best_sim = sk.extract<double>(corpus, RF_NORMALIZED_SIMILARITY, 1, Args<double>{});
  EXPECT(best_sim.size() == 1 && best_sim[0].index == 1 && best_sim[0].score == 1.0);
  {
    std::vector<uint8_t> chars;
    std::vector<uint64_t> offsets{0};
    for (const auto& s : cands) { chars.insert(chars.end(), s.begin(), s.end()); offsets.push_back(chars.size()); }
    auto ds = sk.stream<uint32_t>(chars.data(), offsets.data(), cands.size(), RF_DISTANCE, Args<uint32_t>{});
    EXPECT(ds == d);
  }
  {  // u32 elements: the reference's unicode test (levenshtein.rs:2164-2169)
    const std::u32string a = U"\u0418\u0432\u0430\u043d\u043a\u043e", b = U"\u041f\u0435\u0442\u0440\u0443\u043d\u043a\u043e";
    const uint64_t off[2] = {0, b.size()};
    Corpus cu = Corpus::from_u32(reinterpret_cast<const uint32_t*>(b.data()), off, 1);
    EXPECT(distance::levenshtein::BatchComparator(std::u32string_view(a)).distance(cu)[0] == 5);
  }
  EXPECT(distance::hamming::distance("hamming", "humming") == 1);                     // hamming.rs:198
  EXPECT(distance::hamming::distance_with_args("ham", "hamming", Args<uint32_t>{}.pad(true)) == 4);  // :622-625
  EXPECT(distance::damerau_levenshtein::distance("CA", "ABC") == 2);                  // damerau_levenshtein.rs:226
  EXPECT(distance::prefix::similarity("prefix", "preference") == 4);                  // prefix.rs:122
  EXPECT(distance::postfix::similarity("postfix", "prefix") == 3);                    // postfix.rs:122
  bool threw = false;
  try { distance::hamming::distance("ham", "hamming"); } catch (const Error& e) { threw = e.status == RF_ERR_INVALID_ARG; }
  EXPECT(threw);                                                                        // hamming.rs:617-620
  threw = false;
  try { distance::levenshtein::BatchComparator too_long(std::string(RF_MAX_QUERY_LEN + 1, 'a')); } catch (const Error& e) { threw = e.status == RF_ERR_UNSUPPORTED; }
  EXPECT(threw);
  auto wg = sk.distance_with_args(corpus, Args<uint32_t>{}.weights(1, 2, 3));   // generic weights: Wagner-Fischer route
  EXPECT(wg[1] == 0 && wg[0] == 6 && wg[2] == 22);
  {  // many-vs-many top-k: ties on the distance go to the smaller index
    auto tk = process::cdist_topk(std::vector<std::string>{"South Korea", "aabd", ""}, corpus, 2);
    EXPECT(tk.nq == 3 && tk.at(0, 0)->index == 1 && tk.at(0, 0)->score == 0 && tk.at(0, 1)->index == 0 && tk.at(0, 1)->score == 2);
    EXPECT(tk.at(1, 0)->index == 3 && tk.at(1, 0)->score == 1 && tk.at(2, 0)->index == 2 && tk.at(2, 0)->score == 0);
    auto tc = process::cdist_topk(std::vector<std::string>{"aabd"}, corpus, 3, Args<uint32_t>{}.score_cutoff(1));
    EXPECT(tc.at(0, 0).has_value() && !tc.at(0, 1).has_value());
  }
  {  // one process, the corpus split over a device list (the same device three times on a one-GPU box): identical results
    const std::vector<int> devs{0, 0, 0};
    sharded::Corpus sc(cands, devs);
    sharded::BatchComparator<RF_LEVENSHTEIN> ssk("South Korea", devs);
    EXPECT(sc.size() == cands.size() && ssk.distance(sc) == d);
    auto stk = sharded::cdist_topk(std::vector<std::string>{"South Korea", "aabd"}, sc, 2);
    EXPECT(stk.index[0] == 1 && stk.distance[0] == 0 && stk.index[1] == 0 && stk.distance[1] == 2 && stk.index[2] == 3 && stk.distance[2] == 1);
  }
  std::printf(fails ? "cpp api: %d failure(s)\n" : "cpp api: all ok\n", fails);
  return fails ? 1 : 0;
}
