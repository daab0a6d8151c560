/* rfsynth.h -- synthetic workload generator (SplitMix64, 62-symbol alphanumeric ASCII,
 * lengths uniform in [min_len,max_len], 1/64 of the candidates = query with <= kmax random edits).
 * Host-side test / bench utility in its own library (synth/librfsynth.so), not part of the product ABI.
 * Returns 0 on success, 1 on invalid arguments.  rf_synth_corpus_u8 writes offsets[n+1] and, if chars != NULL, the
 * bytes; call once with chars == NULL to size the buffer (offsets[n] = total). */
#ifndef RFSYNTH_H
#define RFSYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int rf_synth_query_u8(uint64_t seed, uint32_t len, uint8_t* out);
int rf_synth_corpus_u8(uint64_t seed, const uint8_t* query, uint32_t query_len, uint64_t n, uint32_t min_len,
                       uint32_t max_len, uint32_t kmax, uint64_t* offsets, uint8_t* chars, int nthreads);
#ifdef __cplusplus
}
#endif
#endif
