/* rfgpu.h -- C ABI of the B200-native one-vs-many fuzzy-string scoring engine (librfgpu.so).
 *
 * Drop-in boundary for the `BatchComparator` hot path of rapidfuzz-rs 0.5.0.  The reference has no FFI
 * (pure safe Rust, src/lib.rs:78); every entry point below names the reference interface it replaces
 * (paths relative to the reference's src/).  A Rust shim binding these symbols is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers + sizes, no C++/torch types; every function returns rf_status (0 = ok), never aborts.
 *  - `None` (score worse than score_cutoff: common.rs:43-45, :83-85) is UINT32_MAX in u32 outputs and NaN in
 *    f64 outputs.
 *  - handles are immutable after creation; concurrent calls on the same handles are safe
 *    (BatchComparator is Clone + Send + Sync in the reference: levenshtein.rs:1635-1639).
 *  - `*_device` variants take/return device pointers (e.g. torch tensor data_ptr()) and a cudaStream_t
 *    passed as void*; they enqueue work and return without synchronising.
 *  - there is NO CPU fallback: without a usable CUDA device every compute entry point returns RF_ERR_CUDA.
 */
#ifndef RFGPU_H
#define RFGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rf_status {
  RF_OK = 0,
  RF_ERR_INVALID_ARG = 1,
  RF_ERR_UNSUPPORTED = 2, /* e.g. query longer than RF_MAX_QUERY_LEN; generic (non-uniform, non-indel) Levenshtein weights
                             with a query longer than 200 000; 64-bit element values that do not fit the 32-bit symbol domain */
  RF_ERR_CUDA = 3,
  RF_ERR_OOM = 4,
  RF_ERR_NCCL = 5 /* a collective of the sharded (multi-GPU) entry points failed */
} rf_status;

/* metric modules: distance/{levenshtein,indel,lcs_seq,osa,jaro,jaro_winkler,hamming,prefix,postfix,damerau_levenshtein}.rs and
 * fuzz.rs (ratio) */
typedef enum rf_metric {
  RF_LEVENSHTEIN = 0,
  RF_INDEL = 1,
  RF_LCS_SEQ = 2,
  RF_OSA = 3,
  RF_JARO = 4,
  RF_JARO_WINKLER = 5,
  RF_RATIO = 6,
  RF_HAMMING = 7, /* hamming.rs:136-199; see rf_args.pad */
  RF_PREFIX = 8,  /* prefix.rs:47-71: similarity = common prefix length */
  RF_POSTFIX = 9, /* postfix.rs:47-71: similarity = common suffix length */
  RF_DAMERAU_LEVENSHTEIN = 10 /* damerau_levenshtein.rs:111-214 (unrestricted; query <= 200 000 elements) */
} rf_metric;

/* which BatchComparator method: distance / similarity / normalized_distance / normalized_similarity */
typedef enum rf_kind {
  RF_DISTANCE = 0,
  RF_SIMILARITY = 1,
  RF_NORMALIZED_DISTANCE = 2,
  RF_NORMALIZED_SIMILARITY = 3
} rf_kind;

/* POD image of the per-metric `Args` builders (levenshtein.rs:86-126, jaro_winkler.rs:25-62, ...).
 * cutoff_u/hint_u are used by integer-valued results, cutoff_f/hint_f by float-valued ones.
 * score_hint only steers the reference's band search (levenshtein.rs:1069-1088); accepted and ignored. */
typedef struct rf_args {
  uint8_t has_cutoff;
  uint64_t cutoff_u;
  double cutoff_f;
  uint8_t has_hint;
  uint64_t hint_u;
  double hint_f;
  uint64_t insertion_cost, deletion_cost, substitution_cost; /* WeightTable, levenshtein.rs:130-148 */
  double prefix_weight;                                      /* jaro_winkler.rs:31-39, default 0.1 */
  uint8_t reference_quirks; /* 1: RatioBatchComparator divides by max(len1,len2) like fuzz.rs:141 */
  uint8_t pad;              /* hamming::Args::pad (hamming.rs:112-118).  0 (default): a candidate whose length differs
                             * from the query's is Err(DifferentLengthArgs) (hamming.rs:232-234): its result is the None
                             * sentinel and the host-buffer entry points return RF_ERR_INVALID_ARG ("Differing length
                             * arguments provided") after filling the output; the *_device / extract / filter / stream
                             * entry points only write the sentinel.  1: the excess length counts as mismatches. */
} rf_args;

/* Queries may have up to RF_MAX_QUERY_LEN elements (the reference has no limit; its longest own test is 106 514 x
 * 107 244, levenshtein.rs:2139-2161).  The bound only keeps the per-query match tables (2 x 32 B per element) sane. */
#define RF_MAX_QUERY_LEN 4194304u

typedef struct rf_corpus rf_corpus; /* packed candidates resident in one GPU's HBM */
typedef struct rf_batch rf_batch;   /* one cached query == one BatchComparator */

void rf_args_default(rf_args* a); /* Args::default(): no cutoff, no hint, weights (1,1,1), prefix_weight 0.1 */
const char* rf_status_string(rf_status s);
const char* rf_last_error(void); /* thread-local detail of the last non-OK status */
int rf_device_count(void);       /* 0 when no CUDA device is usable */

/* ---- corpus: the candidates the reference receives one iterator at a time
 * (BatchComparator::distance(s2), levenshtein.rs:1740-1777).  chars = concatenated u8 elements,
 * offsets[n+1] = CSR starts (offsets[0] = 0).  Host buffers are copied; caller keeps ownership. */
rf_status rf_corpus_create_u8(const uint8_t* chars, const uint64_t* offsets, uint64_t n, int device,
                              rf_corpus** out);
/* same, CSR starts given as u32 (total bytes < 2^32) -- halves the offset upload */
rf_status rf_corpus_create_u8_off32(const uint8_t* chars, const uint32_t* offsets, uint64_t n, int device,
                                    rf_corpus** out);
/* buffers already on `device` (copied device-to-device on `stream`) */
rf_status rf_corpus_create_device_u8(const uint8_t* d_chars, const uint64_t* d_offsets, uint64_t n,
                                     uint64_t total_chars, int device, void* stream, rf_corpus** out);
/* u32 elements (Rust `char`, u32: the reference accepts any HashableChar, details/common.rs:29-37).  Such a
 * corpus is scored by comparators made with rf_batch_create_u32; results are exact (see there).  A corpus whose
 * elements take at most 255 distinct values is renamed to bytes once, here, and then costs and scores like a u8
 * corpus (rf_set_option("compact_u32_corpus", 0) turns that off).
 * Other element widths (u16, i8 ... i32, and 64-bit elements whose values fit): the reference compares elements
 * numerically (HashableChar::hash_char), so widen them to u32 BY VALUE -- zero-extend unsigned types, keep the
 * two's-complement 32-bit pattern of negative values -- on both the query and the candidate side (what the Python
 * mirror's widen_elems does); do not mix negative values with unsigned values >= 2^31 in one comparison. */
rf_status rf_corpus_create_u32(const uint32_t* elems, const uint64_t* offsets, uint64_t n, int device, rf_corpus** out);
/* Any element type of the reference (HashableChar: u8..u64, i8..i64, char = u32; details/common.rs:29-37).  Elements are
 * compared BY VALUE like hash_char does: the library widens them to the u32 domain it scores in (unsigned zero-extended,
 * signed as the two's-complement pattern of the value).  64-bit values outside [-2^31, 2^32) -> RF_ERR_UNSUPPORTED.  A
 * comparison that would pair negative signed elements on one side with unsigned elements >= 2^31 on the other (equal
 * 32-bit patterns, never equal in the reference: Hash::SIGNED vs Hash::UNSIGNED) is refused with RF_ERR_UNSUPPORTED at
 * scoring time instead of being answered wrongly.  RF_ELEM_U8 is the plain byte path. */
typedef enum rf_elem_type {
  RF_ELEM_U8 = 0, RF_ELEM_U16 = 1, RF_ELEM_U32 = 2, RF_ELEM_U64 = 3,
  RF_ELEM_I8 = 4, RF_ELEM_I16 = 5, RF_ELEM_I32 = 6, RF_ELEM_I64 = 7
} rf_elem_type;
rf_status rf_corpus_create_elems(const void* elems, rf_elem_type type, const uint64_t* offsets, uint64_t n, int device,
                                 rf_corpus** out);
/* Waits for the device to drain first (asynchronous *_device calls may still be reading the corpus). */
rf_status rf_corpus_destroy(rf_corpus* c);
/* A corpus holds its candidates twice: as the CSR copy it was given and as the length-bucketed interleaved layout the
 * single-word kernels read, about 2.2x the candidates' bytes in all.  rf_corpus_release_csr frees the
 * CSR copy (45 % of the footprint) for hosts that only need what the layout serves: Levenshtein / OSA / Indel / LCSseq /
 * ratio with queries of at most 512 elements, Jaro / Jaro-Winkler up to 64, Hamming, Damerau-Levenshtein and generic weights
 * up to 64 (candidates up to 32 000 elements), extract / filter on those, and rf_cdist_topk_*.  Everything else (longer
 * queries, Prefix / Postfix, u32 queries that need a renaming pass, the small-cutoff band kernel -- such calls take the
 * register-column kernel instead, same results) returns RF_ERR_UNSUPPORTED afterwards.  Not reversible; waits for the device
 * to drain; must not run concurrently with calls on the same corpus. */
rf_status rf_corpus_release_csr(rf_corpus* c);
int rf_corpus_has_csr(const rf_corpus* c); /* 0 after rf_corpus_release_csr */
uint64_t rf_corpus_size(const rf_corpus* c);        /* number of candidates */
uint64_t rf_corpus_total_chars(const rf_corpus* c); /* sum of candidate lengths */
int rf_corpus_device(const rf_corpus* c);

/* ---- batch comparator: `BatchComparator::new(query)` -- keeps s1 and builds the pattern-match bit table
 * (levenshtein.rs:1645-1657, pattern_match_vector.rs:203-281). */
rf_status rf_batch_create_u8(rf_metric metric, const uint8_t* query, uint32_t query_len, int device,
                             rf_batch** out);
/* u32-element query.  All metrics of this library depend only on which query / candidate positions hold equal symbols, so
 * renaming symbols is exact (scores identical to the reference's hashmap-based lookup, pattern_match_vector.rs:20-65,
 * :226-280).  Against byte candidates (u8 corpora, the byte streaming entry points) and against u32 corpora that were
 * renamed to bytes at creation, the QUERY is mapped into the candidates' symbol domain (a symbol they cannot contain
 * matches nothing): no extra pass, any number of distinct query symbols.  Against a u32 corpus with more than 255
 * distinct symbols the candidates are renamed to the query's own alphabet on the device per call (one extra pass): to
 * bytes when the query has at most 255 distinct symbols, to 16-bit codes otherwise (at most 65 535 distinct symbols; the
 * multi-word kernels over uint16_t elements with one match-table row per code).  The metrics that compare symbols
 * directly (Hamming / Prefix / Postfix / Damerau-Levenshtein / generic weights) need a query of at most 255 distinct
 * symbols (else RF_ERR_UNSUPPORTED at scoring time). */
rf_status rf_batch_create_u32(rf_metric metric, const uint32_t* query, uint32_t query_len, int device, rf_batch** out);
/* query of any element type (see rf_corpus_create_elems); the comparator scores u8 and u32 / typed corpora alike */
rf_status rf_batch_create_elems(rf_metric metric, const void* query, rf_elem_type type, uint32_t query_len, int device,
                                rf_batch** out);
rf_status rf_batch_destroy(rf_batch* b);
/* Kernel-choice knobs of ONE comparator ("single_word_path", "multi_word_path", "banded_levenshtein", "jaro32", "epilogue_table"; see
 * rf_set_option for their meaning).  A comparator copies the process-wide defaults when it is created; scoring calls
 * read the comparator's copy only, so threads that want different kernels do not race on global state.  Call it
 * before the comparator is shared between threads. */
rf_status rf_batch_set_option(rf_batch* b, const char* name, int value);

/* ---- scoring: one call == the user's loop `for c in candidates { scorer.<kind>_with_args(c, &args) }`.
 * Integer-valued: levenshtein/indel/lcs_seq/osa distance|similarity            -> u32 out[n]
 * Float-valued:   every normalized_*; jaro/jaro_winkler distance|similarity; ratio similarity -> f64 out[n]
 * (levenshtein.rs:1660-1817, lcs_seq.rs:796-949, indel.rs:371-520, osa.rs:463-616, jaro.rs:826-979,
 *  jaro_winkler.rs:409-578, fuzz.rs:102-150).  args == NULL means Args::default(). */
rf_status rf_batch_score_u32(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args,
                             uint32_t* out_host);
rf_status rf_batch_score_f64(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args,
                             double* out_host);
rf_status rf_batch_score_u32_device(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args,
                                    uint32_t* out_device, void* stream);
rf_status rf_batch_score_f64_device(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args,
                                    double* out_device, void* stream);
/* Integer-valued results as BYTES (None = 0xFF): a quarter of the result bytes to download, for corpora and queries whose
 * scores stay below 255 (the result download of rf_batch_score_u32 on BASELINE config 2 takes 3.5x as long as the scan).
 * A score above 254 does not fit: such entries read 0xFF and the host-output call returns RF_ERR_INVALID_ARG after filling
 * the output (the *_device call, which does not synchronise, only writes 0xFF).  Float-valued (metric, kind) pairs ->
 * RF_ERR_INVALID_ARG. */
rf_status rf_batch_score_u8(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint8_t* out_host);
rf_status rf_batch_score_u8_device(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args,
                                   uint8_t* out_device, void* stream);
/* 1 if (metric, kind) yields f64, 0 if u32 */
int rf_result_is_float(rf_metric metric, rf_kind kind);

/* named wrappers, one per reference method */
rf_status rf_batch_distance_u32(const rf_batch* b, const rf_corpus* c, const rf_args* args, uint32_t* out);
rf_status rf_batch_similarity_u32(const rf_batch* b, const rf_corpus* c, const rf_args* args, uint32_t* out);
rf_status rf_batch_distance_f64(const rf_batch* b, const rf_corpus* c, const rf_args* args, double* out);
rf_status rf_batch_similarity_f64(const rf_batch* b, const rf_corpus* c, const rf_args* args, double* out);
rf_status rf_batch_normalized_distance_f64(const rf_batch* b, const rf_corpus* c, const rf_args* args, double* out);
rf_status rf_batch_normalized_similarity_f64(const rf_batch* b, const rf_corpus* c, const rf_args* args, double* out);

/* ---- scoring + post-processing on the device (new on this side; what Python rapidfuzz calls process.extract --
 * the Rust crate leaves the sort / cutoff collection to the caller's loop).  Only the selected (index, score) pairs
 * cross PCIe instead of n scores.
 *   extract: the k (<= 1024) best candidates by (score best-first, index ascending); "best" = smallest for the
 *            distance kinds, largest for the similarity kinds; None scores (args->score_cutoff) never qualify;
 *            *n_out = entries written (< k when fewer candidates qualify).
 *   filter:  every candidate whose score is not None (i.e. passed args->score_cutoff), in index order;
 *            *n_hits = their total number, of which the first min(n_hits, capacity) are written.
 * Host output buffers; u32 / f64 as for rf_batch_score_*. */
rf_status rf_batch_extract_u32(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint32_t k,
                               uint32_t* idx_out, uint32_t* score_out, uint32_t* n_out);
rf_status rf_batch_extract_f64(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint32_t k,
                               uint32_t* idx_out, double* score_out, uint32_t* n_out);
rf_status rf_batch_filter_u32(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint64_t capacity,
                              uint32_t* idx_out, uint32_t* score_out, uint64_t* n_hits);
rf_status rf_batch_filter_f64(const rf_batch* b, const rf_corpus* c, rf_kind kind, const rf_args* args, uint64_t capacity,
                              uint32_t* idx_out, double* score_out, uint64_t* n_hits);

/* ---- streaming: the same loop when the candidates live in HOST memory and are not kept on the GPU
 * (the literal shape of `for c in candidates { scorer.distance(c) }`, levenshtein.rs:1740-1777 / bench_levenshtein.rs:51-58).
 * The CSR corpus is cut into chunks; H2D copy, scan and result D2H of successive chunks overlap on separate
 * streams, so a call costs about (total bytes + 4..8 B offsets per candidate) / PCIe bandwidth.  Pinned
 * (page-locked) chars/offsets/out buffers give full link speed; pageable buffers work, slower.  Results and
 * errors are identical to rf_corpus_create_* + rf_batch_score_*.  Tunables: rf_set_option("stream_chunk_mb" |
 * "stream_chunk_kcand"). */
rf_status rf_batch_stream_u32(const rf_batch* b, const uint8_t* chars, const uint64_t* offsets, uint64_t n,
                              rf_kind kind, const rf_args* args, uint32_t* out_host);
rf_status rf_batch_stream_u32_off32(const rf_batch* b, const uint8_t* chars, const uint32_t* offsets, uint64_t n,
                                    rf_kind kind, const rf_args* args, uint32_t* out_host);
rf_status rf_batch_stream_f64(const rf_batch* b, const uint8_t* chars, const uint64_t* offsets, uint64_t n,
                              rf_kind kind, const rf_args* args, double* out_host);
rf_status rf_batch_stream_f64_off32(const rf_batch* b, const uint8_t* chars, const uint32_t* offsets, uint64_t n,
                                    rf_kind kind, const rf_args* args, double* out_host);

/* host-resident candidates with u32 elements (comparator from rf_batch_create_u32 / rf_batch_create_elems): a chunk
 * crosses the link as 4-byte elements and is renamed to the query's byte alphabet on the device */
rf_status rf_batch_stream_u32_elems32(const rf_batch* b, const uint32_t* elems, const uint64_t* offsets, uint64_t n, rf_kind kind,
                                      const rf_args* args, uint32_t* out_host);
rf_status rf_batch_stream_f64_elems32(const rf_batch* b, const uint32_t* elems, const uint64_t* offsets, uint64_t n, rf_kind kind,
                                      const rf_args* args, double* out_host);
/* Fewer bytes on the link (the call is PCIe-bound): candidates of at most 255 elements described by ONE length byte each
 * instead of a CSR start (lens[i] = length of candidate i, chars = the candidates back to back; the starts are rebuilt
 * on the device by a prefix sum), integer-valued (metric, kind) only.  _u8_: the scores come back as one byte each,
 * None = 0xFF; a score above 254 anywhere makes the call return RF_ERR_INVALID_ARG (use the _u32_ variant then).
 * Chunk boundaries fall on multiples of 4096 candidates. */
rf_status rf_batch_stream_u32_len8(const rf_batch* b, const uint8_t* chars, const uint8_t* lens, uint64_t n, rf_kind kind,
                                   const rf_args* args, uint32_t* out_host);
rf_status rf_batch_stream_u8_len8(const rf_batch* b, const uint8_t* chars, const uint8_t* lens, uint64_t n, rf_kind kind,
                                  const rf_args* args, uint8_t* out_host);

/* ... and fewer still when the corpus has at most 64 distinct symbols (ASCII alphanumerics: 62): rf_pack6_u8 packs the
 * concatenated candidates ONCE on the host, 4 characters into 3 bytes (packed_out: rf_pack6_size(total) bytes; dict_out:
 * 64 bytes, code -> symbol; more than 64 symbols -> RF_ERR_UNSUPPORTED); the _packed6 streaming entry points send the
 * packed stream (27 instead of 36 bytes per config-2 pair) and unpack each chunk on the device before the scan.  Same
 * results, same error behaviour as the _len8 entry points. */
uint64_t rf_pack6_size(uint64_t total_chars);
rf_status rf_pack6_u8(const uint8_t* chars, uint64_t total_chars, uint8_t* packed_out, uint8_t* dict_out, int nthreads);
rf_status rf_batch_stream_u32_len8_packed6(const rf_batch* b, const uint8_t* packed, const uint8_t* dict64, const uint8_t* lens,
                                           uint64_t n, rf_kind kind, const rf_args* args, uint32_t* out_host);
rf_status rf_batch_stream_u8_len8_packed6(const rf_batch* b, const uint8_t* packed, const uint8_t* dict64, const uint8_t* lens,
                                          uint64_t n, rf_kind kind, const rf_args* args, uint8_t* out_host);

/* ---- many-vs-many (new on this side; the reference has no cdist -- SURVEY fact 3): for each of nq queries
 * the k best candidates by (distance ascending, index ascending); fewer than k hits are padded with
 * (UINT32_MAX, UINT32_MAX).  Levenshtein distance (unit weights), queries of length <= 64, k <= 64; with
 * args->has_cutoff only candidates with distance <= cutoff_u qualify.  Queries are host buffers (CSR like the
 * corpus); idx/dist are [nq][k]; the _device variant writes device buffers and returns after the work on
 * `stream` has completed (the per-call scratch is freed on return). */
rf_status rf_cdist_topk_u8(const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c,
                           const rf_args* args, uint32_t k, uint32_t* idx_host, uint32_t* dist_host);
rf_status rf_cdist_topk_u8_device(const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq,
                                  const rf_corpus* c, const rf_args* args, uint32_t k, uint32_t* idx_device,
                                  uint32_t* dist_device, void* stream);

/* the same for the other bit-parallel distances: metric = RF_LEVENSHTEIN | RF_OSA | RF_INDEL | RF_LCS_SEQ (the k smallest
 * DISTANCES of that metric, ties by index; every one of them is bounded below by the length difference, so the length
 * skip of the scan stays exact) */
rf_status rf_cdist_topk_metric_u8(rf_metric metric, const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c,
                                  const rf_args* args, uint32_t k, uint32_t* idx_host, uint32_t* dist_host);
rf_status rf_cdist_topk_metric_u8_device(rf_metric metric, const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq,
                                         const rf_corpus* c, const rf_args* args, uint32_t k, uint32_t* idx_device,
                                         uint32_t* dist_device, void* stream);
/* u32-element queries against a corpus made by rf_corpus_create_u32 that was renamed to bytes at creation (at most 255
 * distinct symbols; a larger alphabet -> RF_ERR_UNSUPPORTED), or against a u8 corpus when every query symbol is a byte */
rf_status rf_cdist_topk_u32(const uint32_t* q_elems, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c,
                            const rf_args* args, uint32_t k, uint32_t* idx_host, uint32_t* dist_host);
rf_status rf_cdist_topk_u32_device(const uint32_t* q_elems, const uint64_t* q_offsets, uint32_t nq, const rf_corpus* c,
                                   const rf_args* args, uint32_t k, uint32_t* idx_device, uint32_t* dist_device, void* stream);

/* Sharded corpora (one contiguous candidate range per GPU): global top-k from the per-shard
 * lists after ONE all-gather.  Part p's rows are idx_parts / dist_parts + p * part_stride, each [nq][k] as written
 * by rf_cdist_topk_u8_device on shard p (shard-local indices, UINT32_MAX padding); index_base_device[p] is the
 * global index of shard p's first candidate (ascending in p).  Output rows: the k best by (distance, global
 * index), padded with (UINT64_MAX, UINT32_MAX).  All pointers are device memory of `device`; asynchronous on
 * `stream`.  parts * k <= 25600. */
rf_status rf_topk_merge_device(const uint32_t* idx_parts, const uint32_t* dist_parts, uint64_t part_stride,
                               const uint64_t* index_base_device, uint32_t parts, uint32_t nq, uint32_t k,
                               uint64_t* idx_out_device, uint32_t* dist_out_device, int device, void* stream);

/* ---- sharded corpora: ONE process, several GPUs of one box.  The reference's BatchComparator is
 * plain data, Clone + Send + Sync (levenshtein.rs:1635-1639), i.e. a Rust host may drive it from any thread over any
 * slice of the candidates; here the split is the library's: contiguous candidate ranges balanced by BYTES, one
 * resident shard per listed device (uploaded and laid out concurrently), the query's tables replicated, per-device
 * streams.  Pairs are independent, so the scan has no exchange step; the collectives are the final ones only:
 *   - *_allgather_device: NCCL all-gather of the per-shard score vectors (grouped ncclBroadcast = all-gather-v, in
 *     place), every device ends with all n scores in candidate order;
 *   - rf_sharded_cdist_topk_u8: ncclAllGather of the per-shard [nq][k] lists + rf_topk_merge_device.
 * Host-destined results (rf_sharded_score_*, rf_sharded_extract_*) need no collective: every device writes its slice of
 * the caller's vector.  Results are identical to the single-GPU entry points on the whole corpus (global candidate
 * indices, u64).  A failing collective returns RF_ERR_NCCL.  `devices` may list a device more than once (one-GPU test
 * boxes): such a list cannot form an NCCL communicator and the same gathers run as event-ordered device copies
 * (rf_set_option("sharded_collective", 1) forces that path for any list; rf_sharded_corpus_uses_nccl tells which).
 * Handles are immutable; concurrent calls are safe (collectives on one corpus are serialised internally). */
typedef struct rf_sharded_corpus rf_sharded_corpus;
typedef struct rf_sharded_batch rf_sharded_batch;
rf_status rf_corpus_create_sharded_u8(const uint8_t* chars, const uint64_t* offsets, uint64_t n, const int* devices, int ndev,
                                      rf_sharded_corpus** out);
rf_status rf_sharded_corpus_destroy(rf_sharded_corpus* c);
uint64_t rf_sharded_corpus_size(const rf_sharded_corpus* c);
int rf_sharded_corpus_shards(const rf_sharded_corpus* c);
/* candidates [*first, *end) live on shard `shard` */
rf_status rf_sharded_corpus_shard_range(const rf_sharded_corpus* c, int shard, uint64_t* first, uint64_t* end);
/* the shard as a plain corpus (owned by the sharded handle), usable with every single-GPU entry point */
const rf_corpus* rf_sharded_corpus_shard(const rf_sharded_corpus* c, int shard);
int rf_sharded_corpus_uses_nccl(const rf_sharded_corpus* c);
/* BatchComparator::new(query), replicated on every listed device (same list, same order, as the corpus) */
rf_status rf_sharded_batch_create_u8(rf_metric metric, const uint8_t* query, uint32_t query_len, const int* devices, int ndev,
                                     rf_sharded_batch** out);
rf_status rf_sharded_batch_create_u32(rf_metric metric, const uint32_t* query, uint32_t query_len, const int* devices, int ndev,
                                      rf_sharded_batch** out);
rf_status rf_sharded_batch_destroy(rf_sharded_batch* b);
/* == rf_batch_score_* over the whole sharded corpus; out_host[n] in candidate order */
rf_status rf_sharded_score_u32(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind, const rf_args* args,
                               uint32_t* out_host);
rf_status rf_sharded_score_f64(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind, const rf_args* args,
                               double* out_host);
/* scores + all-gather: out_device[i] is a buffer of n results on devices[i]; all of them hold every score on return */
rf_status rf_sharded_score_u32_allgather_device(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind,
                                                const rf_args* args, uint32_t* const* out_device);
rf_status rf_sharded_score_f64_allgather_device(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind,
                                                const rf_args* args, double* const* out_device);
/* == rf_batch_extract_* over the whole sharded corpus (global u64 indices) */
rf_status rf_sharded_extract_u32(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind, const rf_args* args, uint32_t k,
                                 uint64_t* idx_out, uint32_t* score_out, uint32_t* n_out);
rf_status rf_sharded_extract_f64(const rf_sharded_batch* b, const rf_sharded_corpus* c, rf_kind kind, const rf_args* args, uint32_t k,
                                 uint64_t* idx_out, double* score_out, uint32_t* n_out);
/* == rf_cdist_topk_u8 over the whole sharded corpus: [nq][k] global indices (UINT64_MAX = none) and distances
 * (UINT32_MAX = none), host buffers.  shards * k <= 25600. */
rf_status rf_sharded_cdist_topk_u8(const uint8_t* q_chars, const uint64_t* q_offsets, uint32_t nq, const rf_sharded_corpus* c,
                                   const rf_args* args, uint32_t k, uint64_t* idx_host, uint32_t* dist_host);
/* == rf_batch_stream_* with the candidate range split by bytes over the batch's devices: every device runs its own
 * chunked H2D / scan / D2H pipeline over its own PCIe link, concurrently */
rf_status rf_sharded_stream_u32(const rf_sharded_batch* b, const uint8_t* chars, const uint64_t* offsets, uint64_t n, rf_kind kind,
                                const rf_args* args, uint32_t* out_host);
rf_status rf_sharded_stream_f64(const rf_sharded_batch* b, const uint8_t* chars, const uint64_t* offsets, uint64_t n, rf_kind kind,
                                const rf_args* args, double* out_host);
/* == rf_batch_stream_*_len8[_packed6] over the batch's devices, with NO static split: the per-device pipelines take their
 * chunks from one shared planner as their slots free up, so a device behind a faster PCIe link takes more chunks and the
 * call runs at the aggregate host-to-device rate of the box instead of waiting for the slowest link. */
rf_status rf_sharded_stream_u32_len8(const rf_sharded_batch* b, const uint8_t* chars, const uint8_t* lens, uint64_t n, rf_kind kind,
                                     const rf_args* args, uint32_t* out_host);
rf_status rf_sharded_stream_u8_len8(const rf_sharded_batch* b, const uint8_t* chars, const uint8_t* lens, uint64_t n, rf_kind kind,
                                    const rf_args* args, uint8_t* out_host);
rf_status rf_sharded_stream_u8_len8_packed6(const rf_sharded_batch* b, const uint8_t* packed, const uint8_t* dict64, const uint8_t* lens,
                                            uint64_t n, rf_kind kind, const rf_args* args, uint8_t* out_host);

/* ---- one process per GPU (MPI / torchrun style hosts): a communicator over the ranks + scoring with the final
 * all-gather of the score vectors (the path's only exchange step, north star: "NCCL all-gather only for the final score
 * vector").  rank 0 calls rf_comm_unique_id and ships the 128 bytes to the other ranks by whatever means the host has;
 * every rank then calls rf_comm_create_rank (collective).  rf_batch_score_*_allgather_device scores this rank's corpus
 * and leaves ALL ranks' results in out_device, rank-major, in candidate order (out_capacity elements available;
 * counts_out[nranks], optional, receives every rank's candidate count).  The transfer is a grouped ncclBroadcast, in place, on an internal
 * high-priority stream ("allgather_chunks" > 1 scans the shard in pieces and sends piece k while piece k+1 is scanned);
 * `stream` is ordered behind the last transfer.  Collective: every rank must
 * make the same call. */
typedef struct rf_comm rf_comm;
rf_status rf_comm_unique_id(void* out128);
rf_status rf_comm_create_rank(const void* id128, int nranks, int rank, int device, rf_comm** out);
rf_status rf_comm_destroy(rf_comm* comm);
int rf_comm_rank(const rf_comm* comm);
int rf_comm_size(const rf_comm* comm);
rf_status rf_batch_score_u32_allgather_device(const rf_batch* b, const rf_corpus* c, rf_comm* comm, rf_kind kind, const rf_args* args,
                                              uint32_t* out_device, uint64_t out_capacity, uint64_t* counts_out, void* stream);
rf_status rf_batch_score_f64_allgather_device(const rf_batch* b, const rf_corpus* c, rf_comm* comm, rf_kind kind, const rf_args* args,
                                              double* out_device, uint64_t out_capacity, uint64_t* counts_out, void* stream);

/* ---- packing and corpus files (host-side; the step before the scoring path).  The reference takes one iterator
 * per candidate (levenshtein.rs:1750-1762); callers holding a Vec<String> pack it once:
 *   rf_pack_u8: n strings given as (pointer, length) -> CSR offsets[n+1] (+ chars[offsets[n]] when chars_out != NULL;
 *               call first with chars_out == NULL to size the buffer).  Parallel copy (nthreads <= 0: all cores).
 *   corpus file: header + offsets (u32 when total < 2^32-16, else u64) + chars, 64-byte aligned sections, mapped
 *               back with mmap -- the accessors' pointers can be passed to rf_corpus_create_u8[_off32] /
 *               rf_batch_stream_* directly; rf_corpus_create_from_file does open + upload + close. */
typedef struct rf_corpus_file rf_corpus_file;
rf_status rf_pack_u8(const uint8_t* const* strings, const uint64_t* lengths, uint64_t n, uint64_t* offsets_out,
                     uint8_t* chars_out, int nthreads);
rf_status rf_corpus_file_write(const char* path, const uint8_t* chars, const uint64_t* offsets, uint64_t n);
rf_status rf_corpus_file_open(const char* path, rf_corpus_file** out);
rf_status rf_corpus_file_close(rf_corpus_file* f);
uint64_t rf_corpus_file_size(const rf_corpus_file* f);
uint64_t rf_corpus_file_total_chars(const rf_corpus_file* f);
uint32_t rf_corpus_file_offset_width(const rf_corpus_file* f); /* 4 or 8: element type of rf_corpus_file_offsets */
const void* rf_corpus_file_offsets(const rf_corpus_file* f);
const uint8_t* rf_corpus_file_chars(const rf_corpus_file* f);
rf_status rf_corpus_create_from_file(const char* path, int device, rf_corpus** out);

/* tuning knobs (process-wide DEFAULTS: "single_word_path", "multi_word_path", "banded_levenshtein", "jaro32" and "epilogue_table" are
 * copied into every comparator at creation -- rf_batch_set_option changes one comparator; "build_interleaved_layout" and
 * "compact_u32_corpus" act at corpus creation; the stream / cdist / sharded knobs are read once at the start of a call):
 *   "build_interleaved_layout" (default 1): corpora created afterwards also keep the length-bucketed,
 *        warp-interleaved copy that the fastest single-word kernel reads (about +1.2x corpus memory);
 *   "single_word_path" (default 0): which kernel scores queries of at most 64 elements:
 *        0 = interleaved layout, per-lane streaming loads with a two-row register pipeline (fastest measured);
 *        1 = CSR kernel (TMA-staged tiles, bucketed by length in shared memory) -- also what corpora without the
 *            interleaved copy and the streaming entry points use;
 *        2 = interleaved layout, rows streamed into per-warp shared-memory rings by TMA bulk copies (Jaro /
 *            Jaro-Winkler, which need random access to the candidate, stay on 0);
 *   "jaro32" (default 1): Jaro / Jaro-Winkler queries of at most 64 elements use the row-wise kernels (32-bit flags up to
 *        32 elements, 64-bit flags on 32-bit halves up to 64; interleaved layout only) with the f64 score algebra looked up
 *        in a per-launch table; 2 = the same kernels with the score algebra computed per pair, 3 = table, 48-register build (5 CTAs per SM)
 *        (both kept for A/B runs and cross-checks); 0 = the generic per-lane routine;
 *   "epilogue_table" (default 1): integer metrics with queries of at most 64 elements on a resident corpus whose candidates
 *        are at most 255 elements long: the score algebra (distance <-> similarity, normalisation, cutoff conversions, the
 *        final score() filter) is evaluated once per launch for every (candidate length, raw result) and looked up per
 *        pair; 0 = evaluated per pair (same results);
 *   "multi_word_path" (default 0): queries of 65..512 elements on a resident corpus: 0 = one thread per candidate, the
 *        whole bit-vector column in registers (interleaved layout), 1 = the sub-warp shuffle kernel (what longer
 *        queries, corpora without the interleaved copy and the streaming entry points use);
 *   "banded_levenshtein" (default 1): multi-word Levenshtein distance with score_cutoff <= 63 edits uses the
 *        one-thread-per-candidate 64-bit Ukkonen-band kernel; 0 = always the multi-word block kernel;
 *   "stream_chunk_mb" (default 64), "stream_chunk_kcand" (default 2048): chunk size of rf_batch_stream_* in
 *        MiB of candidate bytes / thousands (x1024) of candidates, whichever is hit first;
 *   "compact_u32_corpus" (default 1): rf_corpus_create_u32 renames corpora of at most 255 distinct symbols to
 *        bytes once at creation (0 = keep u32 elements and rename per scoring call);
 *   "cdist_slices" (default 0 = automatic, 1..256): corpus slices of rf_cdist_topk_* (work units = slices x queries);
 *   "sharded_collective" (default 0): the gathers of the one-process sharded entry points.  0 = automatic: the top-k lists
 *        travel by NCCL; the score vectors of *_allgather_device by the copy engines over NVLink when every device pair
 *        has peer access (DMA needs no SM, so piece k travels while piece k+1 is scanned), else by NCCL; a device
 *        list with duplicates always uses copies.  1 = copies everywhere, 2 = NCCL everywhere;
 *   "allgather_chunks" (default 0 = automatic, 1..16): pieces a shard is scanned in by the *_allgather_device entry points,
 *        piece k gathered while piece k+1 is scanned (1 = scan everything, then gather).  Automatic = 4 on the
 *        copy-engine path, 1 with NCCL: measured on 2 B200s, NCCL's broadcast kernels find no free CTA slot beside the
 *        persistent scan kernel, so pieces buy nothing there;
 *   "cdist_skip" (default 1): rf_cdist_topk_* skips groups whose length alone puts them beyond the running k-th
 *        distance (0 = score every candidate; for measurements). */
rf_status rf_set_option(const char* name, int value);

/* kernel launches issued by this library in this process so far (bench.py reports the delta) */
uint64_t rf_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RFGPU_H */
