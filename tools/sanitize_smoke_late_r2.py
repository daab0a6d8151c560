#!/usr/bin/env python3
"""Dev helper: the kernels added at the end of round 2, one small pass each (run under compute-sanitizer on the GPU box;
also exec'd by tools/sanitize_smoke.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) if "__file__" in globals() else ROOT
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rapidfuzz-rs_b200"))
import numpy as np
import rapidfuzz_b200 as rf
import synth
from rapidfuzz_b200 import _ffi
from rapidfuzz_b200._scorer import BatchComparatorBase

def bc(metric, q):
    return type("B", (BatchComparatorBase,), {"METRIC": metric})(q)

L = _ffi.lib()
q = synth.synth_query(1, 32)
chars, offsets = synth.synth_corpus(1, q, 3000, 0, 64, 16)
lens8 = np.diff(offsets.astype(np.int64)).astype(np.uint8)
# ---- end of round 2: row-wise Jaro for 33..64 (64-bit flags), the epilogue tables (Jaro: per launch; integer metrics: corpora
#      of >= 65 536 candidates none longer than 255), the per-pair epilogue builds, Hamming / Prefix / Postfix with one-wave grids
rng = np.random.default_rng(11)
lens = rng.choice([0, 1, 2, 8, 31, 32, 33, 63, 64, 65, 66, 67, 100, 129, 130, 255], 66_000)
cb = rng.integers(97, 101, int(lens.sum())).astype(np.uint8)
ob = np.zeros(len(lens) + 1, np.uint64); ob[1:] = np.cumsum(lens)
corpus_b = rf.Corpus(cb, ob)
for qlen in (1, 32, 33, 64):
    qq = rng.integers(97, 101, qlen).astype(np.uint8)
    for m in ("jaro", "jaro_winkler"):
        for opt in (1, 2, 3, 0):
            b = bc(m, qq)
            _ffi.check(L.rf_batch_set_option(b._h, b"jaro32", opt))
            b._score("similarity", corpus_b, None)
            b._score("normalized_distance", corpus_b, rf.Args().score_cutoff(0.3))
            b.close()
    for m in ("levenshtein", "osa", "indel", "lcs_seq", "ratio"):
        for opt in (1, 0):
            b = bc(m, qq)
            _ffi.check(L.rf_batch_set_option(b._h, b"epilogue_table", opt))
            if m != "ratio":
                b._score("similarity", corpus_b, None)
                b._score("normalized_similarity", corpus_b, rf.Args().score_cutoff(0.4))
            else:
                b._score("similarity", corpus_b, None)
            b.close()
    for m, a in (("hamming", rf.Args().pad(True)), ("prefix", None), ("postfix", None)):
        b = bc(m, qq)
        b._score("similarity", corpus_b, a)
        b.close()
packed6, dict64 = rf.pack6(chars)
b = bc("levenshtein", q)
b.stream_len8_packed6("distance", packed6, dict64, lens8, u8_results=True)
b.close()
corpus_b.close()
print("late round-2 sanitize smoke done")
