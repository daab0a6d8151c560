#!/usr/bin/env python3
"""Dev helper: one small pass through every kernel family (run under compute-sanitizer on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rapidfuzz-rs_b200"))
import numpy as np
import rapidfuzz_b200 as rf
from rapidfuzz_b200 import _ffi
from rapidfuzz_b200._scorer import BatchComparatorBase

def bc(metric, q):
    return type("B", (BatchComparatorBase,), {"METRIC": metric})(q)

L = _ffi.lib()
q = rf.synth_query(1, 32)
chars, offsets = rf.synth_corpus(1, q, 3000, 0, 64, 16)
corpus = rf.Corpus(chars, offsets)
for path in (0, 1, 2):
    _ffi.check(L.rf_set_option(b"single_word_path", path))
    for m in ("levenshtein", "indel", "osa", "lcs_seq", "jaro_winkler"):
        for qq in (q, rf.synth_query(2, 50)):
            b = bc(m, qq)
            b._score("distance", corpus, None)
            b._score("normalized_similarity", corpus, rf.Args().score_cutoff(0.5))
            b.close()
_ffi.check(L.rf_set_option(b"single_word_path", 0))
b = bc("levenshtein", q)
b.extract("distance", corpus, k=10)
b.filter("distance", corpus, rf.Args().score_cutoff(10))
b.stream("distance", chars, offsets.astype(np.uint32))
b.close()
rf.cdist_topk([q, rf.synth_query(3, 20), rf.synth_query(4, 64)], corpus, k=10)
rf.cdist_topk([q], corpus, k=40, score_cutoff=30)
q3 = rf.synth_query(3, 256)
c3, o3 = rf.synth_corpus(3, q3, 2000, 64, 256, 48)
corpus3 = rf.Corpus(c3, o3)
for m in ("levenshtein", "indel", "osa", "jaro"):
    b = bc(m, q3)
    b._score("distance", corpus3, None)
    b.close()
b = bc("levenshtein", q3)
b._score("distance", corpus3, rf.Args().score_cutoff(32))
b._score("distance", corpus3, rf.Args().score_cutoff(63))
b._score("distance", corpus3, rf.Args().score_cutoff(100))
b.stream("distance", c3, o3, rf.Args().score_cutoff(32))
b.close()
corpus.close(); corpus3.close()
print("sanitize smoke done")
