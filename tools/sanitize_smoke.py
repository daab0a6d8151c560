#!/usr/bin/env python3
"""Dev helper: one small pass through every kernel family (run under compute-sanitizer on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
corpus = rf.Corpus(chars, offsets)
for path in (0, 1, 2):
    _ffi.check(L.rf_set_option(b"single_word_path", path))
    for m in ("levenshtein", "indel", "osa", "lcs_seq", "jaro_winkler"):
        for qq in (q, synth.synth_query(2, 50)):
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
rf.cdist_topk([q, synth.synth_query(3, 20), synth.synth_query(4, 64)], corpus, k=10)
rf.cdist_topk([q], corpus, k=40, score_cutoff=30)
for sl in (3, 17, 0):   # several corpus slices -> cdist_merge_kernel; 0 = automatic
    _ffi.check(L.rf_set_option(b"cdist_slices", sl))
    rf.cdist_topk([synth.synth_query(10 + i, (8, 32, 47, 64)[i % 4]) for i in range(40)], corpus, k=10)
try:   # sharded merge (rf_topk_merge_device), three shards on one GPU
    import torch
    from rapidfuzz_b200 import sharding
    qs = [synth.synth_query(20 + i, 32) for i in range(8)]
    qo = np.arange(9, dtype=np.uint64) * 32
    parts, starts = [], []
    for r in range(3):
        c_, o_, lo = sharding.local_shard(chars, offsets, 3, r)
        cp = rf.Corpus(c_, o_)
        i_, d_ = sharding.cdist_topk_device(np.concatenate(qs), qo, cp, k=10)
        parts.append(torch.stack([i_, d_], 0)); starts.append(lo)
        torch.cuda.synchronize(); cp.close()
    sharding.merge_topk_device(torch.stack(parts, 0).contiguous(), torch.tensor(starts, dtype=torch.int64, device="cuda"), 10)
    torch.cuda.synchronize()
except ImportError:
    pass
q3 = synth.synth_query(3, 256)
c3, o3 = synth.synth_corpus(3, q3, 2000, 64, 256, 48)
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
# DP metrics / Hamming over the interleaved layout, and the hand-over kernels for long candidates
lens = np.array([0, 1, 9, 40, 64, 70, 200, 33] * 20)
lens[7] = 33000   # ONE candidate past the 16-bit DP cells (a 33000 x 64 DP under the sanitizer is slow)
cl = np.random.default_rng(3).integers(97, 101, int(lens.sum())).astype(np.uint8)
ol = np.zeros(len(lens) + 1, np.uint64); ol[1:] = np.cumsum(lens)
corpus_l = rf.Corpus(cl, ol)
for qq in (q, synth.synth_query(2, 64), synth.synth_query(2, 70)):
    for m, a in (("damerau_levenshtein", None), ("levenshtein", rf.Args().weights(1, 2, 3)), ("hamming", rf.Args().pad(True)),
                 ("jaro_winkler", None), ("jaro", rf.Args().score_cutoff(0.7))):
        b = bc(m, qq)
        b._score("distance", corpus_l, a)
        b._score("distance", corpus, a)
        b.close()
corpus_l.close()
# ---- round 2: register-column kernels (4 / 8 / 12 / 16 limbs), stripe kernel, Jaro long, len8 / elems32 streaming, sharded
for qlen in (100, 256, 300, 500):
    qq = synth.synth_query(7, qlen)
    for m in ("levenshtein", "indel", "osa"):
        b = bc(m, qq)
        b._score("distance", corpus3, None)
        b.close()
qlong = np.random.default_rng(5).integers(97, 101, 17000).astype(np.uint8)
lens = np.array([0, 5, 300, 16999, 17001, 40000])
cl = np.random.default_rng(6).integers(97, 101, int(lens.sum())).astype(np.uint8)
ol = np.zeros(len(lens) + 1, np.uint64); ol[1:] = np.cumsum(lens)
corpus_x = rf.Corpus(cl, ol)
for m in ("levenshtein", "indel", "osa"):
    b = bc(m, qlong)
    b._score("distance", corpus_x, None)
    b.close()
for m in ("jaro", "jaro_winkler"):
    b = bc(m, qlong[:3000])
    b._score("similarity", corpus_x, None)
    b.close()
corpus_x.close()
b = bc("levenshtein", q)
lens8 = np.diff(offsets.astype(np.int64)).astype(np.uint8)
b.stream_len8("distance", chars, lens8)
b.stream_len8("distance", chars, lens8, u8_results=True)
b.close()
b = bc("levenshtein", q.astype(np.uint32) + 1000)
b.stream_elems32("distance", chars.astype(np.uint32) + 1000, offsets)
b.close()
from rapidfuzz_b200 import sharding
devs = [0, 0]
sc = sharding.ShardedCorpus(chars, offsets, devs)
sb = sharding.ShardedBatchComparator("levenshtein", q, devs)
sb.score("distance", sc)
sb.extract("distance", sc, k=5)
import torch
bufs = [torch.empty(len(sc), dtype=torch.int32, device="cuda") for _ in devs]
sb.score_allgather("distance", sc, [x.data_ptr() for x in bufs])
sharding.sharded_cdist_topk(q, np.array([0, 32], np.uint64), sc, k=5)
sb.stream("distance", chars, offsets)
sb.close(); sc.close()
exec(open(os.path.join(ROOT, "tools", "sanitize_smoke_late_r2.py")).read())   # the kernels added at the end of round 2
corpus.close(); corpus3.close()
print("sanitize smoke done")
