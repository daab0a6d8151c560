"""BASELINE.json's FULL sizes on the GPU, checked through size-independent properties (the oracle only sees a
sample): two independent code paths must agree bit for bit on every candidate, bounds that hold for any pair,
planted near-matches are found, and a checksum of the result vector is reproducible."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

import rapidfuzz_b200 as rf
import synth
from rapidfuzz_b200 import _ffi
from oracle import oracle as orc
from gpu_util import gpu_batch


def _bc(metric, q):
    from rapidfuzz_b200._scorer import BatchComparatorBase
    return type("B", (BatchComparatorBase,), {"METRIC": metric})(q)


def test_config2_full_size_1e8():
    """config 2: 1 query len 32 vs 10^8 candidates len 8-64."""
    n = 100_000_000
    q = synth.synth_query(2, 32)
    chars, offsets = synth.synth_corpus(2, q, n, 8, 64, 16)
    lens = np.diff(offsets.astype(np.int64))
    corpus = rf.Corpus(chars, offsets)
    d = gpu_batch("levenshtein", "distance", q, corpus)                       # interleaved-layout kernel
    b = _bc("levenshtein", q)
    s = b.stream("distance", chars, offsets.astype(np.uint32))                # CSR / TMA-tile kernel, chunked
    assert np.array_equal(d, s)                                               # two kernels, two layouts, same bits
    assert np.all(d >= np.abs(lens - 32)) and np.all(d <= np.maximum(lens, 32))
    m = 1_000_000
    exp = orc.batch("levenshtein", "distance", q, chars[: int(offsets[m])], offsets[: m + 1], nthreads=0)
    assert np.array_equal(d[:m], exp)
    assert np.array_equal(d[-m:], orc.batch("levenshtein", "distance", q, chars[int(offsets[n - m]):],
                                            offsets[n - m:] - offsets[n - m], nthreads=0))
    assert (d <= 16).sum() >= n // 64 * 0.9                                   # the planted near-matches (1/64, <= 16 edits)
    # post-processing agrees with numpy on the full vector
    gi, gs = b.extract("distance", corpus, k=100)
    order = np.lexsort((np.arange(n), d))[:100]
    assert np.array_equal(gi, order.astype(np.uint32)) and np.array_equal(gs, d[order])
    fi, fs, tot = b.filter("distance", corpus, rf.Args().score_cutoff(6), capacity=1 << 20)
    hits = np.nonzero(d <= 6)[0]
    assert tot == len(hits) and np.array_equal(fi, hits[: 1 << 20].astype(np.uint32))
    # normalized_similarity is a pure function of (d, lens): size-independent identity
    ns = gpu_batch("levenshtein", "normalized_similarity", q, corpus)
    assert np.array_equal(ns, 1.0 - d / np.maximum(lens, 32))
    b.close()
    corpus.close()


def test_config3_full_size_1e7_banded_equals_block():
    """config 3: query len 256 vs 10^7 candidates len 64-256, score_cutoff 32: the one-word sliding band and the
    multi-word block kernel are different algorithms and must return the same Option for every candidate."""
    n = 10_000_000
    q = synth.synth_query(3, 256)
    chars, offsets = synth.synth_corpus(3, q, n, 64, 256, 48)
    lens = np.diff(offsets.astype(np.int64))
    corpus = rf.Corpus(chars, offsets)
    band = gpu_batch("levenshtein", "distance", q, corpus, cutoff=32)
    _ffi.check(_ffi.lib().rf_set_option(b"banded_levenshtein", 0))
    try:
        block = gpu_batch("levenshtein", "distance", q, corpus, cutoff=32)
    finally:
        _ffi.check(_ffi.lib().rf_set_option(b"banded_levenshtein", 1))
    assert np.array_equal(band, block)
    some = band != 0xFFFFFFFF
    assert np.all(band[some] <= 32) and np.all(band[some] >= np.abs(lens[some] - 256))
    assert not np.any(some & (np.abs(lens - 256) > 32))                       # the length filter (levenshtein.rs:1045-1047)
    assert some.sum() > n // 200
    m = 300_000
    exp = orc.batch("levenshtein", "distance", q, chars[: int(offsets[m])], offsets[: m + 1], nthreads=0, cutoff=32)
    assert np.array_equal(band[:m], exp)
    corpus.close()


def test_config4_full_size_1e8_jaro_winkler():
    """config 4: Jaro-Winkler normalized_similarity, 10^8 candidates: 32-bit row kernel vs the generic per-lane
    routine on every candidate (bit-identical f64), oracle within 1e-6 (and in fact exactly) on a sample."""
    n = 100_000_000
    q = synth.synth_query(4, 32)
    chars, offsets = synth.synth_corpus(4, q, n, 8, 64, 16)
    corpus = rf.Corpus(chars, offsets)
    fast = gpu_batch("jaro_winkler", "normalized_similarity", q, corpus)
    _ffi.check(_ffi.lib().rf_set_option(b"jaro32", 0))
    try:
        generic = gpu_batch("jaro_winkler", "normalized_similarity", q, corpus)
    finally:
        _ffi.check(_ffi.lib().rf_set_option(b"jaro32", 1))
    assert np.array_equal(fast, generic)
    assert np.all((fast >= 0.0) & (fast <= 1.0))
    m = 500_000
    exp = orc.batch("jaro_winkler", "normalized_similarity", q, chars[: int(offsets[m])], offsets[: m + 1], nthreads=0)
    assert np.max(np.abs(fast[:m] - exp)) <= 1e-6
    assert np.array_equal(fast[:m], exp)
    corpus.close()


def test_corpus_beyond_4gib_uses_64bit_offsets():
    """1.25e8 candidates = 4.5 GB of characters: the device keeps u64 CSR offsets (total >= 2^32 - 16).  Resident
    path (interleaved layout built from u64 offsets), CSR path and the streaming pipeline (chunk bases beyond
    2^32) must agree with each other everywhere and with the oracle at both ends of the corpus."""
    n = 125_000_000
    q = synth.synth_query(6, 32)
    chars, offsets = synth.synth_corpus(6, q, n, 8, 64, 16)
    assert int(offsets[n]) >= 2**32
    corpus = rf.Corpus(chars, offsets)
    d = gpu_batch("levenshtein", "distance", q, corpus)
    _ffi.check(_ffi.lib().rf_set_option(b"single_word_path", 1))     # CSR / TMA-tile kernel on the resident corpus
    try:
        d_csr = gpu_batch("levenshtein", "distance", q, corpus)
    finally:
        _ffi.check(_ffi.lib().rf_set_option(b"single_word_path", 0))
    assert np.array_equal(d, d_csr)
    b = _bc("levenshtein", q)
    s = b.stream("distance", chars, offsets)                         # u64 host offsets, chunked
    b.close()
    assert np.array_equal(d, s)
    m = 300_000
    assert np.array_equal(d[:m], orc.batch("levenshtein", "distance", q, chars[: int(offsets[m])], offsets[: m + 1], nthreads=0))
    tail = orc.batch("levenshtein", "distance", q, chars[int(offsets[n - m]):], offsets[n - m:] - offsets[n - m], nthreads=0)
    assert np.array_equal(d[-m:], tail)
    # multi-word banded path reads the same u64 offsets
    q3 = synth.synth_query(7, 100)
    bd = gpu_batch("levenshtein", "distance", q3, corpus, cutoff=40)
    exp = orc.batch("levenshtein", "distance", q3, chars[int(offsets[n - m]):], offsets[n - m:] - offsets[n - m], nthreads=0, cutoff=40)
    assert np.array_equal(bd[-m:], exp)
    corpus.close()


def test_config5_full_size_corpus_cdist_topk():
    """config 5's corpus at full size (10^7 candidates len 8-64, ~440 MB of layout: several L2-sized slices) against 2 000
    queries: the per-query top-10 must be sorted by (distance, index) with distinct indices, every listed distance must
    be the true distance of that pair (oracle on the listed candidates), no distance may undercut the length bound, the
    planted near-matches of query 0 lead its list, one slice vs automatic slices must agree on every entry, and five
    queries are checked against the oracle's top-10 over the WHOLE corpus."""
    n, nq, k = 10_000_000, 2000, 10
    qs = [synth.synth_query(5 + i, 32) for i in range(nq)]
    chars, offsets = synth.synth_corpus(5, qs[0], n, 8, 64, 16)
    lens = np.diff(offsets.astype(np.int64))
    corpus = rf.Corpus(chars, offsets)
    q_chars = np.concatenate(qs)
    q_off = np.arange(nq + 1, dtype=np.uint64) * 32
    idx, dist = rf.cdist_topk((q_chars, q_off), corpus, k=k)
    L = _ffi.lib()
    _ffi.check(L.rf_set_option(b"cdist_slices", 1))
    try:
        idx1, dist1 = rf.cdist_topk((q_chars[: 200 * 32], q_off[:201]), corpus, k=k)
    finally:
        _ffi.check(L.rf_set_option(b"cdist_slices", 0))
    assert np.array_equal(idx1, idx[:200]) and np.array_equal(dist1, dist[:200])
    key = dist.astype(np.int64) * (1 << 32) + idx.astype(np.int64)
    assert np.all(key[:, 1:] > key[:, :-1])                                   # sorted by (distance, index), indices distinct
    assert idx.max() < n
    assert np.all(dist >= np.abs(lens[idx.astype(np.int64)] - 32))            # d >= |len2 - len1|
    assert dist[0, 0] == 0 or dist[0, 0] <= 16                                # query 0's planted near-matches
    for qi in list(range(0, nq, 97)):                                         # listed distances are the true distances
        sel = idx[qi].astype(np.int64)
        sub_chars = np.concatenate([chars[int(offsets[j]):int(offsets[j + 1])] for j in sel])
        sub_off = np.zeros(k + 1, np.uint64)
        sub_off[1:] = np.cumsum(lens[sel])
        assert np.array_equal(orc.batch("levenshtein", "distance", qs[qi], sub_chars, sub_off, nthreads=0), dist[qi])
    for qi in (0, 1, 777, 1500, nq - 1):                                      # complete: the oracle's top-10 over the whole corpus
        d = orc.batch("levenshtein", "distance", qs[qi], chars, offsets, nthreads=0).astype(np.int64)
        best = np.sort(d * (1 << 32) + np.arange(n))[:k]
        assert np.array_equal(idx[qi], (best & 0xFFFFFFFF).astype(np.uint32)) and np.array_equal(dist[qi], (best >> 32).astype(np.uint32))
    corpus.close()
