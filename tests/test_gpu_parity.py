"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> librfgpu.so), against the CPU
oracle on the same seeded inputs.  Integer results must be bit-exact; f64 results are required bit-exact too
(the kernels reproduce the reference's operation order with FMA contraction off), which is stricter than the
1e-6 the north star asks for."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import rapidfuzz_b200 as rf
import synth
from rapidfuzz_b200 import _ffi
from oracle import oracle as orc
from gpu_util import make_corpus, check, gpu_batch, assert_same



PATHS = {"interleaved_ldg": 0, "csr_tiles": 1, "interleaved_tma_ring": 2}


@pytest.fixture(autouse=True, params=list(PATHS))
def single_word_path(request):
    """Every test runs against all three single-word kernels: the length-bucketed interleaved layout read with
    per-lane streaming loads (scan_lb_kernel, default) or through per-warp TMA rings (scan_lbr_kernel), and the
    CSR / TMA-tile path (scan_w1_kernel)."""
    _ffi.check(_ffi.lib().rf_set_option(b"single_word_path", PATHS[request.param]))
    yield request.param
    _ffi.check(_ffi.lib().rf_set_option(b"single_word_path", 0))


INT_METRICS = ["levenshtein", "indel", "lcs_seq", "osa"]
ALL_KINDS = ["distance", "similarity", "normalized_distance", "normalized_similarity"]
EDGE_LENS = [0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 66, 100, 127, 128, 129, 200, 255, 256, 257]


def test_config1_levenshtein_1000():
    """BASELINE config 1: 1 ASCII query len 32 vs 1000 candidates len 8-64, bit-exact."""
    q = synth.synth_query(1, 32)
    chars, offsets = synth.synth_corpus(1, q, 1000, 8, 64, 16)
    check("levenshtein", "distance", q, chars, offsets)


@pytest.mark.parametrize("qlen", [0, 1, 2, 5, 8, 16, 31, 32, 33, 48, 63, 64])
def test_single_word_all_metrics_all_kinds(qlen):
    rng = np.random.default_rng(100 + qlen)
    q = (rng.integers(0, 5, qlen) + 97).astype(np.uint8)
    chars, offsets = make_corpus(rng, 3000, EDGE_LENS, alphabet=5, query=q, high_bytes=True)
    corpus = rf.Corpus(chars, offsets)
    for m in INT_METRICS:
        for kind in ALL_KINDS:
            check(m, kind, q, chars, offsets, corpus)
    for m in ("jaro", "jaro_winkler"):
        for kind in ALL_KINDS:
            check(m, kind, q, chars, offsets, corpus)
    check("ratio", "similarity", q, chars, offsets, corpus)
    check("ratio", "similarity", q, chars, offsets, corpus, reference_quirks=True)
    corpus.close()


@pytest.mark.parametrize("qlen", [20, 32, 40, 64])
def test_single_word_cutoffs(qlen):
    rng = np.random.default_rng(200 + qlen)
    q = (rng.integers(0, 4, qlen) + 97).astype(np.uint8)
    chars, offsets = make_corpus(rng, 2000, [0, 1, 8, 20, 32, 40, 64, 70], alphabet=4, query=q, near_frac=0.6)
    corpus = rf.Corpus(chars, offsets)
    for m in INT_METRICS:
        for c in (0, 1, 2, 3, 4, 5, 10, 31, 32, 64, 2**64 - 1):
            check(m, "distance", q, chars, offsets, corpus, cutoff=c)
            if m != "levenshtein":  # reference similarity+cutoff underflows for Levenshtein
                check(m, "similarity", q, chars, offsets, corpus, cutoff=c)
        for c in (0.0, 0.1, 0.3, 0.5, 0.75, 0.9, 1.0, 1.5, -0.5):
            check(m, "normalized_distance", q, chars, offsets, corpus, cutoff=c)
            check(m, "normalized_similarity", q, chars, offsets, corpus, cutoff=c)
    for m in ("jaro", "jaro_winkler"):
        for kind in ALL_KINDS:
            for c in (0.0, 0.3, 0.6, 0.7, 0.75, 0.85, 0.95, 1.0, 1.1):
                check(m, kind, q, chars, offsets, corpus, cutoff=c)
    check("jaro_winkler", "similarity", q, chars, offsets, corpus, prefix_weight=0.25)
    for c in (0.0, 0.5, 0.9):
        check("ratio", "similarity", q, chars, offsets, corpus, cutoff=c)
    for w in ((2, 2, 2), (1, 1, 2), (3, 3, 7), (0, 0, 5), (1, 2, 3), (2, 1, 1), (3, 5, 4), (0, 1, 1), (7, 7, 9)):
        for c in (None, 3, 10, 100):   # uniform, indel-class, zero and generic (Wagner-Fischer) weight classes
            check("levenshtein", "distance", q, chars, offsets, corpus, weights=w, cutoff=c)
        check("levenshtein", "similarity", q, chars, offsets, corpus, weights=w)
        check("levenshtein", "normalized_similarity", q, chars, offsets, corpus, weights=w)
        check("levenshtein", "normalized_distance", q, chars, offsets, corpus, weights=w, cutoff=0.4)
    corpus.close()


def test_lev_similarity_cutoff_is_none_not_wrapped():
    q = b"kitten"
    c = rf.Corpus.from_strings([b"sitting", b"kitten", b"zzzzzzzzzz"])
    r = rf.distance.levenshtein.BatchComparator(q).similarity_with_args(c, rf.Args().score_cutoff(4))
    assert r.tolist() == [4, 6, None]


@pytest.mark.parametrize("qlen", [65, 66, 100, 128, 129, 200, 256, 300, 512, 513, 700, 2048, 2100, 5000])
def test_multi_word_integer_metrics(qlen):
    rng = np.random.default_rng(300 + qlen)
    q = (rng.integers(0, 4, qlen) + 97).astype(np.uint8)
    lens = [0, 1, 5, 63, 64, 65, 127, 128, 129, 200, 256, 257, qlen - 1, qlen, qlen + 1, qlen + 40]
    n = 600 if qlen <= 700 else 150
    chars, offsets = make_corpus(rng, n, lens, alphabet=4, query=q, near_frac=0.5)
    corpus = rf.Corpus(chars, offsets)
    for m in INT_METRICS:
        check(m, "distance", q, chars, offsets, corpus)
        check(m, "normalized_similarity", q, chars, offsets, corpus)
    for c in (0, 3, 4, 31, 32, 33, 63, 64, 100):
        check("levenshtein", "distance", q, chars, offsets, corpus, cutoff=c)   # <= 63: banded kernel, else block kernel
        check("indel", "distance", q, chars, offsets, corpus, cutoff=c)
    _ffi.check(_ffi.lib().rf_set_option(b"banded_levenshtein", 0))             # same cutoffs through the block kernel
    try:
        for c in (0, 4, 32, 63):
            check("levenshtein", "distance", q, chars, offsets, corpus, cutoff=c)
    finally:
        _ffi.check(_ffi.lib().rf_set_option(b"banded_levenshtein", 1))
    for w, c in (((2, 2, 2), 64), ((3, 3, 3), 100), ((2, 2, 2), 127), ((5, 5, 5), 3)):
        check("levenshtein", "distance", q, chars, offsets, corpus, weights=w, cutoff=c)
    check("lcs_seq", "similarity", q, chars, offsets, corpus, cutoff=qlen // 2)
    check("levenshtein", "distance", q, chars, offsets, corpus, weights=(1, 1, 2))
    if qlen <= 2048:
        check("levenshtein", "distance", q, chars, offsets, corpus, weights=(1, 2, 3), cutoff=qlen)
    check("ratio", "similarity", q, chars, offsets, corpus, cutoff=0.5)
    corpus.close()


@pytest.mark.parametrize("qlen", [65, 100, 128, 200, 256, 257, 700, 2048])
def test_multi_word_jaro(qlen):
    rng = np.random.default_rng(400 + qlen)
    q = (rng.integers(0, 5, qlen) + 97).astype(np.uint8)
    lens = [0, 1, 5, 64, 65, 128, 200, 300, qlen - 1, qlen, qlen + 1, 2 * qlen + 7]
    chars, offsets = make_corpus(rng, 300, lens, alphabet=5, query=q, near_frac=0.5)
    corpus = rf.Corpus(chars, offsets)
    for m in ("jaro", "jaro_winkler"):
        for kind in ALL_KINDS:
            check(m, kind, q, chars, offsets, corpus)
        check(m, "similarity", q, chars, offsets, corpus, cutoff=0.8)
    corpus.close()


def test_config3_shape_banded_vs_oracle():
    """BASELINE config 3 shape (query len 256, candidates 64-256, cutoff 32) at 3e5 candidates, plus other
    cutoffs / query lengths: the three-pass banded path (classify, first columns, near-matches) vs the oracle."""
    for seed, qlen, lo, hi, kmax, n, cut in ((3, 256, 64, 256, 48, 300_000, 32), (13, 1000, 900, 1100, 80, 20_000, 63),
                                            (23, 65, 1, 130, 10, 50_000, 7), (33, 300, 280, 320, 5, 40_000, 0),
                                            (43, 3000, 2990, 3010, 20, 3_000, 20)):
        q = synth.synth_query(seed, qlen)
        chars, offsets = synth.synth_corpus(seed, q, n, lo, hi, kmax)
        corpus = rf.Corpus(chars, offsets)
        got = gpu_batch("levenshtein", "distance", q, corpus, cutoff=cut)
        exp = orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=0, cutoff=cut)
        assert_same(got, exp, ("banded", seed, qlen, cut))
        assert (exp != 0xFFFFFFFF).sum() > 0
        corpus.close()


def test_query_64_vs_long_candidates_and_tile_overflow():
    """Candidates far longer than the staging tile (forces the direct-from-global path) and Jaro's
    block path for query <= 64 with long candidates (jaro.rs:584-595)."""
    rng = np.random.default_rng(5)
    q = (rng.integers(0, 4, 50) + 97).astype(np.uint8)
    chars, offsets = make_corpus(rng, 700, [10, 50, 400, 3000, 40000], alphabet=4, query=q)
    corpus = rf.Corpus(chars, offsets)
    for m in INT_METRICS:
        check(m, "distance", q, chars, offsets, corpus)
    check("jaro_winkler", "normalized_similarity", q, chars, offsets, corpus)
    check("jaro", "similarity", q, chars, offsets, corpus, cutoff=0.4)
    # query <= 32 (row-wise 32-bit Jaro kernel): groups of long candidates take its generic fallback
    for ql in (20, 32):
        q2 = q[:ql]
        for m in ("jaro", "jaro_winkler"):
            check(m, "similarity", q2, chars, offsets, corpus)
            check(m, "normalized_distance", q2, chars, offsets, corpus, cutoff=0.5)
    corpus.close()


def test_empty_and_ragged_inputs():
    q = b"abc"
    # empty corpus
    c0 = rf.Corpus(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(rf.distance.levenshtein.BatchComparator(q).distance(c0)) == 0
    c0.close()
    # all-empty candidates, and one candidate
    chars, offsets = np.zeros(0, np.uint8), np.zeros(8, np.uint64)
    for m in INT_METRICS + ["jaro", "jaro_winkler"]:
        for kind in ALL_KINDS:
            check(m, kind, q, chars, offsets)
            check(m, kind, b"", chars, offsets)
    assert rf.distance.levenshtein.distance(b"CA", b"ABC") == 3               # levenshtein.rs:1378
    assert rf.distance.levenshtein.BatchComparator(b"CA").distance(b"ABC") == 3  # :1632-1633
    assert rf.distance.levenshtein.distance(b"kitten", b"sitting", rf.Args().score_cutoff(2)) is None
    assert abs(rf.fuzz.ratio("this is a test", "this is a test!") - 0.9655172413793104) < 1e-12


G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def test_reference_golden_vectors_on_gpu():
    """Every u8 known-answer vector of the reference (tests/golden/golden.json) through the GPU path."""
    n_run = 0
    for rec in G["cases"]:
        a = rec["args"]
        w = tuple(a["weights"]) if "weights" in a else None
        wide = "s1" not in rec or "s2" not in rec   # non-ASCII cases: u32 elements (code points)
        e1 = np.array(rec["s1_cp"], np.uint32) if "s1_cp" in rec else np.frombuffer(rec["s1"].encode(), np.uint8)
        e2 = np.array(rec["s2_cp"], np.uint32) if "s2_cp" in rec else np.frombuffer(rec["s2"].encode(), np.uint8)
        for q, c in ((e1, e2), (e2, e1)):
            if wide:
                corpus = rf.Corpus.from_u32(c.astype(np.uint32), np.array([0, len(c)], np.uint64))
                q = q.astype(np.uint32)
            else:
                corpus = rf.Corpus(c, np.array([0, len(c)], np.uint64))
            try:
                if rec["expected"] == "error":       # hamming::Error::DifferentLengthArgs
                    with pytest.raises(rf.RfError):
                        _gpu_pad(rec["metric"], rec["kind"], q, corpus, cutoff=a.get("cutoff"), pad=a.get("pad", False))
                    n_run += 1
                    continue
                if rec["metric"] in ("hamming", "prefix", "postfix"):
                    got = _gpu_pad(rec["metric"], rec["kind"], q, corpus, cutoff=a.get("cutoff"), pad=a.get("pad", False))[0]
                else:
                    got = gpu_batch(rec["metric"], rec["kind"], q, corpus, cutoff=a.get("cutoff"), weights=w)[0]
            finally:
                corpus.close()
            exp = rec["expected"]
            is_none = (np.isnan(got) if got.dtype == np.float64 else got == _ffi.NONE_U32)
            if exp is None:
                assert is_none, (rec, got)
            else:
                assert not is_none, (rec, got)
                assert abs(float(got) - exp) <= rec["tol"], (rec, got)
            n_run += 1
    assert n_run > 250
    for metric in ("jaro", "jaro_winkler"):
        m = G["matrices"][metric]
        names = m["names"]
        corpus = rf.Corpus.from_strings(names)
        for c in m["cutoffs"]:
            for i, n1 in enumerate(names):
                got = gpu_batch(metric, "similarity", n1.encode(), corpus, cutoff=c)
                gotd = gpu_batch(metric, "distance", n1.encode(), corpus, cutoff=1.0 - c)
                for j in range(len(names)):
                    sc = m["scores"][i * len(names) + j]
                    if c <= sc:
                        assert abs(got[j] - sc) <= m["tol"], (metric, n1, names[j], c, got[j])
                        assert abs(gotd[j] - (1.0 - sc)) <= m["tol"], (metric, "dist", n1, names[j], c, gotd[j])
                    else:
                        assert np.isnan(got[j]) or abs(got[j] - c) < 1e-9, (metric, n1, names[j], c, got[j])
        corpus.close()


@pytest.mark.parametrize("qlen", [0, 1, 7, 32, 33, 64, 65, 200, 256])
def test_u32_elements_vs_oracle(qlen):
    """u32 elements (Rust char / u32): per-query alphabet renaming on the device, exact against the oracle's
    hashmap-based u32 path; u32 query against a u8 corpus; unicode strings through the host mirror."""
    rng = np.random.default_rng(500 + qlen)
    alphabet = np.array([97, 98, 99, 255, 256, 1048, 0x4E2D, 0x1F600, 0xFFFFFFFF, 0], dtype=np.uint32)
    q = alphabet[rng.integers(0, 6, qlen)]
    lens = [0, 1, 5, 31, 32, 33, 64, 65, 100, max(qlen, 1), qlen + 3]
    cands = []
    for _ in range(1500):
        if qlen and rng.random() < 0.4:
            c = list(q)
            for _ in range(int(rng.integers(0, 8))):
                pos = int(rng.integers(0, len(c) + 1))
                op = rng.integers(0, 3)
                if op == 0 and c:
                    c[min(pos, len(c) - 1)] = alphabet[rng.integers(0, len(alphabet))]
                elif op == 1:
                    c.insert(pos, alphabet[rng.integers(0, len(alphabet))])
                elif c:
                    del c[min(pos, len(c) - 1)]
            cands.append(np.array(c, dtype=np.uint32))
        else:
            cands.append(alphabet[rng.integers(0, len(alphabet), int(rng.choice(lens)))])
    elems = np.concatenate(cands + [np.zeros(0, np.uint32)]).astype(np.uint32)
    offsets = np.zeros(len(cands) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(c) for c in cands])
    corpus = rf.Corpus.from_u32(elems, offsets)
    combos = [(m, k, None) for m in INT_METRICS + ["jaro", "jaro_winkler"] for k in ("distance", "normalized_similarity")]
    combos += [("levenshtein", "distance", 3), ("levenshtein", "distance", 40), ("indel", "similarity", 10),
               ("jaro_winkler", "similarity", 0.8), ("ratio", "similarity", 0.5)]
    for m, kind, cut in combos:
        if m in ("jaro", "jaro_winkler") and qlen > 2048:
            continue
        kw = {} if cut is None else {"cutoff": cut}
        got = gpu_batch(m, kind, q, corpus, **kw)
        exp = orc.batch(m, kind, q, elems, offsets, nthreads=0, **kw)
        assert_same(got, exp, ("u32", m, kind, cut, qlen))
    # top-k on a u32 corpus goes through the same renaming
    if qlen:
        b = _bc("levenshtein", q)
        gi, gs = b.extract("distance", corpus, k=5)
        exp = orc.batch("levenshtein", "distance", q, elems, offsets, nthreads=0)
        order = np.lexsort((np.arange(len(exp)), exp))[:5]
        assert np.array_equal(gi, order.astype(np.uint32)) and np.array_equal(gs, exp[order])
        b.close()
    corpus.close()
    # a u32 query against a byte corpus == the byte query when all its symbols are bytes
    qb = (q % 3 + 97).astype(np.uint32)
    chars, off8 = make_corpus(rng, 500, [0, 5, 40, 70], alphabet=3, query=qb.astype(np.uint8))
    c8 = rf.Corpus(chars, off8)
    assert_same(gpu_batch("levenshtein", "distance", qb, c8), gpu_batch("levenshtein", "distance", qb.astype(np.uint8), c8), "u32 q / u8 corpus")
    c8.close()


@pytest.mark.parametrize("qlen", [1, 2, 5, 31, 32, 33, 40, 63, 64])
def test_jaro_rowwise_kernel_degenerate_lanes(qlen):
    """The row-wise Jaro kernel splits a group's rows into warp-uniform segments from the lanes' window radii.  Padding
    lanes (length 0) and 1 x 1 pairs have a WRAPPED radius (jaro.rs: len/2 - 1): they must not drag the real lanes of
    their group into the wrong segment.  Small alphabet, lengths that put 1 / 65 / 100 long candidates and the padding
    of the last group into shared groups, candidates longer than the truncated length (jaro.rs:553-565)."""
    rng = np.random.default_rng(70 + qlen)
    for n, lo in ((3001, 1), (33, 1), (1, 1), (3001, 0)):         # lo = 0: zero bytes in query and candidates (the layout's padding byte)
        q = rng.integers(lo, 5, qlen).astype(np.uint8)
        lens = rng.choice([0, 1, 3, 8, 31, 32, 33, 50, 64, 65, 100], n)
        lens[-1] = 100                                            # the longest candidates share a group with the padding
        chars = rng.integers(lo, 5, int(lens.sum())).astype(np.uint8)
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        corpus = rf.Corpus(chars, off)
        for m, kind, cut in (("jaro", "similarity", None), ("jaro_winkler", "normalized_similarity", None),
                             ("jaro", "distance", 0.4), ("jaro_winkler", "similarity", 0.8)):
            kw = {} if cut is None else {"cutoff": cut}
            assert_same(gpu_batch(m, kind, q, corpus, **kw), orc.batch(m, kind, q, chars, off, nthreads=0, **kw), (m, kind, cut, qlen, n))
        corpus.close()


def test_dp_metrics_shared_memory_kernels_and_long_candidates():
    """Damerau-Levenshtein and generic-weight Levenshtein with queries of at most 64 elements run over the interleaved
    layout with their DP rows in shared memory (16-bit cells for Damerau-Levenshtein): candidates beyond 32 000 elements
    are handed to the global-scratch kernel through a flag; empty query / empty and very short candidates; query 64 / 65
    (the boundary between the two kernels)."""
    rng = np.random.default_rng(4242)
    for qlen in (0, 1, 17, 64, 65):
        q = rng.integers(97, 101, qlen).astype(np.uint8)
        lens = rng.choice([0, 1, 2, 9, 30, 64, 70], 700)
        lens[13] = 33000                                          # one candidate past the 16-bit cell range
        lens[500] = 32000
        chars = rng.integers(97, 101, int(lens.sum())).astype(np.uint8)
        off = np.zeros(len(lens) + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        corpus = rf.Corpus(chars, off)
        for m, kind, kw in (("damerau_levenshtein", "distance", {}), ("damerau_levenshtein", "normalized_similarity", {}),
                            ("damerau_levenshtein", "distance", {"cutoff": 20}), ("levenshtein", "distance", {"weights": (1, 2, 3)}),
                            ("levenshtein", "similarity", {"weights": (3, 1, 7)}), ("levenshtein", "distance", {"weights": (3, 1, 7), "cutoff": 40})):
            # (similarity + cutoff on Levenshtein is SURVEY Q2: the reference wraps, this library returns None -- not compared)
            assert_same(gpu_batch(m, kind, q, corpus, **kw), orc.batch(m, kind, q, chars, off, nthreads=0, **kw), (m, kind, kw, qlen))
        corpus.close()


@pytest.mark.parametrize("distinct", [10, 255, 256, 400])
def test_u32_corpus_alphabet_compaction(distinct):
    """rf_corpus_create_u32 keeps a corpus of at most 255 distinct symbols as bytes renamed ONCE (codes by ascending
    symbol); a u32 query is renamed through the corpus dictionary, symbols the corpus never contains match nothing.
    256+ distinct symbols fall back to the per-query renaming.  Both must equal the oracle, and each other."""
    rng = np.random.default_rng(900 + distinct)
    alphabet = np.unique(np.concatenate([np.array([0, 97, 255, 256, 0x4E2D, 0xFFFFFFFF], np.uint32),
                                         rng.integers(0, 2**32, 2 * distinct, dtype=np.uint64).astype(np.uint32)]))[:distinct]
    rng.shuffle(alphabet)
    assert len(alphabet) == distinct
    lens = rng.integers(0, 70, 4000)
    elems = alphabet[rng.integers(0, distinct, int(lens.sum()))]
    elems[:distinct] = alphabet                                   # every symbol occurs
    offsets = np.zeros(len(lens) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    L = _ffi.lib()
    absent = np.uint32(0x0BADF00D)
    assert absent not in alphabet
    queries = [alphabet[rng.integers(0, min(distinct, 40), 32)], alphabet[rng.integers(0, distinct, 50)],
               np.concatenate([alphabet[:5], [absent, absent, np.uint32(0x0BADF00E)], alphabet[2:9]]).astype(np.uint32)]
    # a candidate equal to each query (absent symbols cannot occur in the corpus, so only for the first two)
    results = {}
    for opt in (1, 0):
        _ffi.check(L.rf_set_option(b"compact_u32_corpus", opt))
        try:
            corpus = rf.Corpus.from_u32(elems, offsets)
            other = rf.Corpus.from_u32(elems[::-1].copy(), offsets)   # a second dictionary for the same comparator
        finally:
            _ffi.check(L.rf_set_option(b"compact_u32_corpus", 1))
        for qi, q in enumerate(queries):
            for m, kind, cut in (("levenshtein", "distance", None), ("indel", "normalized_similarity", None), ("osa", "distance", 30),
                                 ("jaro_winkler", "similarity", None), ("hamming", "distance", None), ("damerau_levenshtein", "distance", None)):
                b = _bc(m, q)
                a = rf.Args().pad(True) if m == "hamming" else rf.Args()
                if cut is not None:
                    a = a.score_cutoff(cut)
                fill = lambda r: r.filled(np.nan if r.dtype == np.float64 else _ffi.NONE_U32) if isinstance(r, np.ma.MaskedArray) else r
                got = fill(b._score(kind, corpus, a))
                got2 = fill(b._score(kind, other, a))      # same comparator, second dictionary
                got = fill(b._score(kind, corpus, a)) if qi == 0 else got
                b.close()
                kw = {} if cut is None else {"cutoff": cut}
                if m == "hamming":
                    kw["pad"] = True
                exp = orc.batch(m, kind, q, elems, offsets, nthreads=0, **kw)
                exp2 = orc.batch(m, kind, q, elems[::-1].copy(), offsets, nthreads=0, **kw)
                assert_same(got, exp, ("compact", opt, distinct, qi, m, kind))
                assert_same(got2, exp2, ("compact other", opt, distinct, qi, m, kind))
                results[(opt, qi, m)] = got
        corpus.close()
        other.close()
    for (opt, qi, m), v in results.items():
        if opt == 1:
            assert_same(v, results[(0, qi, m)], ("compact == per-query", qi, m))


def test_integer_elements_of_other_widths_compare_by_value():
    """HashableChar covers u8...u64 / i8...i64 (details/common.rs:29-37) and compares numerically: the Python mirror widens
    every integer array to the ABI's u8 / u32 BY VALUE, so i16 -1 never meets u16 65535 or u8 255, i8 / u16 / i32 sequences
    score like their value-renamed u32 images, and a byte query works against a wide corpus."""
    rng = np.random.default_rng(31)
    vals = np.array([-300, -1, 0, 5, 97, 255, 256, 40000], dtype=np.int64)
    lens = rng.integers(0, 40, 800)
    cand = vals[rng.integers(0, len(vals), int(lens.sum()))]
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    q = vals[rng.integers(0, len(vals), 20)]
    ids = {int(v): i + 1 for i, v in enumerate(vals)}                      # reference semantics == equality of VALUES
    ren = lambda a: np.array([ids[int(x)] for x in a], dtype=np.uint32)
    corpus = rf.Corpus.from_elems(cand.astype(np.int32), off)
    for m, kind in (("levenshtein", "distance"), ("jaro_winkler", "similarity"), ("indel", "normalized_similarity")):
        exp = orc.batch(m, kind, ren(q), ren(cand), off, nthreads=0)
        assert_same(gpu_batch(m, kind, q.astype(np.int32), corpus), exp, ("i32", m))
        assert_same(gpu_batch(m, kind, q.astype(np.int64), corpus), exp, ("i64 values that fit", m))
    corpus.close()
    # u16 vs i16: 65535 and -1 share 16 bits but not the value
    cu = rf.Corpus.from_elems(np.array([65535, 7, 65535], dtype=np.uint16), np.array([0, 3], np.uint64))
    assert gpu_batch("levenshtein", "distance", np.array([-1, 7, -1], dtype=np.int16), cu)[0] == 2
    assert gpu_batch("levenshtein", "distance", np.array([65535, 7, 65535], dtype=np.uint16), cu)[0] == 0
    assert gpu_batch("levenshtein", "distance", np.array([7], dtype=np.uint8), cu)[0] == 2          # byte query, wide corpus
    cu.close()
    c8 = rf.Corpus.from_elems(np.array([255, 7], dtype=np.uint8), np.array([0, 2], np.uint64))
    assert gpu_batch("levenshtein", "distance", np.array([-1, 7], dtype=np.int8), c8)[0] == 1      # i8 -1 is not u8 255
    c8.close()
    with pytest.raises(NotImplementedError):
        rf.Corpus.from_elems(np.array([1 << 40], dtype=np.uint64), np.array([0, 1], np.uint64))


def test_u32_host_mirror_and_limits():
    assert rf.distance.levenshtein.distance("Иванко", "Петрунко") == 5            # levenshtein.rs:2164-2169
    assert rf.distance.indel.distance("Иванко", "Петрунко") == 8                  # indel.rs:851-857
    assert rf.distance.levenshtein.BatchComparator("kitten").distance("sittinĝ") == 3   # byte query, wide candidate
    c = rf.Corpus.from_unicode(["Петрунко", "Иванко", "", "abc"])
    assert rf.distance.levenshtein.BatchComparator("Иванко").distance(c).tolist() == [5, 0, 6, 6]
    # more distinct symbols than a byte alphabet holds: against byte / byte-renamed corpora the query is mapped into THEIR
    # symbol domain, against other u32 corpora the candidates are renamed to 16-bit codes
    big = rf.distance.levenshtein.BatchComparator(np.arange(1000, 1300, dtype=np.uint32))
    strs = ["Петрунко", "Иванко", "", "abc"]     # 1000..1299 covers most of the Cyrillic block: some letters do match
    cps = np.array([ord(ch) for s_ in strs for ch in s_], dtype=np.uint32)
    offs = np.cumsum([0] + [len(s_) for s_ in strs]).astype(np.uint64)
    assert np.array_equal(big.distance(c), orc.batch("levenshtein", "distance", np.arange(1000, 1300, dtype=np.uint32), cps, offs, nthreads=0))
    _ffi.check(_ffi.lib().rf_set_option(b"compact_u32_corpus", 0))
    try:
        c_raw = rf.Corpus.from_unicode(strs)
        assert np.array_equal(big.distance(c_raw), orc.batch("levenshtein", "distance", np.arange(1000, 1300, dtype=np.uint32), cps, offs, nthreads=0))
        c_raw.close()
    finally:
        _ffi.check(_ffi.lib().rf_set_option(b"compact_u32_corpus", 1))
    big.close()
    with pytest.raises(rf.RfError):          # u32 corpus with a byte comparator handle
        b = rf.distance.levenshtein.BatchComparator(b"abc")
        out = np.zeros(4, np.uint32)
        _ffi.check(_ffi.lib().rf_batch_score_u32(b._h, c._h, 0, None, out.ctypes.data))
    c.close()


@pytest.mark.parametrize("qlen", [0, 1, 3, 4, 17, 32, 64, 65, 300])
def test_hamming_prefix_postfix_vs_oracle(qlen):
    """distance::{hamming, prefix, postfix}: every kind and cutoff ladder against the oracle; Hamming's pad / error."""
    rng = np.random.default_rng(700 + qlen)
    q = (rng.integers(0, 3, qlen) + 97).astype(np.uint8)
    chars, offsets = make_corpus(rng, 3000, [0, 1, 2, 3, 4, 5, 15, 16, 17, 33, 64, 100, max(qlen, 1), qlen + 1], alphabet=3,
                                 query=q, near_frac=0.5)
    corpus = rf.Corpus(chars, offsets)
    for m, extra in (("prefix", {}), ("postfix", {}), ("hamming", {"pad": True})):
        for kind in ALL_KINDS:
            check2 = lambda **kw: assert_same(_gpu_pad(m, kind, q, corpus, **kw, **extra),
                                              orc.batch(m, kind, q, chars, offsets, nthreads=0, **kw, **extra), (m, kind, kw, qlen))
            check2()
            for c in ((0, 1, 2, 5, 40, 2**64 - 1) if kind in ("distance", "similarity") else (0.0, 0.2, 0.5, 0.9, 1.0)):
                check2(cutoff=c)
    # Hamming without pad: candidates of another length are Err(DifferentLengthArgs)
    exp = orc.batch("hamming", "distance", q, chars, offsets, nthreads=0, allow_differing=True)
    lens = np.diff(offsets.astype(np.int64))
    if np.any(lens != qlen):
        with pytest.raises(rf.RfError) as ei:
            _gpu_pad("hamming", "distance", q, corpus)
        assert ei.value.status == _ffi.RF_ERR_INVALID_ARG and "Differing length" in str(ei.value)
    same = np.nonzero(lens == qlen)[0]
    if len(same):
        sub_off = np.zeros(len(same) + 1, np.uint64)
        sub_off[1:] = np.cumsum(lens[same])
        sub_chars = np.concatenate([chars[int(offsets[i]):int(offsets[i + 1])] for i in same] + [np.zeros(0, np.uint8)]).astype(np.uint8)
        sc = rf.Corpus(sub_chars, sub_off)
        assert_same(_gpu_pad("hamming", "distance", q, sc), exp[same], "hamming equal lengths")
        sc.close()
    corpus.close()


@pytest.mark.parametrize("qlen", [0, 1, 2, 5, 32, 64, 65, 200])
def test_damerau_levenshtein_vs_oracle(qlen):
    rng = np.random.default_rng(900 + qlen)
    q = (rng.integers(0, 3, qlen) + 97).astype(np.uint8)
    chars, offsets = make_corpus(rng, 1500, [0, 1, 2, 3, 8, 20, 33, 64, 70, max(qlen, 1), qlen + 2], alphabet=3, query=q, near_frac=0.5)
    corpus = rf.Corpus(chars, offsets)
    for kind in ALL_KINDS:
        check("damerau_levenshtein", kind, q, chars, offsets, corpus)
    for c in (0, 1, 2, 3, 10, 64, 2**64 - 1):
        check("damerau_levenshtein", "distance", q, chars, offsets, corpus, cutoff=c)
    for c in (0.0, 0.3, 0.7, 1.0):
        check("damerau_levenshtein", "normalized_similarity", q, chars, offsets, corpus, cutoff=c)
        check("damerau_levenshtein", "normalized_distance", q, chars, offsets, corpus, cutoff=c)
    corpus.close()
    assert rf.distance.damerau_levenshtein.distance("CA", "ABC") == 2                      # damerau_levenshtein.rs:226
    assert rf.distance.damerau_levenshtein.distance("Иванко", "Петрунко") == 5            # :695-698 (u32 elements)


def _gpu_pad(metric, kind, q, corpus, cutoff=None, pad=False):
    b = _bc(metric, q)
    a = rf.Args().pad(pad)
    if cutoff is not None:
        a = a.score_cutoff(cutoff)
    try:
        r = b._score(kind, corpus, a)
    finally:
        b.close()
    if isinstance(r, np.ma.MaskedArray):
        return r.filled(np.nan if r.dtype == np.float64 else _ffi.NONE_U32)
    return r


def test_hamming_prefix_postfix_known_answers():
    assert rf.distance.hamming.distance("hamming", "humming") == 1                                   # hamming.rs:198
    assert rf.distance.hamming.distance("ham", "hamming", rf.Args().pad(True)) == 4                  # :622-625
    assert rf.distance.hamming.distance("hammers", "hamming", rf.Args().pad(True).score_cutoff(2)) is None   # :584-591
    with pytest.raises(rf.RfError):
        rf.distance.hamming.distance("ham", "hamming")                                               # :617-620
    assert rf.distance.hamming.distance("hamming", "h\u9999mm\u00fcng") == 2                         # :612 (u32 elements)
    assert rf.distance.prefix.similarity("prefix", "preference") == 4                                # prefix.rs:122
    assert rf.distance.postfix.BatchComparator("postfix").similarity("prefix") == 3                  # postfix.rs:256


def test_unsupported_is_loud():
    c = rf.Corpus.from_strings([b"abc"])
    assert rf.distance.levenshtein.BatchComparator(b"abcd").distance_with_args(c, rf.Args().weights(1, 2, 3)).tolist() == [2]
    with pytest.raises(rf.RfError) as ei:   # generic weights (Wagner-Fischer) are limited to queries of 200 000 elements
        rf.distance.levenshtein.BatchComparator(np.full(200_001, 97, np.uint8)).distance_with_args(c, rf.Args().weights(1, 2, 3))
    assert ei.value.status == _ffi.RF_ERR_UNSUPPORTED
    with pytest.raises(rf.RfError):
        rf.distance.levenshtein.BatchComparator(np.zeros(_ffi.RF_MAX_QUERY_LEN + 1, np.uint8))
    with pytest.raises(rf.RfError):   # integer-valued result requested through the f64 entry point
        b = rf.distance.levenshtein.BatchComparator(b"abc")
        out = np.zeros(1)
        _ffi.check(_ffi.lib().rf_batch_score_f64(b._h, c._h, 0, None, out.ctypes.data))
    c.close()


def test_large_corpus_properties():
    """2e6 synthetic candidates (config-2 shape): GPU vs multi-threaded oracle, plus size-independent
    properties: d(q,q-planted)==0 hits exist, |len1-len2| <= d <= max(len1,len2), d(q,c) symmetric under
    swapping roles for a sample, normalized_similarity == 1 - d/max."""
    q = synth.synth_query(2, 32)
    n = 2_000_000
    chars, offsets = synth.synth_corpus(2, q, n, 8, 64, 16)
    corpus = rf.Corpus(chars, offsets)
    d = gpu_batch("levenshtein", "distance", q, corpus)
    exp = orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=0)
    assert_same(d, exp, "large lev")
    lens = np.diff(offsets.astype(np.int64))
    assert np.all(d >= np.abs(lens - 32)) and np.all(d <= np.maximum(lens, 32))
    assert (d <= 16).sum() > n // 200      # planted near-matches are found
    ns = gpu_batch("levenshtein", "normalized_similarity", q, corpus)
    assert np.array_equal(ns, 1.0 - d / np.maximum(lens, 32))
    jw = gpu_batch("jaro_winkler", "normalized_similarity", q, corpus)
    expjw = orc.batch("jaro_winkler", "normalized_similarity", q, chars, offsets, nthreads=0)
    assert np.max(np.abs(jw - expjw)) <= 1e-6   # north-star tolerance
    assert_same(jw, expjw, "large jw (bit-exact)")
    # role swap on a sample: d(q, c) == d(c, q)
    cq = rf.Corpus.from_strings([bytes(q)])
    for i in range(0, n, n // 50):
        c = chars[int(offsets[i]):int(offsets[i + 1])]
        assert gpu_batch("levenshtein", "distance", c, cq)[0] == d[i]
    cq.close()
    corpus.close()


def _stream(metric, kind, q, chars, offsets, **kw):
    from rapidfuzz_b200._scorer import Args, BatchComparatorBase
    b = type("B", (BatchComparatorBase,), {"METRIC": metric})(q)
    a = Args()
    if kw.get("cutoff") is not None:
        a = a.score_cutoff(kw["cutoff"])
    try:
        return b.stream(kind, chars, offsets, a)
    finally:
        b.close()


@pytest.mark.parametrize("chunk_mb,chunk_kcand", [(64, 2048), (1, 2048), (64, 1), (1, 3)])
def test_streaming_matches_oracle(chunk_mb, chunk_kcand):
    """rf_batch_stream_*: host-resident CSR candidates, chunked H2D/scan/D2H pipeline; chunk boundaries at
    arbitrary (unaligned) byte offsets; u32 and u64 offsets; every family; ragged + empty candidates."""
    L = _ffi.lib()
    _ffi.check(L.rf_set_option(b"stream_chunk_mb", chunk_mb))
    _ffi.check(L.rf_set_option(b"stream_chunk_kcand", chunk_kcand))
    try:
        rng = np.random.default_rng(77)
        q = (rng.integers(0, 5, 29) + 97).astype(np.uint8)
        chars, offsets = make_corpus(rng, 20000, [0, 1, 3, 8, 20, 33, 64, 70, 300], alphabet=5, query=q)
        for off in (offsets, offsets.astype(np.uint32)):
            for m, kind, kw in (("levenshtein", "distance", {}), ("levenshtein", "distance", {"cutoff": 6}),
                                ("indel", "normalized_similarity", {}), ("osa", "distance", {}),
                                ("lcs_seq", "similarity", {}), ("jaro_winkler", "similarity", {}),
                                ("jaro", "normalized_distance", {"cutoff": 0.4})):
                got = _stream(m, kind, q, chars, off, **kw)
                exp = orc.batch(m, kind, q, chars, offsets, nthreads=0, **kw)
                assert_same(got, exp, ("stream", m, kind, kw, off.dtype))
        # multi-word query + cutoff (config 3 shape) and multi-word Jaro through the same pipeline
        q3 = synth.synth_query(3, 256)
        c3, o3 = synth.synth_corpus(3, q3, 30000, 64, 256, 48)
        assert_same(_stream("levenshtein", "distance", q3, c3, o3, cutoff=32),
                    orc.batch("levenshtein", "distance", q3, c3, o3, nthreads=0, cutoff=32), "stream mw")
        assert_same(_stream("jaro", "similarity", q3[:100], c3[: int(o3[2000])], o3[:2001]),
                    orc.batch("jaro", "similarity", q3[:100], c3[: int(o3[2000])], o3[:2001], nthreads=0), "stream jaro mw")
        # empty corpus / all-empty candidates
        assert len(_stream("levenshtein", "distance", q, np.zeros(0, np.uint8), np.zeros(1, np.uint64))) == 0
        z = np.zeros(6, np.uint64)
        assert_same(_stream("levenshtein", "distance", q, np.zeros(0, np.uint8), z),
                    orc.batch("levenshtein", "distance", q, np.zeros(0, np.uint8), z, nthreads=0), "stream empties")
    finally:
        _ffi.check(L.rf_set_option(b"stream_chunk_mb", 64))
        _ffi.check(L.rf_set_option(b"stream_chunk_kcand", 2048))


def test_streaming_large_equals_resident():
    """5e6 config-2 candidates: the streaming entry point and the resident-corpus path agree bit for bit."""
    q = synth.synth_query(2, 32)
    chars, offsets = synth.synth_corpus(2, q, 5_000_000, 8, 64, 16)
    corpus = rf.Corpus(chars, offsets)
    d = gpu_batch("levenshtein", "distance", q, corpus)
    corpus.close()
    s = _stream("levenshtein", "distance", q, chars, offsets.astype(np.uint32))
    assert np.array_equal(d, s)
    exp = orc.batch("levenshtein", "distance", q, chars[: int(offsets[300000])], offsets[:300001], nthreads=0)
    assert np.array_equal(s[:300000], exp)


def test_corpus_file_to_gpu(tmp_path):
    """Corpus file -> resident corpus (rf_corpus_create_from_file) and -> streaming scan straight from the mapping."""
    q = synth.synth_query(4, 32)
    chars, offsets = synth.synth_corpus(4, q, 50_000, 0, 64, 16)
    path = str(tmp_path / "c.rfc")
    rf.write_corpus_file(path, chars, offsets)
    exp = orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=0)
    corpus = rf.Corpus.from_file(path)
    assert len(corpus) == 50_000 and corpus.total_chars == len(chars)
    assert np.array_equal(gpu_batch("levenshtein", "distance", q, corpus), exp)
    corpus.close()
    with rf.CorpusFile(path) as f:
        assert np.array_equal(_stream("levenshtein", "distance", q, f.chars, f.offsets), exp)
    strings = [bytes(chars[int(offsets[i]):int(offsets[i + 1])]) for i in range(2000)]
    pc, po = rf.pack_strings(strings)
    c2 = rf.Corpus(pc, po)
    assert np.array_equal(gpu_batch("levenshtein", "distance", q, c2), exp[:2000])
    c2.close()


def _bc(metric, q):
    from rapidfuzz_b200._scorer import BatchComparatorBase
    return type("B", (BatchComparatorBase,), {"METRIC": metric})(q)


@pytest.mark.parametrize("n", [1, 37, 5000, 300_000])
def test_extract_and_filter_vs_oracle(n):
    """On-device post-processing: k best by (score best-first, index ascending) and cutoff compaction in index
    order, against the oracle's full score vector sorted / filtered with numpy."""
    q = synth.synth_query(9, 32)
    chars, offsets = synth.synth_corpus(9, q, n, 8, 64, 16)
    corpus = rf.Corpus(chars, offsets)
    idxs = np.arange(n)
    for metric, kind, cut in (("levenshtein", "distance", None), ("levenshtein", "distance", 12),
                              ("levenshtein", "normalized_similarity", None), ("indel", "similarity", None),
                              ("lcs_seq", "similarity", 10), ("osa", "distance", None),
                              ("jaro_winkler", "similarity", None), ("jaro", "normalized_distance", 0.45),
                              ("ratio", "similarity", 0.4)):
        kw = {} if cut is None else {"cutoff": cut}
        exp = orc.batch(metric, kind, q, chars, offsets, nthreads=0, **kw)
        is_f = exp.dtype == np.float64
        valid = ~np.isnan(exp) if is_f else (exp != 0xFFFFFFFF)
        desc = kind in ("similarity", "normalized_similarity") or metric == "ratio"
        order = np.lexsort((idxs, -exp.astype(np.float64) if desc else exp.astype(np.float64)))
        order = order[valid[order]]
        b = _bc(metric, q)
        a = rf.Args() if cut is None else rf.Args().score_cutoff(cut)
        for k in (1, 5, 64, 1000):
            gi, gs = b.extract(kind, corpus, k=k, args=a)
            e = order[:k]
            assert np.array_equal(gi, e.astype(np.uint32)), (metric, kind, cut, k, n)
            assert np.array_equal(gs, exp[e]), (metric, kind, cut, k, n)
        if cut is not None:
            hits = np.nonzero(valid)[0]
            for cap in (None, 3, 0):
                gi, gs, tot = b.filter(kind, corpus, a, capacity=cap)
                m = len(hits) if cap is None else min(cap, len(hits))
                assert tot == len(hits), (metric, kind, cut, cap, tot, len(hits))
                assert np.array_equal(gi, hits[:m].astype(np.uint32)) and np.array_equal(gs, exp[hits[:m]])
        b.close()
    corpus.close()


def _oracle_topk(queries, chars, offsets, k, cutoff=None):
    n = len(offsets) - 1
    idx = np.full((len(queries), k), 0xFFFFFFFF, dtype=np.uint32)
    dist = np.full((len(queries), k), 0xFFFFFFFF, dtype=np.uint32)
    for qi, q in enumerate(queries):
        d = orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=0).astype(np.int64)
        keys = d * (1 << 32) + np.arange(n)
        if cutoff is not None:
            keys = keys[d <= cutoff]
        keys = np.sort(keys)[:k]
        idx[qi, :len(keys)] = (keys & 0xFFFFFFFF).astype(np.uint32)
        dist[qi, :len(keys)] = (keys >> 32).astype(np.uint32)
    return idx, dist


@pytest.mark.parametrize("qlens,n,k", [([32], 50000, 10), ([0, 1, 7, 20, 32], 20000, 10), ([5, 33, 40, 64], 30000, 7),
                                        ([32], 100, 10), ([16], 5, 10), ([32], 300000, 64)])
def test_cdist_topk_vs_oracle_full_matrix(qlens, n, k):
    """config 5 shape (small): per-query top-k by (distance, index) against the oracle's full matrix."""
    rng = np.random.default_rng(n + k)
    queries = []
    for rep in range(12):
        for ql in qlens:
            queries.append(synth.synth_query(1000 + rep, ql))
    base = queries[0] if len(queries[0]) else synth.synth_query(5, 32)
    chars, offsets = synth.synth_corpus(5, base, n, 8, 64, 16)
    corpus = rf.Corpus(chars, offsets)
    for cutoff in (None, 20, 3):
        idx, dist = rf.cdist_topk(queries, corpus, k=k, score_cutoff=cutoff)
        eidx, edist = _oracle_topk(queries, chars, offsets, k, cutoff)
        assert np.array_equal(dist, edist), (qlens, n, k, cutoff)
        assert np.array_equal(idx, eidx), (qlens, n, k, cutoff)
    corpus.close()


@pytest.mark.parametrize("slices,skip", [(1, 1), (1, 0), (3, 1), (17, 1), (256, 1), (0, 0)])
def test_cdist_work_decomposition_is_invisible(slices, skip):
    """The (slice, query) work units, the merge of several slices and the skipping of groups by length against the
    running k-th bound must not change a single entry: 330 queries of mixed length (more units than CTAs when there is
    one slice), ties on the distance decided by the index, with and without a cutoff."""
    queries = [synth.synth_query(2000 + i, (8, 20, 32, 32, 32, 47, 64, 3)[i % 8]) for i in range(330)]
    chars, offsets = synth.synth_corpus(11, queries[2], 40000, 1, 64, 16)
    corpus = rf.Corpus(chars, offsets)
    L = _ffi.lib()
    _ffi.check(L.rf_set_option(b"cdist_slices", slices))
    _ffi.check(L.rf_set_option(b"cdist_skip", skip))
    try:
        for k, cutoff in ((10, None), (1, None), (33, 12), (5, 0)):
            idx, dist = rf.cdist_topk(queries, corpus, k=k, score_cutoff=cutoff)
            eidx, edist = _oracle_topk(queries, chars, offsets, k, cutoff)
            assert np.array_equal(dist, edist), (slices, skip, k, cutoff)
            assert np.array_equal(idx, eidx), (slices, skip, k, cutoff)
    finally:
        _ffi.check(L.rf_set_option(b"cdist_slices", 0))
        _ffi.check(L.rf_set_option(b"cdist_skip", 1))
    corpus.close()


@pytest.mark.parametrize("world,k", [(3, 10), (8, 4), (2, 64)])
def test_sharded_cdist_device_merge(world, k):
    """SURVEY 8e on one GPU: the corpus is cut into `world` byte-balanced shards, every shard is scored on its own
    (what each rank does), the per-shard lists are stacked the way one all_gather_into_tensor lays them out and merged
    on the device (rf_topk_merge_device).  Must equal the oracle's global top-k, ties by GLOBAL index."""
    import torch
    from rapidfuzz_b200 import sharding
    queries = [synth.synth_query(3000 + i, (32, 12, 50)[i % 3]) for i in range(40)]
    q_off = np.zeros(len(queries) + 1, dtype=np.uint64)
    q_off[1:] = np.cumsum([len(q) for q in queries])
    q_chars = np.concatenate(queries)
    chars, offsets = synth.synth_corpus(12, queries[0], 30000, 1, 64, 16)
    for cutoff in (None, 14):
        parts, starts = [], []
        for r in range(world):
            c, o, lo = sharding.local_shard(chars, offsets, world, r)
            corpus = rf.Corpus(c, o)
            i, d = sharding.cdist_topk_device(q_chars, q_off, corpus, k=k, score_cutoff=cutoff)
            parts.append(torch.stack([i, d], dim=0))
            starts.append(lo)
            corpus.close()
        gi, gd = sharding.merge_topk_device(torch.stack(parts, dim=0).contiguous(),
                                            torch.tensor(starts, dtype=torch.int64, device="cuda"), k)
        torch.cuda.synchronize()
        eidx, edist = _oracle_topk(queries, chars, offsets, k, cutoff)
        assert np.array_equal(gd.cpu().numpy().view(np.uint32), edist), (world, k, cutoff)
        got = gi.cpu().numpy()
        exp = np.where(eidx == 0xFFFFFFFF, -1, eidx.astype(np.int64))
        assert np.array_equal(got, exp), (world, k, cutoff)


def test_cpp_host_mirror_known_answers(tmp_path):
    """Compiles tests/cpp/test_cpp_api.cpp against the header-only C++ mirror and runs it on the GPU."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "test_cpp_api")
    libdir = os.path.join(root, "rapidfuzz-rs_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           "-I", os.path.join(root, "rapidfuzz-rs_b200", "cpp"),
                           os.path.join(root, "tests", "cpp", "test_cpp_api.cpp"), "-o", exe,
                           "-L", libdir, "-lrfgpu", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all ok" in out.stdout
