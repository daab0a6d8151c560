"""Round-2 GPU parity tests (through the C ABI, against the CPU oracle): the fixes of ADVICE r1 and the round's new
entry points."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import rapidfuzz_b200 as rf
import synth
from rapidfuzz_b200 import _ffi
from rapidfuzz_b200._scorer import Args, BatchComparatorBase
from oracle import oracle as orc
from gpu_util import make_corpus, check, gpu_batch, assert_same


def _bc(metric, q, device=0):
    return type("B", (BatchComparatorBase,), {"METRIC": metric})(q, device)


@pytest.mark.parametrize("qlen", [1, 29, 64, 200])
def test_streaming_with_u32_query_renames_the_candidates(qlen):
    """ADVICE r1 (rf_api.cu stream_impl): a comparator made by rf_batch_create_u32 keeps its tables over renamed bytes;
    the streaming entry points must rename the candidates' bytes the same way (they used to score the raw bytes)."""
    rng = np.random.default_rng(qlen)
    qb = (rng.integers(0, 6, qlen) + 97).astype(np.uint8)
    chars, offsets = make_corpus(rng, 6000, [0, 1, 8, 20, 33, 64, 70, 300], alphabet=6, query=qb, high_bytes=True)
    # (a) all-byte symbols given as u32: must equal the byte query; (b) symbols beyond a byte never match a byte
    for q32 in (qb.astype(np.uint32), np.where(np.arange(qlen) % 3 == 0, 0x4E2D, qb).astype(np.uint32)):
        for m, kind, kw in (("levenshtein", "distance", {}), ("levenshtein", "distance", {"cutoff": 7}),
                            ("indel", "normalized_similarity", {}), ("jaro_winkler", "similarity", {}),
                            ("osa", "distance", {})):
            b = _bc(m, q32)
            a = Args().score_cutoff(kw["cutoff"]) if kw else Args()
            for off in (offsets, offsets.astype(np.uint32)):
                got = b.stream(kind, chars, off, a)
                exp = orc.batch(m, kind, q32, chars.astype(np.uint32), offsets, nthreads=0, **kw)
                assert_same(got, exp, ("stream u32 query", m, kind, kw, qlen, off.dtype))
            b.close()


def test_cdist_rejects_u32_corpora_loudly():
    """ADVICE r1 (rf_api.cu cdist_impl): a corpus made by rf_corpus_create_u32 holds dictionary codes (or u32 elements),
    which rf_cdist_topk_u8 must not scan with tables built over raw bytes."""
    elems = np.array([97, 98, 99, 0x4E2D, 97, 98], dtype=np.uint32)
    offs = np.array([0, 3, 6], dtype=np.uint64)
    for compact in (1, 0):
        _ffi.check(_ffi.lib().rf_set_option(b"compact_u32_corpus", compact))
        try:
            c = rf.Corpus.from_u32(elems, offs)
            with pytest.raises(rf.RfError) as ei:
                rf.cdist_topk([b"abc"], c, k=1)
            assert ei.value.status == _ffi.RF_ERR_UNSUPPORTED
            c.close()
        finally:
            _ffi.check(_ffi.lib().rf_set_option(b"compact_u32_corpus", 1))


def test_corrupt_offsets_are_refused_not_scanned():
    """ADVICE r1 (rf_io.cpp / corpus_create_host): decreasing or out-of-range CSR starts -> RF_ERR_INVALID_ARG, and the
    process stays usable (no sticky CUDA fault)."""
    q = synth.synth_query(1, 16)
    chars, offsets = synth.synth_corpus(1, q, 5000, 1, 40, 4)
    for dt in (np.uint64, np.uint32):
        for pos, val in ((100, int(offsets[102]) + 1), (4000, int(offsets[5000]) + 77), (7, 2**32 + 5 if dt == np.uint64 else 2**32 - 1)):
            bad = offsets.astype(dt).copy()
            bad[pos] = val
            with pytest.raises(rf.RfError) as ei:
                rf.Corpus(chars, bad)
            assert ei.value.status == _ffi.RF_ERR_INVALID_ARG, (dt, pos)
    elems = chars.astype(np.uint32)
    bad = offsets.copy()
    bad[9] = bad[11] + 3
    with pytest.raises(rf.RfError):
        rf.Corpus.from_u32(elems, bad)
    check("levenshtein", "distance", q, chars, offsets)   # still healthy


def test_compact_sub_cache_survives_many_corpora():
    """ADVICE r1 (compact_sub): the per-dictionary sub-comparators are no longer evicted under a running call; 70 compact
    corpora against one u32 comparator, then the first one again."""
    q = np.array([0x4E2D, 0x6587, 97, 98, 0x4E2D], dtype=np.uint32)
    b = _bc("levenshtein", q)
    corpora = []
    for i in range(70):
        elems = np.array([0x4E2D, 0x6587, 97 + (i % 20), 98, 0x4E2D, 1000 + i], dtype=np.uint32)
        corpora.append((rf.Corpus.from_u32(elems, np.array([0, 5, 6], dtype=np.uint64)), elems))
    for c, elems in corpora + corpora[:3]:
        got = b.distance(c)
        exp = orc.batch("levenshtein", "distance", q, elems, np.array([0, 5, 6], dtype=np.uint64), nthreads=0)
        assert np.array_equal(got, exp)
    for c, _ in corpora:
        c.close()
    b.close()


# ------------------------------------------------------------------------------------------------ sharded C ABI
def _device_lists():
    """One-GPU boxes exercise the split / gather / merge logic with a repeated device (copy-based gathers); with two or
    more GPUs the same tests also run over distinct devices, i.e. through NCCL."""
    n = _ffi.lib().rf_device_count()
    lists = [[0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    return lists


@pytest.mark.parametrize("devices", _device_lists() if True else [], ids=lambda d: "dev" + "".join(map(str, d)))
def test_sharded_corpus_equals_single_gpu_equals_oracle(devices):
    """the multi-GPU split behind the C ABI.  sharded == single-GPU == oracle for scores (host vector and
    all-gathered device vectors), extract and cdist top-k with global indices."""
    import torch
    from rapidfuzz_b200 import sharding
    rng = np.random.default_rng(len(devices))
    q = synth.synth_query(2, 32)
    chars, offsets = synth.synth_corpus(2, q, 120_001, 8, 64, 16)
    sc = sharding.ShardedCorpus(chars, offsets, devices)
    assert len(sc) == 120_001 and sc.uses_nccl == (len(devices) > 1 and len(set(devices)) == len(devices))
    rg = sc.shard_ranges()
    assert rg[0][0] == 0 and rg[-1][1] == 120_001 and all(a[1] == b[0] for a, b in zip(rg[:-1], rg[1:]))
    byts = [int(offsets[b] - offsets[a]) for a, b in rg]
    assert max(byts) - min(byts) <= 2 * 64
    single = rf.Corpus(chars, offsets)
    for metric, kind, cut in (("levenshtein", "distance", None), ("levenshtein", "distance", 9), ("indel", "normalized_similarity", None),
                              ("jaro_winkler", "similarity", 0.6), ("lcs_seq", "similarity", None)):
        sb = sharding.ShardedBatchComparator(metric, q, devices)
        a = Args().score_cutoff(cut) if cut is not None else Args()
        got = sb.score(kind, sc, a)
        kw = {} if cut is None else {"cutoff": cut}
        exp = orc.batch(metric, kind, q, chars, offsets, nthreads=0, **kw)
        assert_same(got, exp, ("sharded score", metric, kind, cut, devices))
        assert_same(gpu_batch(metric, kind, q, single, **kw), exp, "single")
        # all-gather: every device ends with the whole vector
        dt = torch.float64 if got.dtype == np.float64 else torch.int32
        bufs = [torch.full((len(sc),), -1, dtype=dt, device="cuda:%d" % d) for d in devices]
        sb.score_allgather(kind, sc, [b.data_ptr() for b in bufs], a)
        for b in bufs:
            gb = b.cpu().numpy()
            assert_same(gb if got.dtype == np.float64 else gb.view(np.uint32), exp, ("allgather", metric, devices))
        # extract: k best with global indices
        gi, gs = sb.extract(kind, sc, k=7, args=a)
        desc = kind in ("similarity", "normalized_similarity")
        valid = ~np.isnan(exp) if exp.dtype == np.float64 else (exp != 0xFFFFFFFF)
        keyv = np.where(valid, -exp.astype(np.float64) if desc else exp.astype(np.float64), np.inf)
        order = np.lexsort((np.arange(len(exp)), keyv))[:7]
        order = order[valid[order]]
        assert np.array_equal(gi, order.astype(np.uint64)) and np.array_equal(gs, exp[order]), (metric, kind)
        sb.close()
    # many-vs-many top-k over the shards == the oracle's global top-k
    qs = [synth.synth_query(100 + i, L) for i, L in enumerate((32, 20, 7, 32, 1))]
    q_chars = np.concatenate(qs)
    q_off = np.zeros(len(qs) + 1, dtype=np.uint64)
    q_off[1:] = np.cumsum([len(x) for x in qs])
    for k, cut in ((10, None), (3, 14)):
        gi, gd = sharding.sharded_cdist_topk(q_chars, q_off, sc, k=k, score_cutoff=cut)
        for qi, qq in enumerate(qs):
            d = orc.batch("levenshtein", "distance", qq, chars, offsets, nthreads=0).astype(np.int64)
            keys = np.sort(d * (1 << 32) + np.arange(len(d)))
            if cut is not None:
                keys = keys[(keys >> 32) <= cut]
            keys = keys[:k]
            m = len(keys)
            assert np.array_equal(gi[qi][:m], (keys & 0xFFFFFFFF).astype(np.uint64)) and np.array_equal(gd[qi][:m], (keys >> 32).astype(np.uint32))
            assert np.all(gi[qi][m:] == np.uint64(0xFFFFFFFFFFFFFFFF)) and np.all(gd[qi][m:] == 0xFFFFFFFF)
    single.close()
    sc.close()


@pytest.mark.parametrize("devices", _device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_sharded_stream_and_edge_shapes(devices):
    from rapidfuzz_b200 import sharding
    rng = np.random.default_rng(5)
    q = (rng.integers(0, 5, 29) + 97).astype(np.uint8)
    chars, offsets = make_corpus(rng, 30_000, [0, 1, 3, 8, 20, 33, 64, 70, 300], alphabet=5, query=q)
    sb = sharding.ShardedBatchComparator("levenshtein", q, devices)
    exp = orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=0)
    assert_same(sb.stream("distance", chars, offsets), exp, "sharded stream")
    # ... and the dynamically balanced forms: one length byte per candidate, optional byte results / 6-bit packed characters
    L = _ffi.lib()
    for mb, kc in ((64, 2048), (1, 4)):
        _ffi.check(L.rf_set_option(b"stream_chunk_mb", mb))
        _ffi.check(L.rf_set_option(b"stream_chunk_kcand", kc))
        try:
            c8, o8 = make_corpus(np.random.default_rng(6), 30_000, [0, 1, 3, 8, 20, 33, 64, 70, 255], alphabet=5, query=q)
            e8 = orc.batch("levenshtein", "distance", q, c8, o8, nthreads=0)
            lens8 = np.diff(o8.astype(np.int64)).astype(np.uint8)
            assert_same(sb.stream_len8("distance", c8, lens8), e8, ("sharded len8", mb))
            assert int(e8.max()) <= 254
            assert np.array_equal(sb.stream_len8("distance", c8, lens8, u8_results=True), e8.astype(np.uint8)), ("sharded len8 u8", mb)
            packed, d64 = rf.pack6(c8)
            assert np.array_equal(sb.stream_len8("distance", packed, lens8, u8_results=True, dict64=d64), e8.astype(np.uint8)), ("sharded packed6", mb)
        finally:
            _ffi.check(L.rf_set_option(b"stream_chunk_mb", 64))
            _ffi.check(L.rf_set_option(b"stream_chunk_kcand", 2048))
    # fewer candidates than shards, empty corpus, all-empty candidates
    for n in (0, 1, 2):
        sc = sharding.ShardedCorpus(chars[: int(offsets[n])], offsets[: n + 1], devices)
        assert_same(sb.score("distance", sc), exp[:n], ("tiny", n))
        gi, gd = sharding.sharded_cdist_topk(q, np.array([0, len(q)], np.uint64), sc, k=2)
        order = np.lexsort((np.arange(n), exp[:n]))[:2]
        assert np.array_equal(gi[0][: len(order)], order.astype(np.uint64))
        sc.close()
    z = np.zeros(6, np.uint64)
    sc = sharding.ShardedCorpus(np.zeros(0, np.uint8), z, devices)
    assert_same(sb.score("distance", sc), np.full(5, 29, np.uint32), "empties")
    sc.close()
    # mismatched device lists are refused
    if len(devices) > 1:
        other = sharding.ShardedBatchComparator("levenshtein", q, devices[:1])
        sc = sharding.ShardedCorpus(chars, offsets, devices)
        with pytest.raises(rf.RfError) as ei:
            other.score("distance", sc)
        assert ei.value.status == _ffi.RF_ERR_INVALID_ARG
        other.close()
        sc.close()
    sb.close()


# ------------------------------------------------------------------------------------------------ long / multi-word queries
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.timeout(900)
def test_ocr_large_band_on_gpu():
    """the reference's own test_large_band (levenshtein.rs:2139-2161: 106 514 x 107 244 -> 5278, None at
    cutoff 2500, 5278 with score_hint 0) through rf_batch_score_u32, i.e. the long-query stripe kernel."""
    z = np.load(os.path.join(GOLD, "ocr.npz"))
    a, b = z["OCR_EXAMPLE1"], z["OCR_EXAMPLE2"]
    assert len(a) == 106514 and len(b) == 107244
    lev = rf.distance.levenshtein.BatchComparator(a)
    assert lev.distance(b) == 5278
    assert lev.distance_with_args(b, rf.Args().score_cutoff(2500)) is None
    assert lev.distance_with_args(b, rf.Args().score_hint(0)) == 5278
    assert lev.distance_with_args(b, rf.Args().score_cutoff(5278).score_hint(31)) == 5278
    assert lev.distance_with_args(b, rf.Args().score_cutoff(5277)) is None
    # the same query against a ragged corpus, every integer family, vs the oracle (which runs the reference's block + band code)
    cands = [b, a, b[:50000], np.zeros(0, np.uint8), a[:200], a[1000:90000], b[::-1].copy()]
    chars = np.concatenate(cands).astype(np.uint8)
    offsets = np.zeros(len(cands) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(c) for c in cands])
    corpus = rf.Corpus(chars, offsets)
    for m, kind, kw in (("levenshtein", "distance", {}), ("levenshtein", "normalized_similarity", {}), ("indel", "distance", {}),
                        ("lcs_seq", "similarity", {}), ("osa", "distance", {}), ("levenshtein", "distance", {"cutoff": 60000}),
                        ("levenshtein", "distance", {"cutoff": 40})):
        got = gpu_batch(m, kind, a, corpus, **kw)
        exp = orc.batch(m, kind, a, chars, offsets, nthreads=0, **kw)
        assert_same(got, exp, ("ocr", m, kind, kw))
    corpus.close()
    lev.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("qlen", [16384, 16385, 16449, 20000, 32768, 40001])
def test_queries_beyond_16384_vs_oracle(qlen):
    """The 16 384 cap of round 1 is gone: stripes of 256 blocks, carries parked between stripes (1, 2 and 3 stripes,
    stripe ends on and off block boundaries)."""
    rng = np.random.default_rng(qlen)
    q = (rng.integers(0, 4, qlen) + 97).astype(np.uint8)
    lens = [0, 1, 63, 64, 65, 5000, qlen - 1, qlen, qlen + 1, qlen + 777]
    chars, offsets = make_corpus(rng, 40, lens, alphabet=4, query=q, near_frac=0.5)
    corpus = rf.Corpus(chars, offsets)
    for m in ("levenshtein", "indel", "lcs_seq", "osa"):
        check(m, "distance", q, chars, offsets, corpus)
    check("levenshtein", "normalized_similarity", q, chars, offsets, corpus, cutoff=0.9)
    check("levenshtein", "distance", q, chars, offsets, corpus, cutoff=1000)
    check("levenshtein", "distance", q, chars, offsets, corpus, cutoff=12)      # banded kernel: any query length
    check("levenshtein", "distance", q, chars, offsets, corpus, weights=(1, 1, 2))
    check("ratio", "similarity", q, chars, offsets, corpus)
    corpus.close()


@pytest.mark.parametrize("qlen", [65, 96, 127, 128, 129, 191, 192, 193, 255, 256, 257, 320, 383, 384, 385, 448, 511, 512, 513])
def test_register_column_kernel_vs_shuffle_kernel_vs_oracle(qlen):
    """queries of 65..512 elements on a resident corpus run one thread per candidate with the whole
    bit-vector column in registers (scan_lbn_kernel, 4/8/12/16 limbs); multi_word_path=1 is the sub-warp shuffle
    kernel.  Both must equal the oracle, limb and plane boundaries included."""
    rng = np.random.default_rng(1000 + qlen)
    q = (rng.integers(0, 5, qlen) + 97).astype(np.uint8)
    lens = [0, 1, 7, 8, 9, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 700, qlen - 1, qlen, qlen + 1]
    chars, offsets = make_corpus(rng, 3000, lens, alphabet=5, query=q, near_frac=0.5, high_bytes=True)
    corpus = rf.Corpus(chars, offsets)
    try:
        for path in (0, 1):
            _ffi.check(_ffi.lib().rf_set_option(b"multi_word_path", path))
            for m in ("levenshtein", "indel", "lcs_seq", "osa"):
                check(m, "distance", q, chars, offsets, corpus)
            check("levenshtein", "normalized_distance", q, chars, offsets, corpus)
            check("levenshtein", "distance", q, chars, offsets, corpus, cutoff=70)
            check("levenshtein", "similarity", q, chars, offsets, corpus)
            check("osa", "normalized_similarity", q, chars, offsets, corpus, cutoff=0.5)
            check("ratio", "similarity", q, chars, offsets, corpus, cutoff=0.3)
            check("levenshtein", "distance", q, chars, offsets, corpus, weights=(2, 2, 2))
            check("levenshtein", "distance", q, chars, offsets, corpus, weights=(1, 1, 2))
    finally:
        _ffi.check(_ffi.lib().rf_set_option(b"multi_word_path", 0))
        corpus.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("qlen", [2049, 2112, 3000, 8191, 20000])
def test_jaro_queries_beyond_2048_vs_oracle(qlen):
    """VERDICT r1 missing #6: Jaro / Jaro-Winkler had a 2048-element query cap; the reference has none
    (jaro.rs:286-337, :370-420).  Warp-per-candidate kernel with the flags in scratch, vs the oracle's block path."""
    rng = np.random.default_rng(qlen)
    q = (rng.integers(0, 6, qlen) + 97).astype(np.uint8)
    lens = [0, 1, 2, 64, 65, 1000, qlen // 2, qlen - 1, qlen, qlen + 1, 2 * qlen + 5]
    chars, offsets = make_corpus(rng, 60, lens, alphabet=6, query=q, near_frac=0.5)
    corpus = rf.Corpus(chars, offsets)
    for m in ("jaro", "jaro_winkler"):
        for kind in ("similarity", "normalized_distance"):
            check(m, kind, q, chars, offsets, corpus)
        check(m, "similarity", q, chars, offsets, corpus, cutoff=0.7)
        check(m, "distance", q, chars, offsets, corpus, cutoff=0.25)
    corpus.close()
    # the streaming entry point takes the same kernel
    b = _bc("jaro_winkler", q)
    assert_same(b.stream("similarity", chars, offsets), orc.batch("jaro_winkler", "similarity", q, chars, offsets, nthreads=0), "jaro long stream")
    b.close()


@pytest.mark.parametrize("chunk_mb,chunk_kcand", [(64, 2048), (1, 2048), (64, 4), (1, 8)])
def test_streaming_with_length_bytes_and_byte_results(chunk_mb, chunk_kcand):
    """rf_batch_stream_{u32,u8}_len8: one length byte per candidate on the wire, offsets rebuilt on the device, optional
    byte results; must equal the CSR streaming entry point and the oracle for every chunking."""
    L = _ffi.lib()
    _ffi.check(L.rf_set_option(b"stream_chunk_mb", chunk_mb))
    _ffi.check(L.rf_set_option(b"stream_chunk_kcand", chunk_kcand))
    try:
        rng = np.random.default_rng(11)
        q = (rng.integers(0, 5, 29) + 97).astype(np.uint8)
        chars, offsets = make_corpus(rng, 50_001, [0, 1, 3, 8, 20, 33, 64, 70, 255], alphabet=5, query=q)
        lens = np.diff(offsets.astype(np.int64)).astype(np.uint8)
        for m, kind, kw in (("levenshtein", "distance", {}), ("levenshtein", "distance", {"cutoff": 6}), ("indel", "distance", {}),
                            ("lcs_seq", "similarity", {}), ("osa", "distance", {}), ("hamming", "distance", {"pad": True})):
            b = _bc(m, q)
            a = Args()
            if "cutoff" in kw:
                a = a.score_cutoff(kw["cutoff"])
            if kw.get("pad"):
                a = a.pad(True)
            exp = orc.batch(m, kind, q, chars, offsets, nthreads=0, **kw)
            assert_same(b.stream_len8(kind, chars, lens, a), exp, ("len8", m, kind, kw))
            if exp[exp != 0xFFFFFFFF].max(initial=0) <= 254:
                got8 = b.stream_len8(kind, chars, lens, a, u8_results=True)
                exp8 = np.where(exp == 0xFFFFFFFF, 255, exp).astype(np.uint8)
                assert np.array_equal(got8, exp8), ("len8 u8", m, kind, kw)
            b.close()
        # a u32-query comparator renames the bytes here too
        q32 = np.where(np.arange(29) % 4 == 0, 0x4E2D, q).astype(np.uint32)
        b = _bc("levenshtein", q32)
        assert_same(b.stream_len8("distance", chars, lens), orc.batch("levenshtein", "distance", q32, chars.astype(np.uint32), offsets, nthreads=0), "len8 wide")
        b.close()
        # scores above 254 cannot come back as bytes: loud
        b = _bc("levenshtein", np.full(300, 97, np.uint8))
        with pytest.raises(rf.RfError) as ei:
            b.stream_len8("distance", chars, lens, u8_results=True)
        assert ei.value.status == _ffi.RF_ERR_INVALID_ARG
        assert_same(b.stream_len8("distance", chars, lens), orc.batch("levenshtein", "distance", np.full(300, 97, np.uint8), chars, offsets, nthreads=0), "len8 long query")
        b.close()
        # empty input
        b = _bc("levenshtein", q)
        assert len(b.stream_len8("distance", np.zeros(0, np.uint8), np.zeros(0, np.uint8))) == 0
        b.close()
    finally:
        _ffi.check(L.rf_set_option(b"stream_chunk_mb", 64))
        _ffi.check(L.rf_set_option(b"stream_chunk_kcand", 2048))


# ------------------------------------------------------------------------------------------------ concurrency / options
def test_eight_threads_share_handles():
    """rfgpu.h promises that concurrent calls on shared handles are safe (BatchComparator is
    Clone + Send + Sync in the reference, levenshtein.rs:1635-1639).  8 host threads share ONE corpus and the same
    comparators and mix score / extract / filter / stream / cdist / u32-renaming calls (ctypes drops the GIL inside the
    library); every result must equal the oracle."""
    from rapidfuzz_b200 import sharding
    q = synth.synth_query(21, 32)
    q2 = synth.synth_query(22, 200)
    chars, offsets = synth.synth_corpus(21, q, 200_000, 8, 64, 16)
    lens = np.diff(offsets.astype(np.int64)).astype(np.uint8)
    corpus = rf.Corpus(chars, offsets)
    lev, jw, ind, mw = _bc("levenshtein", q), _bc("jaro_winkler", q), _bc("indel", q), _bc("levenshtein", q2)
    wide = _bc("levenshtein", q.astype(np.uint32))
    exp_lev = orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=0)
    exp_cut = orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=0, cutoff=10)
    exp_jw = orc.batch("jaro_winkler", "similarity", q, chars, offsets, nthreads=0)
    exp_ind = orc.batch("indel", "normalized_similarity", q, chars, offsets, nthreads=0)
    exp_mw = orc.batch("levenshtein", "distance", q2, chars, offsets, nthreads=0)
    order = np.lexsort((np.arange(len(exp_lev)), exp_lev))[:9]
    hits = np.nonzero(exp_cut != 0xFFFFFFFF)[0]
    qs = [synth.synth_query(300 + i, 32) for i in range(4)]
    exp_cd = []
    for qq in qs:
        d = orc.batch("levenshtein", "distance", qq, chars, offsets, nthreads=0).astype(np.int64)
        exp_cd.append(np.sort(d * (1 << 32) + np.arange(len(d)))[:5])
    errors = []

    def work(tid):
        try:
            for it in range(6):
                op = (tid + it) % 8
                if op == 0:
                    assert np.array_equal(lev.distance(corpus), exp_lev), "score"
                elif op == 1:
                    r = lev.distance_with_args(corpus, Args().score_cutoff(10))
                    assert np.array_equal(r.filled(0xFFFFFFFF), exp_cut), "cutoff"
                elif op == 2:
                    assert np.array_equal(jw.similarity(corpus), exp_jw), "jaro_winkler"
                elif op == 3:
                    gi, gs = lev.extract("distance", corpus, k=9)
                    assert np.array_equal(gi, order.astype(np.uint32)) and np.array_equal(gs, exp_lev[order]), "extract"
                    fi, fs, tot = lev.filter("distance", corpus, Args().score_cutoff(10))
                    assert tot == len(hits) and np.array_equal(fi, hits.astype(np.uint32)), "filter"
                elif op == 4:
                    assert np.array_equal(lev.stream("distance", chars, offsets), exp_lev), "stream"
                    assert np.array_equal(lev.stream_len8("distance", chars, lens, u8_results=True), exp_lev.astype(np.uint8)), "len8"
                elif op == 5:
                    gi, gd = rf.cdist_topk([bytes(x) for x in qs], corpus, k=5)
                    for qi, keys in enumerate(exp_cd):
                        assert np.array_equal(gi[qi], (keys & 0xFFFFFFFF).astype(np.uint32)) and np.array_equal(gd[qi], (keys >> 32).astype(np.uint32)), "cdist"
                elif op == 6:
                    assert np.array_equal(ind.normalized_similarity(corpus), exp_ind), "indel"
                    assert np.array_equal(wide.distance(corpus), exp_lev), "u32 query on a byte corpus"
                else:
                    assert np.array_equal(mw.distance(corpus), exp_mw), "multi-word"
        except BaseException as e:   # noqa: BLE001 -- reported by the main thread
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for b in (lev, jw, ind, mw, wide):
        b.close()
    corpus.close()


def gpu_batch_with(b, kind, corpus, cutoff=None, prefix_weight=0.1):
    """gpu_util.gpu_batch on an existing comparator (whose options were set by the test)."""
    a = Args()
    if cutoff is not None:
        a = a.score_cutoff(cutoff)
    r = b._score(kind, corpus, a.prefix_weight(prefix_weight))
    if isinstance(r, np.ma.MaskedArray):
        return r.filled(np.nan if r.dtype == np.float64 else _ffi.NONE_U32)
    return r


@pytest.mark.parametrize("qlen", [1, 2, 7, 32, 33, 47, 64])
def test_jaro_table_epilogue_equals_per_pair_epilogue_equals_oracle(qlen):
    """The row-wise Jaro kernels look the whole f64 score algebra up in a per-launch table indexed by (candidate length,
    common characters, transpositions / 2, prefix) -- built by the same device functions that option jaro32=2 runs per
    pair.  All four kinds, with and without cutoffs (incl. > 0.7, which back-translates through the prefix:
    jaro_winkler.rs:125-133), two prefix weights, 1 x 1 pairs, empty candidates, candidates up to the longest the
    row-wise kernels take and beyond (jaro.rs:553-565): table == per pair == generic routine == oracle, bit for bit."""
    L = _ffi.lib()
    rng = np.random.default_rng(900 + qlen)
    q = rng.integers(97, 101, qlen).astype(np.uint8)
    n = 6000
    lens = rng.choice([0, 1, 2, 3, 7, 8, 9, 31, 32, 33, 40, 63, 64, 65, 66, 67, 100, 128, 129, 130, 200], n)
    chars = rng.integers(97, 101, int(lens.sum())).astype(np.uint8)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    for i in range(0, n, 7):                       # near matches: shared prefixes, many common characters
        a, b = int(off[i]), int(off[i + 1])
        m = min(b - a, qlen)
        chars[a:a + m] = q[:m]
    corpus = rf.Corpus(chars, off)
    for metric in ("jaro", "jaro_winkler"):
        variants = []
        for opt in (1, 2, 3, 0):
            b = _bc(metric, q)
            _ffi.check(L.rf_batch_set_option(b._h, b"jaro32", opt))
            variants.append(b)
        for kind in ("similarity", "distance", "normalized_similarity", "normalized_distance"):
            for kw in ({}, {"cutoff": 0.0}, {"cutoff": 0.5}, {"cutoff": 0.71}, {"cutoff": 0.9}, {"cutoff": 1.0}, {"cutoff": 1.2},
                       {"prefix_weight": 0.25, "cutoff": 0.8}):
                if "prefix_weight" in kw and metric == "jaro":
                    continue
                exp = orc.batch(metric, kind, q, chars, off, nthreads=0, **kw)
                for opt, b in zip((1, 2, 3, 0), variants):
                    assert_same(gpu_batch_with(b, kind, corpus, **kw), exp, (metric, kind, kw, "jaro32=%d" % opt, qlen))
        for b in variants:
            b.close()
    corpus.close()


@pytest.mark.parametrize("qlen", [0, 1, 20, 32, 33, 64])
def test_integer_epilogue_table_equals_per_pair_epilogue_equals_oracle(qlen):
    """Resident corpora of at least 65 536 candidates, none longer than 255 elements: the score algebra of the integer
    metrics (details/distance.rs:154-275) is looked up in a per-launch table over (candidate length, raw result) that
    finish_int / finish_norm fill themselves.  Every metric x kind, cutoff ladders, uniform and indel-class weights, the
    ratio quirk: table == per pair (option epilogue_table = 0) == oracle."""
    L = _ffi.lib()
    rng = np.random.default_rng(1200 + qlen)
    q = (rng.integers(0, 4, qlen) + 97).astype(np.uint8)
    n = 66_000
    lens = rng.choice([0, 1, 2, 8, 20, 31, 32, 33, 63, 64, 65, 100, 254, 255], n, p=[.05, .05, .05, .15, .15, .1, .1, .1, .05, .05, .05, .04, .03, .03])
    chars = (rng.integers(0, 4, int(lens.sum())) + 97).astype(np.uint8)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    for i in range(0, n, 5):                       # near matches so that small cutoffs keep some candidates
        a, b = int(off[i]), int(off[i + 1])
        m = min(b - a, qlen)
        chars[a:a + m] = q[:m]
    corpus = rf.Corpus(chars, off)
    cases = []
    for m in ("levenshtein", "osa", "indel", "lcs_seq"):
        for kind in ("distance", "similarity"):
            for c in (None, 0, 3, 30, 300):
                if m == "levenshtein" and kind == "similarity" and c is not None:
                    continue
                cases.append((m, kind, {} if c is None else {"cutoff": c}))
        for kind in ("normalized_distance", "normalized_similarity"):
            for c in (None, 0.0, 0.25, 0.8, 1.0):
                cases.append((m, kind, {} if c is None else {"cutoff": c}))
    cases += [("ratio", "similarity", {}), ("ratio", "similarity", {"cutoff": 0.6}), ("ratio", "similarity", {"reference_quirks": True}),
              ("levenshtein", "distance", {"weights": (2, 2, 2), "cutoff": 40}), ("levenshtein", "normalized_similarity", {"weights": (1, 1, 2)}),
              ("levenshtein", "similarity", {"weights": (3, 3, 3)})]
    for metric, kind, kw in cases:
        exp = orc.batch(metric, kind, q, chars, off, nthreads=0, **kw)
        for opt in (1, 0):
            b = _bc(metric, q)
            _ffi.check(L.rf_batch_set_option(b._h, b"epilogue_table", opt))
            a = Args()
            if "cutoff" in kw:
                a = a.score_cutoff(kw["cutoff"])
            if "weights" in kw:
                a = a.weights(*kw["weights"])
            a = a.reference_quirks(kw.get("reference_quirks", False))
            r = b._score(kind, corpus, a)
            if isinstance(r, np.ma.MaskedArray):
                r = r.filled(np.nan if r.dtype == np.float64 else _ffi.NONE_U32)
            assert_same(r, exp, (metric, kind, kw, "epilogue_table=%d" % opt, qlen))
            b.close()
    corpus.close()


def test_byte_results_on_a_resident_corpus():
    """rf_batch_score_u8 / _device: the u32 scores narrowed on the device (None -> 0xFF), == rf_batch_score_u32 == oracle;
    a score above 254 is refused loudly after the output is filled; float-valued kinds are refused."""
    import torch
    L = _ffi.lib()
    q = synth.synth_query(5, 32)
    chars, offsets = synth.synth_corpus(5, q, 40_003, 0, 64, 16)
    corpus = rf.Corpus(chars, offsets)
    for metric, kind, cut in (("levenshtein", "distance", None), ("levenshtein", "distance", 9), ("indel", "similarity", 30),
                              ("osa", "distance", None), ("lcs_seq", "similarity", None), ("hamming", "distance", None)):
        b = _bc(metric, q)
        a = Args() if cut is None else Args().score_cutoff(cut)
        if metric == "hamming":
            a = a.pad(True)
        kw = {} if cut is None else {"cutoff": cut}
        if metric == "hamming":
            kw["pad"] = True
        exp = orc.batch(metric, kind, q, chars, offsets, nthreads=0, **kw)
        exp8 = np.where(exp == 0xFFFFFFFF, 255, exp).astype(np.uint8)
        assert exp[exp != 0xFFFFFFFF].max() <= 254
        got = b.score_u8(kind, corpus, a)
        assert np.array_equal(got, exp8), (metric, kind, cut)
        for shift in (0, 1):       # a byte-aligned caller buffer too; the bytes around the result stay untouched
            buf = torch.full((len(corpus) + 16,), 0xAB, dtype=torch.uint8, device="cuda")
            out = buf[4 + shift: 4 + shift + len(corpus)]
            ca = a._c(False)
            _ffi.check(L.rf_batch_score_u8_device(b._h, corpus._h, _ffi.KINDS[kind], C.byref(ca), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            h = buf.cpu().numpy()
            assert np.array_equal(h[4 + shift: 4 + shift + len(corpus)], exp8), (metric, kind, cut, "device", shift)
            assert np.all(h[: 4 + shift] == 0xAB) and np.all(h[4 + shift + len(corpus):] == 0xAB), "wrote outside the result"
        b.close()
    b = _bc("levenshtein", q)
    with pytest.raises(rf.RfError):
        b.score_u8("normalized_distance", corpus)
    b.close()
    corpus.close()
    # one candidate of 300 elements: its distance does not fit a byte
    lens = np.array([5, 300, 7])
    ch = np.full(int(lens.sum()), 120, np.uint8)
    off = np.zeros(4, np.uint64); off[1:] = np.cumsum(lens)
    c2 = rf.Corpus(ch, off)
    b = _bc("levenshtein", q)
    with pytest.raises(rf.RfError, match="254"):
        b.score_u8("distance", c2)
    b.close(); c2.close()


def test_corpus_without_its_csr_copy():
    """rf_corpus_release_csr frees the CSR copy; everything the interleaved layout serves gives the same results as before
    (incl. a small cutoff on a 200-element query, which then takes the register-column kernel instead of the band kernel),
    the rest is refused loudly -- never a wrong answer or a fault."""
    rng = np.random.default_rng(77)
    q = (rng.integers(0, 5, 32) + 97).astype(np.uint8)
    chars, off = make_corpus(rng, 9000, [0, 1, 8, 20, 33, 64, 70, 200, 260], alphabet=5, query=q)
    corpus = rf.Corpus(chars, off)
    assert corpus.has_csr
    corpus.release_csr()
    assert not corpus.has_csr
    q200 = (rng.integers(0, 5, 200) + 97).astype(np.uint8)
    q64 = (rng.integers(0, 5, 64) + 97).astype(np.uint8)
    for metric, kind, qq, kw in (("levenshtein", "distance", q, {}), ("levenshtein", "normalized_similarity", q64, {"cutoff": 0.3}),
                                 ("osa", "distance", q64, {}), ("indel", "similarity", q, {}), ("ratio", "similarity", q, {}),
                                 ("jaro_winkler", "similarity", q, {}), ("jaro", "distance", q64, {"cutoff": 0.6}),
                                 ("hamming", "distance", q, {"pad": True}), ("damerau_levenshtein", "distance", q, {}),
                                 ("levenshtein", "distance", q, {"weights": (1, 2, 3)}),
                                 ("levenshtein", "distance", q200, {}), ("levenshtein", "distance", q200, {"cutoff": 20}),
                                 ("indel", "distance", q200, {"cutoff": 150})):
        b = _bc(metric, qq)
        a = Args()
        if "cutoff" in kw:
            a = a.score_cutoff(kw["cutoff"])
        if "weights" in kw:
            a = a.weights(*kw["weights"])
        if "pad" in kw:
            a = a.pad(True)
        r = b._score(kind, corpus, a)
        if isinstance(r, np.ma.MaskedArray):
            r = r.filled(np.nan if r.dtype == np.float64 else _ffi.NONE_U32)
        assert_same(r, orc.batch(metric, kind, qq, chars, off, nthreads=0, **kw), (metric, kind, len(qq), kw))
        b.close()
    b = _bc("levenshtein", q)
    ti, ts = b.extract("distance", corpus, k=7)
    exp = orc.batch("levenshtein", "distance", q, chars, off, nthreads=0)
    order = np.lexsort((np.arange(len(exp)), exp))[:7]
    assert np.array_equal(ti, order.astype(np.uint32)) and np.array_equal(ts, exp[order])
    _ffi.check(_ffi.lib().rf_batch_set_option(b._h, b"single_word_path", 1))   # the CSR kernel is gone: the knob is ignored
    assert np.array_equal(b.distance(corpus), exp)
    b.close()
    idx, dist = rf.cdist_topk([q, q64[:40]], corpus, k=5)
    assert np.array_equal(idx[0], order[:5].astype(idx.dtype))
    qlong = (rng.integers(0, 5, 600) + 97).astype(np.uint8)
    for metric, qq in (("levenshtein", qlong), ("jaro", q200), ("prefix", q), ("postfix", q)):
        b = _bc(metric, qq)
        with pytest.raises(rf.RfError, match="release_csr"):
            b._score("similarity" if metric in ("prefix", "postfix", "jaro") else "distance", corpus, None)
        b.close()
    # u32 query on the byte corpus: the table-driven metrics map the QUERY into the byte domain (no pass over the candidates)
    # and keep working; the metrics that compare symbols directly would have to rename the candidates -> refused
    qw = q.astype(np.uint32)
    qw[::3] += 1000
    full = rf.Corpus(chars, off)
    b = _bc("levenshtein", qw)
    assert np.array_equal(b._score("distance", corpus, None), b._score("distance", full, None))
    b.close()
    b = _bc("hamming", qw)
    with pytest.raises(rf.RfError, match="release_csr"):
        b._score("distance", corpus, Args().pad(True))
    b.close()
    full.close()
    corpus.close()


@pytest.mark.parametrize("qlen", [65, 2048, 2049, 5000, 70_000])
def test_dp_metrics_with_long_queries(qlen):
    """Generic Levenshtein weights (weighted_wagner_fischer, levenshtein.rs:212-259) and Damerau-Levenshtein
    (damerau_levenshtein.rs:111-214) have no query-length limit in the reference.  Thread-per-candidate kernels with the DP
    rows in global scratch: the grid shrinks with the query so the rows stay within 2 GB; queries beyond 48 K elements need
    the opt-in shared-memory size for their copy of the query.  (Round 1 refused everything beyond 2048.)"""
    rng = np.random.default_rng(qlen)
    q = (rng.integers(0, 4, qlen) + 97).astype(np.uint8)
    n = 20 if qlen > 10_000 else 300
    lens = rng.choice([0, 1, 17, 64] if qlen > 10_000 else [0, 1, 17, 64, 200], n)   # (one thread walks len1 x len2 cells)
    chars = (rng.integers(0, 4, int(lens.sum())) + 97).astype(np.uint8)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    corpus = rf.Corpus(chars, off)
    for metric, kw in (("levenshtein", {"weights": (1, 2, 3)}), ("levenshtein", {"weights": (3, 1, 2), "cutoff": qlen}),
                       ("damerau_levenshtein", {}), ("damerau_levenshtein", {"cutoff": qlen - 10})):
        b = _bc(metric, q)
        a = Args()
        if "weights" in kw:
            a = a.weights(*kw["weights"])
        if "cutoff" in kw:
            a = a.score_cutoff(kw["cutoff"])
        r = b._score("distance", corpus, a)
        if isinstance(r, np.ma.MaskedArray):
            r = r.filled(_ffi.NONE_U32)
        assert_same(r, orc.batch(metric, "distance", q, chars, off, nthreads=0, **kw), (metric, kw, qlen))
        b.close()
    corpus.close()


def test_options_are_per_comparator():
    """The kernel-choice knobs are copied into a comparator at creation (rf_batch_set_option changes one comparator):
    two comparators with different settings give the same results side by side, and flipping the process-wide default
    afterwards does not touch them."""
    L = _ffi.lib()
    q = synth.synth_query(5, 32)
    q3 = synth.synth_query(6, 256)
    chars, offsets = synth.synth_corpus(5, q, 50_000, 8, 64, 16)
    c3, o3 = synth.synth_corpus(6, q3, 20_000, 64, 256, 48)
    corpus, corpus3 = rf.Corpus(chars, offsets), rf.Corpus(c3, o3)
    exp = orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=0)
    expj = orc.batch("jaro", "similarity", q, chars, offsets, nthreads=0)
    exp3 = orc.batch("levenshtein", "distance", q3, c3, o3, nthreads=0, cutoff=40)
    a, b, c = _bc("levenshtein", q), _bc("levenshtein", q), _bc("levenshtein", q)
    _ffi.check(L.rf_batch_set_option(b._h, b"single_word_path", 1))
    _ffi.check(L.rf_batch_set_option(c._h, b"single_word_path", 2))
    j0, j1 = _bc("jaro", q), _bc("jaro", q)
    _ffi.check(L.rf_batch_set_option(j1._h, b"jaro32", 0))
    m0, m1, m2 = _bc("levenshtein", q3), _bc("levenshtein", q3), _bc("levenshtein", q3)
    _ffi.check(L.rf_batch_set_option(m1._h, b"multi_word_path", 1))
    _ffi.check(L.rf_batch_set_option(m2._h, b"banded_levenshtein", 0))
    _ffi.check(L.rf_set_option(b"single_word_path", 1))     # a default flipped later must not reach existing comparators
    try:
        n0 = L.rf_kernel_launch_count()
        for x in (a, b, c):
            assert np.array_equal(x.distance(corpus), exp)
        for x in (j0, j1):
            assert np.array_equal(x.similarity(corpus), expj)
        for x in (m0, m1, m2):
            assert np.array_equal(x.distance_with_args(corpus3, Args().score_cutoff(40)).filled(0xFFFFFFFF), exp3)
        assert L.rf_kernel_launch_count() > n0
        with pytest.raises(rf.RfError):
            _ffi.check(L.rf_batch_set_option(a._h, b"stream_chunk_mb", 1))
    finally:
        _ffi.check(L.rf_set_option(b"single_word_path", 0))
    for x in (a, b, c, j0, j1, m0, m1, m2):
        x.close()
    corpus.close()
    corpus3.close()


# ------------------------------------------------------------------------------------------------ element types
def test_typed_element_entry_points_compare_by_value():
    """VERDICT r1 missing #5: native u16 / i8 ... i64 / u64 C entries (HashableChar, details/common.rs:29-37).  The library
    widens BY VALUE; i16 -1 never equals u16 65535; the one collision of the 32-bit domain (negative signed vs unsigned
    >= 2^31) is refused loudly, never answered wrongly; 64-bit values that do not fit are refused at creation."""
    rng = np.random.default_rng(8)
    vals = np.array([-300, -1, 0, 5, 97, 255, 256, 30000], dtype=np.int64)   # all fit i16: no wrap-around in the casts below
    ids = {int(v): i + 1 for i, v in enumerate(vals)}
    ren = lambda a: np.array([ids[int(x)] for x in a], dtype=np.uint32)
    lens = rng.integers(0, 40, 600)
    cand = vals[rng.integers(0, len(vals), int(lens.sum()))]
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    q = vals[rng.integers(0, len(vals), 24)]
    for cdt in (np.int16, np.int32, np.int64):
        corpus = rf.Corpus.from_typed(cand.astype(cdt), off)
        for qdt in (np.int16, np.int32, np.int64):
            for m, kind in (("levenshtein", "distance"), ("jaro_winkler", "similarity"), ("lcs_seq", "similarity")):
                B = type("B", (BatchComparatorBase,), {"METRIC": m})
                b = B.from_typed(q.astype(qdt))
                exp = orc.batch(m, kind, ren(q), ren(cand), off, nthreads=0)
                is_f = exp.dtype == np.float64
                out = np.empty(len(lens), exp.dtype)
                fn = _ffi.lib().rf_batch_score_f64 if is_f else _ffi.lib().rf_batch_score_u32
                _ffi.check(fn(b._h, corpus._h, _ffi.KINDS[kind], None, out.ctypes.data))
                assert_same(out, exp, (cdt, qdt, m))
                b.close()
        corpus.close()
    L = _ffi.lib()

    def dist(qarr, carr):
        c = rf.Corpus.from_typed(carr, np.array([0, len(carr)], np.uint64))
        b = type("B", (BatchComparatorBase,), {"METRIC": "levenshtein"}).from_typed(qarr)
        out = np.zeros(1, np.uint32)
        try:
            _ffi.check(L.rf_batch_score_u32(b._h, c._h, 0, None, out.ctypes.data))
        finally:
            b.close()
            c.close()
        return int(out[0])
    assert dist(np.array([-1, 7, -1], np.int16), np.array([65535, 7, 65535], np.uint16)) == 2      # same 16 bits, different values
    assert dist(np.array([65535, 7], np.uint16), np.array([65535, 7], np.uint64)) == 0
    assert dist(np.array([-1, 7], np.int8), np.array([255, 7], np.uint8)) == 1
    assert dist(np.array([-1, 7], np.int8), np.array([-1, 7], np.int64)) == 0
    assert dist(np.array([200, 7], np.uint8), np.array([200, 7], np.int16)) == 0
    with pytest.raises(rf.RfError) as ei:       # negative i32 vs u32 >= 2^31: same 32-bit pattern, never equal in the reference
        dist(np.array([-1], np.int32), np.array([0xFFFFFFFF], np.uint32))
    assert ei.value.status == _ffi.RF_ERR_UNSUPPORTED
    with pytest.raises(rf.RfError) as ei:
        rf.Corpus.from_typed(np.array([1 << 40], np.uint64), np.array([0, 1], np.uint64))
    assert ei.value.status == _ffi.RF_ERR_UNSUPPORTED


def test_u32_streaming_and_cdist_entry_points():
    """VERDICT r1 missing #5: rf_batch_stream_*_elems32 and rf_cdist_topk_u32 (the u32 counterparts of the byte entry points)."""
    rng = np.random.default_rng(3)
    alphabet = np.array([97, 98, 99, 0x4E2D, 0x6587, 0x1F600, 1048, 0], dtype=np.uint32)
    q = alphabet[rng.integers(0, 6, 30)]
    cands = []
    for _ in range(9000):
        if rng.random() < 0.4:
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
            cands.append(alphabet[rng.integers(0, len(alphabet), int(rng.choice([0, 1, 5, 33, 64, 70, 300])))])
    elems = np.concatenate(cands).astype(np.uint32)
    offsets = np.zeros(len(cands) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(c) for c in cands])
    L = _ffi.lib()
    for mb, kc in ((64, 2048), (1, 1)):
        _ffi.check(L.rf_set_option(b"stream_chunk_mb", mb))
        _ffi.check(L.rf_set_option(b"stream_chunk_kcand", kc))
        try:
            for m, kind, kw in (("levenshtein", "distance", {}), ("levenshtein", "distance", {"cutoff": 5}),
                                ("jaro_winkler", "similarity", {}), ("indel", "normalized_similarity", {})):
                b = _bc(m, q)
                a = Args().score_cutoff(kw["cutoff"]) if kw else Args()
                assert_same(b.stream_elems32(kind, elems, offsets, a), orc.batch(m, kind, q, elems, offsets, nthreads=0, **kw), ("elems32", m, kind, mb))
                b.close()
        finally:
            _ffi.check(L.rf_set_option(b"stream_chunk_mb", 64))
            _ffi.check(L.rf_set_option(b"stream_chunk_kcand", 2048))
    # cdist on the (compacted) u32 corpus
    corpus = rf.Corpus.from_u32(elems, offsets)
    qs = [q, alphabet[rng.integers(0, 8, 12)], np.array([0x4E2D, 5555, 97], np.uint32), np.zeros(0, np.uint32)]
    q_off = np.zeros(len(qs) + 1, np.uint64)
    q_off[1:] = np.cumsum([len(x) for x in qs])
    gi, gd = rf.cdist_topk_u32(np.concatenate(qs), q_off, corpus, k=6)
    for qi, qq in enumerate(qs):
        d = orc.batch("levenshtein", "distance", qq, elems, offsets, nthreads=0).astype(np.int64)
        keys = np.sort(d * (1 << 32) + np.arange(len(d)))[:6]
        assert np.array_equal(gi[qi], (keys & 0xFFFFFFFF).astype(np.uint32)) and np.array_equal(gd[qi], (keys >> 32).astype(np.uint32)), qi
    corpus.close()
    _ffi.check(L.rf_set_option(b"compact_u32_corpus", 0))
    try:
        c2 = rf.Corpus.from_u32(elems, offsets)
        with pytest.raises(rf.RfError) as ei:
            rf.cdist_topk_u32(np.concatenate(qs), q_off, c2, k=6)
        assert ei.value.status == _ffi.RF_ERR_UNSUPPORTED
        c2.close()
    finally:
        _ffi.check(L.rf_set_option(b"compact_u32_corpus", 1))


# ------------------------------------------------------------------------------------------------ rf_comm (process-per-GPU form)
def _comm_rank_worker(rank, nranks, uid, q, chars, offsets, out, errors, chunks):
    try:
        import torch
        L = _ffi.lib()
        n = len(offsets) - 1
        lo, hi = (n * rank) // nranks // 3 * 3, (n * (rank + 1)) // nranks // 3 * 3 if rank + 1 < nranks else n   # unequal, unaligned shards
        if rank == 0:
            lo = 0
        sub_off = (offsets[lo:hi + 1] - offsets[lo]).astype(np.uint64)
        sub_chars = chars[int(offsets[lo]): int(offsets[hi])]
        corpus = rf.Corpus(sub_chars, sub_off, device=rank)
        comm = C.c_void_p()
        _ffi.check(L.rf_comm_create_rank(uid, nranks, rank, rank, C.byref(comm)))
        res = {}
        for metric, kind, is_f in (("levenshtein", "distance", False), ("jaro_winkler", "similarity", True), ("hamming", "distance", False)):
            b = _bc(metric, q, device=rank)
            a = Args().pad(True)._c(is_f)
            full = torch.full((n,), -1, dtype=torch.float64 if is_f else torch.int32, device="cuda:%d" % rank)
            counts = (C.c_uint64 * nranks)()
            fn = L.rf_batch_score_f64_allgather_device if is_f else L.rf_batch_score_u32_allgather_device
            with torch.cuda.device(rank):
                st = torch.cuda.current_stream().cuda_stream
                for _ in range(2):   # the second call reuses the cached counts
                    _ffi.check(fn(b._h, corpus._h, comm, _ffi.KINDS[kind], C.byref(a), full.data_ptr(), n, counts, st))
                torch.cuda.synchronize()
            res[metric] = full.cpu().numpy()
            assert sum(counts) == n and counts[rank] == hi - lo
            b.close()
        out[rank] = res
        L.rf_comm_destroy(comm)
        corpus.close()
    except BaseException as e:   # noqa: BLE001
        errors.append((rank, repr(e)))


@pytest.mark.parametrize("chunks", [4, 1, 16])
def test_comm_allgather_overlapped_equals_oracle(chunks):
    """rf_comm_* + rf_batch_score_*_allgather_device: every rank ends with ALL ranks' scores in candidate order; the shard
    is scanned in pieces whose transfer overlaps the next piece's scan.  One rank per visible GPU (threads of this
    process stand in for the host's processes); a single GPU runs the 1-rank form."""
    L = _ffi.lib()
    nranks = min(L.rf_device_count(), 8)
    q = synth.synth_query(2, 32)
    chars, offsets = synth.synth_corpus(2, q, 700_003, 8, 64, 16)
    uid = C.create_string_buffer(128)
    _ffi.check(L.rf_comm_unique_id(uid))
    _ffi.check(L.rf_set_option(b"allgather_chunks", chunks))
    out, errors = {}, []
    try:
        ths = [threading.Thread(target=_comm_rank_worker, args=(r, nranks, uid, q, chars, offsets, out, errors, chunks)) for r in range(nranks)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    finally:
        _ffi.check(L.rf_set_option(b"allgather_chunks", 4))
    assert not errors, errors
    exp = {"levenshtein": orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=0),
           "jaro_winkler": orc.batch("jaro_winkler", "similarity", q, chars, offsets, nthreads=0),
           "hamming": orc.batch("hamming", "distance", q, chars, offsets, nthreads=0, pad=True)}
    for r in range(nranks):
        assert np.array_equal(out[r]["levenshtein"].view(np.uint32), exp["levenshtein"]), r
        assert np.array_equal(out[r]["jaro_winkler"], exp["jaro_winkler"]), r
        assert np.array_equal(out[r]["hamming"].view(np.uint32), exp["hamming"]), r


@pytest.mark.parametrize("chunk_mb,chunk_kcand", [(64, 2048), (1, 2048), (1, 4)])
def test_streaming_6bit_packed_characters(chunk_mb, chunk_kcand):
    """rf_pack6_u8 + rf_batch_stream_*_len8_packed6: the candidates cross PCIe 6-bit packed (62-symbol alphabet) and are
    unpacked per chunk on the device; identical to the oracle for every chunking, byte results included."""
    L = _ffi.lib()
    _ffi.check(L.rf_set_option(b"stream_chunk_mb", chunk_mb))
    _ffi.check(L.rf_set_option(b"stream_chunk_kcand", chunk_kcand))
    try:
        q = synth.synth_query(2, 32)
        chars, offsets = synth.synth_corpus(2, q, 300_007, 0, 64, 16)
        lens = np.diff(offsets.astype(np.int64)).astype(np.uint8)
        packed, d64 = rf.pack6(chars)
        for m, kind, kw in (("levenshtein", "distance", {}), ("levenshtein", "distance", {"cutoff": 9}), ("indel", "distance", {}),
                            ("osa", "distance", {})):
            b = _bc(m, q)
            a = Args().score_cutoff(kw["cutoff"]) if kw else Args()
            exp = orc.batch(m, kind, q, chars, offsets, nthreads=0, **kw)
            assert_same(b.stream_len8_packed6(kind, packed, d64, lens, a), exp, ("packed6", m, kw, chunk_mb))
            got8 = b.stream_len8_packed6(kind, packed, d64, lens, a, u8_results=True)
            assert np.array_equal(got8, np.where(exp == 0xFFFFFFFF, 255, exp).astype(np.uint8)), ("packed6 u8", m, kw)
            b.close()
        # a query symbol outside the corpus' dictionary simply never matches
        q2 = q.copy()
        q2[::3] = ord("#")
        b = _bc("levenshtein", q2)
        assert_same(b.stream_len8_packed6("distance", packed, d64, lens), orc.batch("levenshtein", "distance", q2, chars, offsets, nthreads=0), "packed6 foreign symbol")
        b.close()
    finally:
        _ffi.check(L.rf_set_option(b"stream_chunk_mb", 64))
        _ffi.check(L.rf_set_option(b"stream_chunk_kcand", 2048))


@pytest.mark.parametrize("qlen,distinct", [(300, 300), (700, 400), (5000, 3000), (20000, 9000)])
def test_u32_queries_with_more_than_255_distinct_symbols(qlen, distinct):
    """VERDICT r1 missing #5: the reference's per-block hashmap takes any alphabet (pattern_match_vector.rs:20-65, :226-280).
    A u32 query with hundreds of distinct symbols is mapped into the candidates' symbol domain (byte corpora, byte
    streaming, u32 corpora renamed to bytes at creation): exact against the oracle's u32 path, every table-driven metric."""
    rng = np.random.default_rng(qlen)
    base = np.concatenate([np.arange(97, 123), np.arange(0x4E00, 0x4E00 + distinct)]).astype(np.uint32)
    q = base[rng.integers(0, len(base), qlen)]
    q[: distinct] = base[26: 26 + distinct][: min(distinct, qlen)]     # make sure the query really holds `distinct` symbols
    assert len(np.unique(q)) > 255
    # (a) byte candidates: only the query's letters can ever match
    chars, offsets = make_corpus(rng, 300, [0, 1, 64, 65, 300, qlen // 2, qlen], alphabet=26, base=97,
                                 query=np.where(q < 256, q, 35).astype(np.uint8), near_frac=0.5)
    corpus8 = rf.Corpus(chars, offsets)
    # (b) u32 candidates over a 200-symbol alphabet (renamed to bytes at creation): half of the query's symbols occur in it
    alpha_c = np.concatenate([np.arange(97, 123), np.arange(0x4E00, 0x4E00 + 170)]).astype(np.uint32)
    cands = []
    for _ in range(200):
        if rng.random() < 0.5:
            c = q.copy()
            hit = (rng.random(qlen) < 0.1) | ~np.isin(c, alpha_c)      # the corpus itself stays within its 196 symbols
            c[hit] = alpha_c[rng.integers(0, len(alpha_c), int(hit.sum()))]
            cands.append(c[: int(rng.integers(qlen // 2, qlen + 1))])
        else:
            cands.append(alpha_c[rng.integers(0, len(alpha_c), int(rng.choice([0, 5, 100, qlen])))])
    elems = np.concatenate(cands).astype(np.uint32)
    off32 = np.zeros(len(cands) + 1, np.uint64)
    off32[1:] = np.cumsum([len(c) for c in cands])
    corpus32 = rf.Corpus.from_u32(elems, off32)
    metrics = [("levenshtein", "distance", {}), ("levenshtein", "distance", {"cutoff": 40}), ("indel", "normalized_similarity", {}),
               ("lcs_seq", "similarity", {}), ("osa", "distance", {})]
    if qlen <= 5000:
        metrics += [("jaro_winkler", "similarity", {}), ("jaro", "distance", {"cutoff": 0.5})]
    for m, kind, kw in metrics:
        assert_same(gpu_batch(m, kind, q, corpus8, **kw), orc.batch(m, kind, q, chars.astype(np.uint32), offsets, nthreads=0, **kw), ("wide q / u8 corpus", m, kw))
        assert_same(gpu_batch(m, kind, q, corpus32, **kw), orc.batch(m, kind, q, elems, off32, nthreads=0, **kw), ("wide q / compact u32 corpus", m, kw))
    # (c) u32 candidates with a LARGE alphabet of their own (near-copies of the query keep its symbols): 16-bit codes
    cands = []
    for _ in range(120):
        if rng.random() < 0.6:
            c = q.copy()
            hit = rng.random(qlen) < 0.05
            c[hit] = base[rng.integers(0, len(base), int(hit.sum()))] + (rng.random(int(hit.sum())) < 0.3) * 50000
            lo_ = int(rng.integers(0, qlen // 4 + 1))
            cands.append(c[lo_: int(rng.integers(qlen // 2, qlen + 1))])
        else:
            cands.append((base[rng.integers(0, len(base), int(rng.choice([0, 1, 70, qlen])))] + 7).astype(np.uint32))
    elems_c = np.concatenate(cands).astype(np.uint32)
    off_c = np.zeros(len(cands) + 1, np.uint64)
    off_c[1:] = np.cumsum([len(c) for c in cands])
    corpus_big = rf.Corpus.from_u32(elems_c, off_c)
    assert len(np.unique(elems_c)) > 255
    for m, kind, kw in metrics:
        assert_same(gpu_batch(m, kind, q, corpus_big, **kw), orc.batch(m, kind, q, elems_c, off_c, nthreads=0, **kw), ("wide q / big-alphabet u32 corpus", m, kw))
    corpus_big.close()
    b = _bc("levenshtein", q)
    assert_same(b.stream("distance", chars, offsets), orc.batch("levenshtein", "distance", q, chars.astype(np.uint32), offsets, nthreads=0), "wide q stream")
    with pytest.raises(rf.RfError) as ei:      # symbol-comparing metrics still need a byte alphabet of the query's own
        gpu_batch("damerau_levenshtein", "distance", q[:300], corpus8)
    assert ei.value.status == _ffi.RF_ERR_UNSUPPORTED
    b.close()
    corpus8.close()
    corpus32.close()


@pytest.mark.parametrize("metric", ["osa", "indel", "lcs_seq", "levenshtein"])
def test_cdist_topk_other_metrics_vs_oracle(metric):
    """VERDICT r1 missing #7: many-vs-many top-k was Levenshtein only.  rf_cdist_topk_metric_u8 for OSA / Indel / LCSseq
    distance against the oracle's full distance matrix (queries of 0 ... 64 elements, with and without a cutoff, several
    corpus slices)."""
    rng = np.random.default_rng(12)
    base = synth.synth_query(31, 40)
    chars, offsets = synth.synth_corpus(31, base, 40_000, 0, 64, 12)
    corpus = rf.Corpus(chars, offsets)
    qs = [base, base[:7], synth.synth_query(32, 64), synth.synth_query(33, 33), np.zeros(0, np.uint8), synth.synth_query(34, 1)]
    n = len(offsets) - 1
    L = _ffi.lib()
    for slices in (0, 5):
        _ffi.check(L.rf_set_option(b"cdist_slices", slices))
        try:
            for k, cut in ((10, None), (40, None), (6, 20)):
                gi, gd = rf.cdist_topk([bytes(x) for x in qs], corpus, k=k, score_cutoff=cut, metric=metric)
                for qi, qq in enumerate(qs):
                    d = orc.batch(metric, "distance", qq, chars, offsets, nthreads=0).astype(np.int64)
                    keys = np.sort(d * (1 << 32) + np.arange(n))
                    if cut is not None:
                        keys = keys[(keys >> 32) <= cut]
                    keys = keys[:k]
                    m = len(keys)
                    assert np.array_equal(gi[qi][:m], (keys & 0xFFFFFFFF).astype(np.uint32)), (metric, qi, k, cut, slices)
                    assert np.array_equal(gd[qi][:m], (keys >> 32).astype(np.uint32)), (metric, qi, k, cut, slices)
                    assert np.all(gi[qi][m:] == 0xFFFFFFFF)
        finally:
            _ffi.check(L.rf_set_option(b"cdist_slices", 0))
    corpus.close()
