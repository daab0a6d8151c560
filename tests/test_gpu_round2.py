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
