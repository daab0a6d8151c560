"""Helpers shared by the GPU parity tests: seeded corpora and oracle-vs-GPU comparison through the C ABI."""
import numpy as np

import rapidfuzz_b200 as rf
from rapidfuzz_b200 import _ffi
from rapidfuzz_b200._scorer import Args, BatchComparatorBase
from oracle import oracle as orc


def make_corpus(rng, n, lens, alphabet=6, base=97, query=None, near_frac=0.3, high_bytes=False):
    """n candidates with lengths drawn from `lens`; a fraction are edited copies of `query`."""
    cands = []
    for _ in range(n):
        if query is not None and len(query) and rng.random() < near_frac:
            b = list(query)
            for _ in range(int(rng.integers(0, 10))):
                op, pos = rng.integers(0, 3), int(rng.integers(0, len(b) + 1))
                if op == 0 and b:
                    b[min(pos, len(b) - 1)] = base + int(rng.integers(0, alphabet))
                elif op == 1:
                    b.insert(pos, base + int(rng.integers(0, alphabet)))
                elif b:
                    del b[min(pos, len(b) - 1)]
            c = np.array(b, dtype=np.uint8)
        else:
            l = int(rng.choice(lens))
            c = (rng.integers(0, alphabet, l) + base).astype(np.uint8)
            if high_bytes and l and rng.random() < 0.3:
                c[rng.integers(0, l)] = int(rng.integers(128, 256))
        cands.append(c)
    chars = np.concatenate(cands + [np.zeros(0, np.uint8)]).astype(np.uint8)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(c) for c in cands])
    return chars, offsets


def gpu_batch(metric, kind, query, corpus, cutoff=None, weights=None, prefix_weight=0.1, reference_quirks=False):
    """Raw sentinel-carrying result (u32 with 0xFFFFFFFF / f64 with NaN) like oracle.batch."""
    cls = type("B", (BatchComparatorBase,), {"METRIC": metric})
    b = cls(query)
    a = Args()
    if cutoff is not None:
        a = a.score_cutoff(cutoff)
    if weights is not None:
        a = a.weights(*weights)
    a = a.prefix_weight(prefix_weight).reference_quirks(reference_quirks)
    try:
        r = b._score(kind, corpus, a)
    finally:
        b.close()
    if isinstance(r, np.ma.MaskedArray):
        is_f = r.dtype == np.float64
        return r.filled(np.nan if is_f else _ffi.NONE_U32)
    return r


def assert_same(got, exp, ctx):
    assert got.dtype == exp.dtype, (ctx, got.dtype, exp.dtype)
    if got.dtype == np.float64:
        gn, en = np.isnan(got), np.isnan(exp)
        bad = np.nonzero((gn != en) | (~gn & ~en & (got != exp)))[0]
    else:
        bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, (ctx, "mismatches", len(bad), "first", int(bad[0]), got[bad[0]], exp[bad[0]])


def check(metric, kind, query, chars, offsets, corpus=None, **kw):
    own = corpus is None
    if own:
        corpus = rf.Corpus(chars, offsets)
    try:
        got = gpu_batch(metric, kind, query, corpus, **kw)
    finally:
        if own:
            corpus.close()
    exp = orc.batch(metric, kind, query, chars, offsets, nthreads=0, **kw)
    assert_same(got, exp, (metric, kind, bytes(query)[:40], kw))
