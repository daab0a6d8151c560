"""The CPU oracle against every known-answer vector the reference's own tests hold for this path
(tests/golden/golden.json, made by tests/golden/make_golden.py from /root/reference), plus the
4-way-consistency idea of the reference's test helpers (levenshtein.rs:1847-1875) and a textbook-DP
cross-check on random inputs."""
import json
import math
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")
G = json.load(open(os.path.join(GOLD, "golden.json")))


def _strs(rec):
    def one(k):
        return rec[k] if k in rec else "".join(chr(c) for c in rec[k + "_cp"])
    return one("s1"), one("s2")


def _kw(rec):
    a = rec["args"]
    kw = {}
    if "cutoff" in a:
        kw["cutoff"] = a["cutoff"]
    if "weights" in a:
        kw["weights"] = tuple(a["weights"])
    if "pad" in a:
        kw["pad"] = a["pad"]
    return kw


def _check(got, exp, tol, ctx):
    if exp is None:
        assert got is None, ctx
    else:
        assert got is not None, ctx
        if tol == 0:
            assert got == exp, ctx
        else:
            assert abs(got - exp) <= tol, ctx


@pytest.mark.parametrize("idx", range(len(G["cases"])))
def test_golden_case(idx):
    rec = G["cases"][idx]
    s1, s2 = _strs(rec)
    kw = _kw(rec)
    for a, b in ((s1, s2), (s2, s1)):   # BatchComparator::new(s1)(s2) and ::new(s2)(s1)
        if rec["expected"] == "error":       # hamming::Error::DifferentLengthArgs
            with pytest.raises(orc.DifferentLengthArgs):
                orc.pair(rec["metric"], rec["kind"], a, b, **kw)
            continue
        got = orc.pair(rec["metric"], rec["kind"], a, b, **kw)
        _check(got, rec["expected"], rec["tol"], (rec, a, b, got))
    if rec["expected"] == "error":
        return
    if all(ord(c) < 128 for c in s1 + s2):  # .chars() vs .bytes() (levenshtein.rs:1877-1890)
        got32 = orc.pair(rec["metric"], rec["kind"], s1, s2, dtype=np.uint32, **kw)
        _check(got32, rec["expected"], rec["tol"], (rec, "u32", got32))


@pytest.mark.parametrize("metric", ["jaro", "jaro_winkler"])
def test_golden_matrix(metric):
    m = G["matrices"][metric]
    names, n = m["names"], len(m["names"])
    for c in m["cutoffs"]:
        for i, n1 in enumerate(names):
            for j, n2 in enumerate(names):
                sc = m["scores"][i * n + j]
                exp = sc if c <= sc else None
                got = orc.pair(metric, "similarity", n1, n2, cutoff=c)
                _check(got, exp, m["tol"], (metric, n1, n2, c, got))
                gotd = orc.pair(metric, "distance", n1, n2, cutoff=1.0 - c)
                _check(gotd, None if exp is None else 1.0 - exp, m["tol"], (metric, "dist", n1, n2, c, gotd))


def test_ocr_large_band():
    # levenshtein.rs:2139-2161: 106514 x 107244 elements -> 5278; None at cutoff 2500; 5278 with score_hint(0)
    z = np.load(os.path.join(GOLD, "ocr.npz"))
    a, b = z["OCR_EXAMPLE1"], z["OCR_EXAMPLE2"]
    assert len(a) == 106514 and len(b) == 107244
    assert orc.pair("levenshtein", "distance", a, b) == 5278
    assert orc.pair("levenshtein", "distance", a, b, cutoff=2500) is None
    assert orc.pair("levenshtein", "distance", a, b, hint=0) == 5278


def _rand_pairs(rng, n, lens, alphabet):
    for _ in range(n):
        l1, l2 = rng.choice(lens), rng.choice(lens)
        a = rng.integers(0, alphabet, l1).astype(np.uint8) + 97
        if rng.random() < 0.5 and l1 > 0:   # related pair: mutate a
            b = list(a)
            for _ in range(rng.integers(0, 6)):
                op = rng.integers(0, 3)
                pos = rng.integers(0, len(b) + 1)
                if op == 0 and b:
                    b[min(pos, len(b) - 1)] = 97 + rng.integers(0, alphabet)
                elif op == 1:
                    b.insert(pos, 97 + rng.integers(0, alphabet))
                elif b:
                    del b[min(pos, len(b) - 1)]
            b = np.array(b, dtype=np.uint8)
        else:
            b = rng.integers(0, alphabet, l2).astype(np.uint8) + 97
        yield a, b


LENS = [0, 1, 2, 3, 5, 8, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 200, 255, 256, 257, 300]


def test_oracle_vs_textbook_integer_metrics():
    rng = np.random.default_rng(7)
    for a, b in _rand_pairs(rng, 1500, LENS, 4):
        d = orc.tb("levenshtein", a, b)
        l = orc.tb("lcs", a, b)
        o = orc.tb("osa", a, b)
        assert orc.pair("levenshtein", "distance", a, b) == d
        assert orc.pair("lcs_seq", "similarity", a, b) == l
        assert orc.pair("indel", "distance", a, b) == len(a) + len(b) - 2 * l
        assert orc.pair("osa", "distance", a, b) == o
        # cutoff semantics: Some(d) iff d <= cutoff, through every dispatcher branch
        for c in (0, 1, 2, 3, 4, 5, 31, 32, 33, 64, max(len(a), len(b)), 2**64 - 1):
            got = orc.pair("levenshtein", "distance", a, b, cutoff=c)
            assert got == (d if d <= c else None), (bytes(a), bytes(b), c, got, d)
            goti = orc.pair("indel", "distance", a, b, cutoff=c)
            di = len(a) + len(b) - 2 * l
            assert goti == (di if di <= c else None), (bytes(a), bytes(b), c, goti, di)
            gots = orc.pair("lcs_seq", "similarity", a, b, cutoff=c)
            assert gots == (l if l >= c else None), (bytes(a), bytes(b), c, gots, l)
        # weighted routes: (1,1,2) -> indel (levenshtein.rs:1321), (1,2,3) -> generic Wagner-Fischer (:1330)
        assert orc.pair("levenshtein", "distance", a, b, weights=(1, 1, 2)) == orc.tb("levenshtein", a, b, 1, 1, 2)
        assert orc.pair("levenshtein", "distance", a, b, weights=(1, 2, 3)) == orc.tb("levenshtein", a, b, 1, 2, 3)


def test_oracle_block_and_small_band_kernels_direct():
    """Force the block / small-band kernels (levenshtein.rs:509-617, :769-1019) on inputs the dispatcher
    would route elsewhere."""
    rng = np.random.default_rng(11)
    for a, b in _rand_pairs(rng, 800, [65, 66, 100, 128, 129, 200, 256, 300], 3):
        if len(a) == 0 or len(b) == 0:
            continue
        d = orc.tb("levenshtein", a, b)
        for c in (4, 8, 16, 31, 32, 33, 64, 100, 1000):
            if c < abs(len(a) - len(b)):
                continue
            got = orc.lev_block(a, b, c)
            assert got == (d if d <= min(c, max(len(a), len(b))) else None), (bytes(a), bytes(b), c, got, d)
            if len(a) > 64 and 2 * c + 1 <= 64:
                g2 = orc.lev_small_band(a, b, c)
                g2 = g2 if (g2 is not None and g2 <= c) else None
                assert g2 == (d if d <= c else None), (bytes(a), bytes(b), c, g2, d)


def test_oracle_vs_textbook_jaro():
    rng = np.random.default_rng(13)
    for a, b in _rand_pairs(rng, 1500, LENS, 5):
        j = orc.tb("jaro", a, b)
        jw = orc.tb("jaro_winkler", a, b, 0.1)
        assert abs(orc.pair("jaro", "similarity", a, b) - j) < 1e-12, (bytes(a), bytes(b))
        assert abs(orc.pair("jaro_winkler", "similarity", a, b) - jw) < 1e-12
        ns = orc.pair("jaro_winkler", "normalized_similarity", a, b)
        assert abs(ns - jw) < 1e-12
        for c in (0.3, 0.7, 0.75, 0.9, 1.0):
            got = orc.pair("jaro_winkler", "similarity", a, b, cutoff=c)
            assert (got is None) == (jw < c) or abs(jw - c) < 1e-9, (bytes(a), bytes(b), c, got, jw)


def test_batch_matches_pair_and_sentinels():
    rng = np.random.default_rng(17)
    cands = [b for _, b in _rand_pairs(rng, 300, [0, 1, 5, 20, 40, 64, 70], 4)]
    chars = np.concatenate([c for c in cands] + [np.zeros(0, np.uint8)])
    offsets = np.concatenate([[0], np.cumsum([len(c) for c in cands])]).astype(np.uint64)
    q = b"abcdabcdabcdabcdabcd"
    out = orc.batch("levenshtein", "distance", q, chars, offsets, cutoff=10)
    outf = orc.batch("jaro_winkler", "normalized_similarity", q, chars, offsets, cutoff=0.5)
    for i, c in enumerate(cands):
        p = orc.pair("levenshtein", "distance", q, c, cutoff=10)
        assert out[i] == (0xFFFFFFFF if p is None else p)
        pf = orc.pair("jaro_winkler", "normalized_similarity", q, c, cutoff=0.5)
        assert (math.isnan(outf[i]) and pf is None) or outf[i] == pf


def test_ratio_quirk_q1():
    # fuzz.rs:94-96 documents 0.9655 for both; the literal batch path divides by max(len1,len2)
    a, b = "this is a test", "this is a test!"
    assert abs(orc.pair("ratio", "similarity", a, b) - 28 / 29) < 1e-12
    assert abs(orc.pair("ratio", "similarity", a, b, reference_quirks=True) - 14 / 15) < 1e-12


def test_damerau_levenshtein_zhao_vs_textbook():
    """The restated Zhao/Sahni linear-space algorithm == Lowrance-Wagner full-matrix Damerau-Levenshtein; OSA >= DL."""
    rng = np.random.default_rng(123)
    for _ in range(4000):
        a = (rng.integers(0, 3, int(rng.integers(0, 14))) + 97).astype(np.uint8)
        b = (rng.integers(0, 3, int(rng.integers(0, 14))) + 97).astype(np.uint8)
        d = orc.pair("damerau_levenshtein", "distance", a, b)
        assert d == orc.tb("damerau_levenshtein", a, b), (bytes(a), bytes(b))
        assert d == orc.pair("damerau_levenshtein", "distance", b, a)
        assert orc.tb("levenshtein", a, b) >= orc.tb("osa", a, b) >= d
        for c in (0, 1, 2, 5):
            assert orc.pair("damerau_levenshtein", "distance", a, b, cutoff=c) == (d if d <= c else None)
