"""CPU check of the exact per-candidate arithmetic the CUDA kernels execute (csrc/rf_core.cuh compiled for
the host by tests/core_host_shim.cpp) against the oracle: bit tricks (top-aligned Myers, popcount score,
funnel-shift byte reader), Jaro flag/transposition passes and the score algebra incl. cutoffs."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "core_host.so")


@pytest.fixture(scope="module")
def core():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(HERE, "core_host_shim.cpp")
    hdr = os.path.join(HERE, "..", "rapidfuzz-rs_b200", "csrc", "rf_core.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-x", "c++", "-o", SO, src])
    lib = C.CDLL(SO)
    lib.core_raw.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
    lib.core_raw.restype = C.c_uint32
    lib.core_score.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_uint64,
                               C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_int, C.c_uint32,
                               C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
    lib.core_score.restype = C.c_int
    lib.core_lev_band.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.core_lev_band.restype = C.c_uint32
    lib.core_jaro32.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_double, C.c_uint32]
    lib.core_jaro32.restype = C.c_double
    lib.core_jaro_generic.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_double]
    lib.core_jaro_generic.restype = C.c_double
    return lib


def _pairs(rng, n, qlens, clens, alphabet):
    for _ in range(n):
        l1, l2 = int(rng.choice(qlens)), int(rng.choice(clens))
        a = (rng.integers(0, alphabet, l1) + 97).astype(np.uint8)
        if rng.random() < 0.5 and l1 > 0:
            b = list(a)
            for _ in range(rng.integers(0, 8)):
                op, pos = rng.integers(0, 3), rng.integers(0, len(b) + 1)
                if op == 0 and b:
                    b[min(pos, len(b) - 1)] = 97 + rng.integers(0, alphabet)
                elif op == 1:
                    b.insert(pos, 97 + rng.integers(0, alphabet))
                elif b:
                    del b[min(pos, len(b) - 1)]
            b = np.array(b, dtype=np.uint8)
        else:
            b = (rng.integers(0, alphabet, l2) + 97).astype(np.uint8)
        yield a, b


QL = [1, 2, 3, 7, 8, 16, 31, 32, 33, 40, 63, 64]
CL = [0, 1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 63, 64, 65, 100, 130, 200]


def _score(core, metric, kind, a, b, mis=0, cutoff=None, weights=(1, 1, 1), pw=0.1, quirks=False):
    is_f = orc.result_is_float(metric, kind)
    ou, of = C.c_uint32(0), C.c_double(0.0)
    cu = int(cutoff) if (cutoff is not None and not is_f) else 0
    cf = float(cutoff) if (cutoff is not None and is_f) else 0.0
    r = core.core_score(orc.METRICS[metric], orc.KINDS[kind], a.ctypes.data, len(a), b.ctypes.data, len(b),
                        0 if cutoff is None else 1, cu, cf, weights[0], weights[1], weights[2], pw, 1 if quirks else 0,
                        mis, C.byref(ou), C.byref(of))
    if r == 1:
        return None if math.isnan(of.value) else of.value
    return None if ou.value == 0xFFFFFFFF else ou.value


def test_raw_kernels_vs_textbook(core):
    rng = np.random.default_rng(3)
    for a, b in _pairs(rng, 3000, QL, CL, 4):
        d, l, o = orc.tb("levenshtein", a, b), orc.tb("lcs", a, b), orc.tb("osa", a, b)
        for bits in ((32, 64) if len(a) <= 32 else (64,)):
            mis = int(rng.integers(0, 4))
            assert core.core_raw(0, bits, a.ctypes.data, len(a), b.ctypes.data, len(b), mis) == d, (bytes(a), bytes(b), bits)
            assert core.core_raw(1, bits, a.ctypes.data, len(a), b.ctypes.data, len(b), mis) == l
            assert core.core_raw(2, bits, a.ctypes.data, len(a), b.ctypes.data, len(b), mis) == o, (bytes(a), bytes(b), bits)


INT_METRICS = ["levenshtein", "indel", "lcs_seq", "osa"]


def test_score_algebra_vs_oracle(core):
    rng = np.random.default_rng(5)
    for a, b in _pairs(rng, 1200, [0] + QL, CL, 4):
        mis = int(rng.integers(0, 4))
        for m in INT_METRICS:
            for kind in ("distance", "similarity"):
                for c in (None, 0, 1, 2, 3, 4, 5, 10, 31, 32, 64, 2**64 - 1):
                    if m == "levenshtein" and kind == "similarity" and c is not None:
                        continue  # reference wraps; we return None (tested separately)
                    exp = orc.pair(m, kind, a, b, cutoff=c)
                    got = _score(core, m, kind, a, b, mis, cutoff=c)
                    assert got == exp, (m, kind, bytes(a), bytes(b), c, got, exp)
            for kind in ("normalized_distance", "normalized_similarity"):
                for c in (None, 0.0, 0.1, 0.25, 0.3, 0.5, 0.75, 0.9, 1.0, 1.5, -0.5):
                    exp = orc.pair(m, kind, a, b, cutoff=c)
                    got = _score(core, m, kind, a, b, mis, cutoff=c)
                    assert got == exp, (m, kind, bytes(a), bytes(b), c, got, exp)
        for c in (None, 0.0, 0.3, 0.5, 0.9, 1.0):
            for quirks in (False, True):
                exp = orc.pair("ratio", "similarity", a, b, cutoff=c, reference_quirks=quirks)
                got = _score(core, "ratio", "similarity", a, b, mis, cutoff=c, quirks=quirks)
                assert got == exp, ("ratio", bytes(a), bytes(b), c, quirks, got, exp)
        for w in ((2, 2, 2), (1, 1, 2), (3, 3, 7), (0, 0, 5)):
            for c in (None, 0, 3, 10, 100):
                exp = orc.pair("levenshtein", "distance", a, b, cutoff=c, weights=w)
                got = _score(core, "levenshtein", "distance", a, b, mis, cutoff=c, weights=w)
                assert got == exp, ("wlev", bytes(a), bytes(b), c, w, got, exp)


def test_lev_similarity_cutoff_quirk_q2(core):
    # reference: maximum - usize::MAX wraps (release) / panics (debug); ours: None when dist > max - cutoff
    a, b = np.frombuffer(b"kitten", np.uint8), np.frombuffer(b"sitting", np.uint8)
    assert _score(core, "levenshtein", "similarity", a, b, cutoff=4) == 4
    assert _score(core, "levenshtein", "similarity", a, b, cutoff=5) is None


def test_jaro_family_vs_oracle_bit_exact(core):
    rng = np.random.default_rng(9)
    for a, b in _pairs(rng, 2500, [0] + QL, CL, 5):
        for m in ("jaro", "jaro_winkler"):
            for kind in ("distance", "similarity", "normalized_distance", "normalized_similarity"):
                for c in (None, 0.0, 0.3, 0.6, 0.7, 0.75, 0.85, 0.95, 1.0, 1.1):
                    exp = orc.pair(m, kind, a, b, cutoff=c)
                    got = _score(core, m, kind, a, b, int(rng.integers(0, 4)), cutoff=c)
                    assert got == exp, (m, kind, bytes(a), bytes(b), c, got, exp)
        exp = orc.pair("jaro_winkler", "similarity", a, b, prefix_weight=0.25)
        assert _score(core, "jaro_winkler", "similarity", a, b, pw=0.25) == exp


def test_jaro_generic_multiword_vs_oracle(core):
    rng = np.random.default_rng(21)
    for a, b in _pairs(rng, 1500, [1, 5, 64, 65, 100, 128, 129, 200, 300, 700], CL + [300, 500, 900], 5):
        for c in (0.0, 0.5, 0.8):
            exp = orc.pair("jaro", "similarity", a, b, cutoff=c)
            exp = 0.0 if exp is None else exp   # raw _similarity returns 0.0 below the cutoff
            got = core.core_jaro_generic(a.ctypes.data, len(a), b.ctypes.data, len(b), c)
            if got < c:
                got = 0.0
            assert got == exp, (bytes(a), bytes(b), c, got, exp)


def test_banded_levenshtein_vs_oracle(core):
    """LevBand64 (sliding 64-bit Ukkonen band, any query length, cutoff <= 63) == exact distance or None."""
    rng = np.random.default_rng(33)
    n = 0
    for a, b in _pairs(rng, 2500, [1, 2, 5, 31, 32, 33, 63, 64, 65, 100, 128, 129, 200, 256, 300], CL + [250, 256, 260, 300], 3):
        if len(b) == 0:
            continue
        exact = orc.tb("levenshtein", a, b)
        for k in (0, 1, 2, 3, 5, 8, 16, 31, 32, 33, 48, 62, 63):
            exp = exact if exact <= k else 0xFFFFFFFF
            for ce in (0, 1, 8):
                got = core.core_lev_band(a.ctypes.data, len(a), b.ctypes.data, len(b), k, ce)
                assert got == exp, (bytes(a), bytes(b), k, ce, got, exp)
                n += 1
    assert n > 50000


def test_jaro32_rows_vs_oracle_bit_exact(core):
    """Row-wise 32-bit Jaro passes (query <= 32, truncated candidate <= 64) == the oracle's raw similarity."""
    rng = np.random.default_rng(41)
    n = 0
    for a, b in _pairs(rng, 4000, [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 24, 31, 32], CL, 5):
        for c in (0.0, 0.5, 0.7, 0.8, 0.95, 1.0, 1.1):
            exp = orc.pair("jaro", "similarity", a, b, cutoff=c)
            exp = 0.0 if exp is None else exp
            for extra in (0, 2):
                got = core.core_jaro32(a.ctypes.data, len(a), b.ctypes.data, len(b), c, extra)
                if math.isnan(got):
                    continue  # outside the fast path's domain (long candidate)
                if got < c:
                    got = 0.0
                assert got == exp, (bytes(a), bytes(b), c, extra, got, exp)
                n += 1
    assert n > 20000


def test_div3_exact_is_correctly_rounded():
    """div3_exact(x) == x / 3.0 bit for bit on the values the Jaro formula can produce and on random doubles."""
    import ctypes
    lib = ctypes.CDLL(SO)
    lib.core_div3.argtypes = [ctypes.c_double]
    lib.core_div3.restype = ctypes.c_double
    rng = np.random.default_rng(1)
    xs = [a / b + c / d + e / f for a in range(0, 33, 3) for b in range(1, 65, 7) for c in range(0, 33, 5)
          for d in range(1, 65, 9) for e in range(0, 33, 4) for f in range(1, 33, 5)]
    xs += list(rng.random(20000) * 3.0) + list(rng.random(2000) * 1e-3) + [0.0, 3.0, 1.0, 2.0, 1e-300, 2.9999999999999996]
    for x in xs:
        assert lib.core_div3(x) == x / 3.0, x


def test_simple_metrics_vs_oracle(core):
    """Word-wise Hamming (pad semantics) / Prefix / Postfix == the oracle's restatement of hamming.rs / common.rs."""
    core.core_simple.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    core.core_simple.restype = C.c_uint32
    rng = np.random.default_rng(77)
    for a, b in _pairs(rng, 3000, [0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 31, 33, 64, 100], CL, 3):
        if rng.random() < 0.3 and len(a) and len(b):      # long shared prefix / suffix
            k = int(rng.integers(0, min(len(a), len(b)) + 1))
            b = b.copy()
            if rng.random() < 0.5:
                b[:k] = a[:k]
            elif k:
                b[-k:] = a[-k:]
        assert core.core_simple(0, a.ctypes.data, len(a), b.ctypes.data, len(b)) == orc.pair("hamming", "distance", a, b, pad=True)
        assert core.core_simple(1, a.ctypes.data, len(a), b.ctypes.data, len(b)) == orc.pair("prefix", "similarity", a, b)
        assert core.core_simple(2, a.ctypes.data, len(a), b.ctypes.data, len(b)) == orc.pair("postfix", "similarity", a, b)


def test_weighted_wagner_fischer_vs_oracle_and_textbook(core):
    core.core_wf.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64]
    core.core_wf.restype = C.c_uint64
    rng = np.random.default_rng(88)
    for a, b in _pairs(rng, 1500, [0] + QL, CL, 3):
        for w in ((1, 2, 3), (2, 1, 1), (3, 5, 4), (0, 1, 1), (7, 7, 9), (1, 1, 1)):
            got = core.core_wf(a.ctypes.data, len(a), b.ctypes.data, len(b), *w)
            assert got == orc.tb("levenshtein", a, b, *w), (bytes(a), bytes(b), w)
            assert got == orc.pair("levenshtein", "distance", a, b, weights=w), (bytes(a), bytes(b), w)


def test_damerau_zhao_vs_oracle(core):
    core.core_dl.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    core.core_dl.restype = C.c_uint32
    rng = np.random.default_rng(99)
    for a, b in _pairs(rng, 3000, [0] + QL, CL, 3):
        got = core.core_dl(a.ctypes.data, len(a), b.ctypes.data, len(b))
        assert got == orc.pair("damerau_levenshtein", "distance", a, b), (bytes(a), bytes(b), got)


def test_jaro_epilogue_table_dimensions_cover_every_reachable_pair():
    """The row-wise Jaro kernels look their f64 result up at [original candidate length][common characters]
    [transpositions / 2][prefix] (rf_kernels.cu jaro_epi_table).  They score a pair only when the TRUNCATED candidate
    length (jaro.rs:553-565) is at most 64, so the table's first dimension must cover every ORIGINAL length that can reach
    them: max(64, 131 - 2 * len1).  Exhaustive check of that bound against a restatement of jaro_bounds."""
    def truncated_len2(len1, len2):
        if len2 > len1:
            bound = len2 // 2 - 1
            return min(len2, len1 + bound)
        return len2
    for len1 in range(1, 65):
        l2max = 131 - 2 * len1 if 131 - 2 * len1 > 64 else 64
        reach = [len2 for len2 in range(0, 400) if truncated_len2(len1, len2) <= 64]
        assert max(reach) == l2max, (len1, max(reach), l2max)
        assert reach == list(range(0, l2max + 1)), len1      # contiguous: nothing beyond the bound comes back under 64
