"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs --
never by the product package."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

METRICS = {"levenshtein": 0, "indel": 1, "lcs_seq": 2, "osa": 3, "jaro": 4, "jaro_winkler": 5, "ratio": 6,
           "hamming": 7, "prefix": 8, "postfix": 9, "damerau_levenshtein": 10}
KINDS = {"distance": 0, "similarity": 1, "normalized_distance": 2, "normalized_similarity": 3}
U64_MAX = 2**64 - 1


class OrcArgs(C.Structure):
    _fields_ = [("has_cutoff", C.c_uint8), ("cutoff_u", C.c_uint64), ("cutoff_f", C.c_double),
                ("has_hint", C.c_uint8), ("hint_u", C.c_uint64), ("hint_f", C.c_double),
                ("ins", C.c_uint64), ("del_", C.c_uint64), ("sub", C.c_uint64),
                ("prefix_weight", C.c_double), ("reference_quirks", C.c_uint8), ("pad", C.c_uint8)]


class DifferentLengthArgs(ValueError):
    """hamming::Error::DifferentLengthArgs (hamming.rs:121-136): unequal lengths without Args::pad(true)."""


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "rf_oracle.hpp", "rf_textbook.hpp", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        for name in ("orc_batch_u8", "orc_batch_u32"):
            f = getattr(_lib, name)
            f.argtypes = [C.c_int, C.c_int, vp, C.c_uint64, vp, vp, C.c_uint64, C.POINTER(OrcArgs), vp, vp, C.c_int]
            f.restype = C.c_int
        for name in ("orc_pair_u8", "orc_pair_u32"):
            f = getattr(_lib, name)
            f.argtypes = [C.c_int, C.c_int, vp, C.c_uint64, vp, C.c_uint64, C.POINTER(OrcArgs),
                          C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_int)]
            f.restype = C.c_int
        for name in ("orc_lev_block_u8", "orc_lev_small_band_u8"):
            f = getattr(_lib, name)
            f.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.c_uint64]
            f.restype = C.c_uint64
        _lib.orc_tb_levenshtein_u8.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
        _lib.orc_tb_levenshtein_u8.restype = C.c_uint64
        for name in ("orc_tb_lcs_u8", "orc_tb_osa_u8", "orc_tb_damerau_levenshtein_u8"):
            f = getattr(_lib, name)
            f.argtypes = [vp, C.c_uint64, vp, C.c_uint64]
            f.restype = C.c_uint64
        _lib.orc_tb_jaro_u8.argtypes = [vp, C.c_uint64, vp, C.c_uint64]
        _lib.orc_tb_jaro_u8.restype = C.c_double
        _lib.orc_tb_jaro_winkler_u8.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.c_double]
        _lib.orc_tb_jaro_winkler_u8.restype = C.c_double
        _lib.orc_result_is_float.argtypes = [C.c_int, C.c_int]
        _lib.orc_result_is_float.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def make_args(cutoff=None, hint=None, weights=None, prefix_weight=0.1, reference_quirks=False, is_float=False, pad=False,
              allow_differing=False):
    a = OrcArgs()
    a.pad = 1 if pad else 0
    a.ins, a.del_, a.sub = weights if weights is not None else (1, 1, 1)
    a.prefix_weight = prefix_weight
    a.reference_quirks = 1 if reference_quirks else 0
    if cutoff is not None:
        a.has_cutoff = 1
        if is_float:
            a.cutoff_f = float(cutoff)
        else:
            a.cutoff_u = int(cutoff)
    if hint is not None:
        a.has_hint = 1
        if is_float:
            a.hint_f = float(hint)
        else:
            a.hint_u = int(hint)
    return a


def _as_arr(s, dtype=None):
    """bytes/str/list/ndarray -> contiguous ndarray of u8 or u32 elements."""
    if isinstance(s, np.ndarray):
        return np.ascontiguousarray(s)
    if isinstance(s, (bytes, bytearray)):
        return np.frombuffer(bytes(s), dtype=np.uint8)
    if isinstance(s, str):
        cps = [ord(c) for c in s]
        if dtype is None:
            dtype = np.uint8 if all(c < 256 for c in cps) else np.uint32
        return np.array(cps, dtype=dtype)
    return np.array(list(s), dtype=dtype or np.uint32)


def result_is_float(metric, kind):
    return bool(lib().orc_result_is_float(METRICS[metric], KINDS[kind]))


def pair(metric, kind, s1, s2, dtype=None, **kw):
    """One (query, candidate) score through the oracle's BatchComparator path. Returns value or None."""
    is_f = result_is_float(metric, kind)
    a = make_args(is_float=is_f, **kw)
    q, s = _as_arr(s1, dtype), _as_arr(s2, dtype)
    if q.dtype != s.dtype:
        q, s = q.astype(np.uint32), s.astype(np.uint32)
    fn = lib().orc_pair_u8 if q.dtype == np.uint8 else lib().orc_pair_u32
    ou, of, some = C.c_uint64(0), C.c_double(0.0), C.c_int(0)
    rc = fn(METRICS[metric], KINDS[kind], q.ctypes.data, len(q), s.ctypes.data, len(s), C.byref(a),
            C.byref(ou), C.byref(of), C.byref(some))
    if rc == 3:
        raise DifferentLengthArgs("Differing length arguments provided")
    assert rc == 0
    if not some.value:
        return None
    return of.value if is_f else ou.value


def batch(metric, kind, query, chars, offsets, nthreads=1, **kw):
    """One query vs a packed corpus. Returns u32 array (UINT32_MAX = None) or f64 array (NaN = None)."""
    is_f = result_is_float(metric, kind)
    a = make_args(is_float=is_f, **kw)
    q = _as_arr(query)
    chars = np.ascontiguousarray(chars)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    if q.dtype != chars.dtype:
        q = q.astype(chars.dtype)
    fn = lib().orc_batch_u8 if chars.dtype == np.uint8 else lib().orc_batch_u32
    out = np.empty(n, dtype=np.float64 if is_f else np.uint32)
    rc = fn(METRICS[metric], KINDS[kind], q.ctypes.data, len(q), chars.ctypes.data, offsets.ctypes.data, n,
            C.byref(a), None if is_f else out.ctypes.data, out.ctypes.data if is_f else None, nthreads)
    if rc == 3 and not kw.get("allow_differing"):
        raise DifferentLengthArgs("Differing length arguments provided")
    assert rc in (0, 3)
    return out


def tb(name, a, b, *extra):
    """Textbook DP cross-checks: name in levenshtein|lcs|osa|damerau_levenshtein|jaro|jaro_winkler (u8 only)."""
    a, b = _as_arr(a, np.uint8), _as_arr(b, np.uint8)
    l = lib()
    if name == "levenshtein":
        w = extra if extra else (1, 1, 1)
        return l.orc_tb_levenshtein_u8(a.ctypes.data, len(a), b.ctypes.data, len(b), *w)
    if name == "jaro_winkler":
        return l.orc_tb_jaro_winkler_u8(a.ctypes.data, len(a), b.ctypes.data, len(b), extra[0] if extra else 0.1)
    return getattr(l, "orc_tb_%s_u8" % name)(a.ctypes.data, len(a), b.ctypes.data, len(b))


def lev_block(q, s, cutoff):
    q, s = _as_arr(q, np.uint8), _as_arr(s, np.uint8)
    r = lib().orc_lev_block_u8(q.ctypes.data, len(q), s.ctypes.data, len(s), cutoff)
    return None if r == U64_MAX else r


def lev_small_band(q, s, cutoff):
    q, s = _as_arr(q, np.uint8), _as_arr(s, np.uint8)
    r = lib().orc_lev_small_band_u8(q.ctypes.data, len(q), s.ctypes.data, len(s), cutoff)
    return None if r == U64_MAX else r


def max_threads():
    return lib().orc_max_threads()
