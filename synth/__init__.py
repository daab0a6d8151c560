"""Deterministic synthetic workloads for bench.py and the tests: a tiny host-only library
(synth/librfsynth.so, g++ -fopenmp) of its own, so that neither the product library depends on a test utility nor the
CPU reference arm of bench.py maps the product library."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librfsynth.so")
_lib = None


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("rf_synth.cpp", "rfsynth.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in src):
        return LIB_PATH
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-shared", "-o", LIB_PATH, src[0]])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB_PATH)
        l.rf_synth_query_u8.restype = C.c_int
        l.rf_synth_query_u8.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p]
        l.rf_synth_corpus_u8.restype = C.c_int
        l.rf_synth_corpus_u8.argtypes = [C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_void_p, C.c_void_p, C.c_int]
        _lib = l
    return _lib


def synth_query(seed, length):
    out = np.empty(length, dtype=np.uint8)
    if lib().rf_synth_query_u8(seed, length, out.ctypes.data) != 0:
        raise ValueError("rf_synth_query_u8: invalid argument")
    return out


def synth_corpus(seed, query, n, min_len, max_len, kmax, nthreads=0, pinned=False):
    """Deterministic synthetic candidates. Returns (chars u8, offsets u64); pinned=True allocates both as page-locked
    torch tensors (bench.py's end-to-end leg)."""
    query = np.ascontiguousarray(query, dtype=np.uint8)
    l = lib()
    if pinned:
        import torch
        offsets_t = torch.empty(n + 1, dtype=torch.int64).pin_memory()
        offsets = offsets_t.numpy().view(np.uint64)
    else:
        offsets = np.empty(n + 1, dtype=np.uint64)
    if l.rf_synth_corpus_u8(seed, query.ctypes.data, len(query), n, min_len, max_len, kmax, offsets.ctypes.data, None, nthreads) != 0:
        raise ValueError("rf_synth_corpus_u8: invalid argument")
    total = int(offsets[n])
    if pinned:
        chars_t = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
        chars = chars_t.numpy()[:total]
    else:
        chars = np.empty(total, dtype=np.uint8)
    if l.rf_synth_corpus_u8(seed, query.ctypes.data, len(query), n, min_len, max_len, kmax, offsets.ctypes.data,
                            chars.ctypes.data, nthreads) != 0:
        raise ValueError("rf_synth_corpus_u8: invalid argument")
    return chars, offsets
