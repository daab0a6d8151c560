"""CPU-side checks of the boundary: the C-ABI library loads without a GPU and exports every symbol
include/rfgpu.h declares; argument validation and the no-CPU-fallback rule hold; the synthetic workload
generator is deterministic."""
import ctypes as C
import os
import re

import numpy as np

import rapidfuzz_b200 as rf
import synth
from rapidfuzz_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported():
    hdr = open(os.path.join(ROOT, "include", "rfgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)   # declarations only, not prose
    declared = set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = C.CDLL(_ffi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)


def test_args_default_and_result_types():
    a = _ffi.RfArgs()
    _ffi.lib().rf_args_default(C.byref(a))
    assert (a.has_cutoff, a.has_hint, a.insertion_cost, a.deletion_cost, a.substitution_cost) == (0, 0, 1, 1, 1)
    assert abs(a.prefix_weight - 0.1) < 1e-15
    isf = _ffi.lib().rf_result_is_float
    assert isf(0, 0) == 0 and isf(0, 1) == 0 and isf(0, 2) == 1 and isf(0, 3) == 1
    assert isf(4, 0) == 1 and isf(5, 1) == 1 and isf(6, 1) == 1


def test_no_cpu_fallback_without_device():
    if _ffi.lib().rf_device_count() > 0:
        return
    try:
        rf.Corpus.from_strings([b"abc"])
    except rf.RfError as e:
        assert e.status == _ffi.RF_ERR_CUDA
    else:
        raise AssertionError("corpus creation must fail loudly without a CUDA device")
    try:
        rf.distance.levenshtein.BatchComparator(b"abc")
    except rf.RfError as e:
        assert e.status == _ffi.RF_ERR_CUDA
    else:
        raise AssertionError("batch creation must fail loudly without a CUDA device")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rapidfuzz-rs_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in src.lower() or f == "rf_core.cuh" and "oracle" not in src, (dp, f)


def test_synth_is_deterministic_and_shaped():
    q = synth.synth_query(2, 32)
    assert bytes(q) == bytes(synth.synth_query(2, 32)) and len(q) == 32
    c1, o1 = synth.synth_corpus(2, q, 20000, 8, 64, 16, nthreads=1)
    c2, o2 = synth.synth_corpus(2, q, 20000, 8, 64, 16, nthreads=4)
    assert np.array_equal(c1, c2) and np.array_equal(o1, o2)
    lens = np.diff(o1.astype(np.int64))
    assert lens.min() >= 8 and lens.max() <= 64 and abs(lens.mean() - 36) < 1.0
    alnum = set(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789")
    assert set(np.unique(c1).tolist()) <= alnum
    c3, _ = synth.synth_corpus(3, q, 20000, 8, 64, 16)
    assert not np.array_equal(c1[:1000], c3[:1000])


def test_late_round2_entry_points_validate_arguments_without_a_device():
    """rf_batch_score_u8[_device], rf_corpus_release_csr, rf_corpus_has_csr: NULL handles are refused with
    RF_ERR_INVALID_ARG before anything touches a device (so this runs on the CPU-only build box too)."""
    L = _ffi.lib()
    out = np.zeros(4, np.uint8)
    assert L.rf_batch_score_u8(None, None, 0, None, out.ctypes.data) == _ffi.RF_ERR_INVALID_ARG
    assert L.rf_batch_score_u8_device(None, None, 0, None, out.ctypes.data, None) == _ffi.RF_ERR_INVALID_ARG
    assert L.rf_corpus_release_csr(None) == _ffi.RF_ERR_INVALID_ARG
    assert L.rf_corpus_has_csr(None) == 0
    assert L.rf_set_option(b"epilogue_table", 0) == _ffi.RF_OK and L.rf_set_option(b"epilogue_table", 1) == _ffi.RF_OK
    assert L.rf_set_option(b"jaro32", 2) == _ffi.RF_OK and L.rf_set_option(b"jaro32", 1) == _ffi.RF_OK
    assert L.rf_set_option(b"no_such_option", 1) == _ffi.RF_ERR_INVALID_ARG
