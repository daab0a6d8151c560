#!/usr/bin/env python3
"""e2e streaming call (rf_batch_stream_u8_len8_packed6 / _len8, pinned host buffers, config-2 workload) against the chunk
size knobs: wall time per call, link rate.  One JSON line per setting."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rapidfuzz-rs_b200"))
import numpy as np
import torch
import rapidfuzz_b200 as rf
import synth
from rapidfuzz_b200 import _ffi

L = _ffi.lib()
n = int(1e8 * float(os.environ.get("RF_CFG_SCALE", "1.0")))
q = synth.synth_query(2, 32)
chars, off = synth.synth_corpus(2, q, n, 8, 64, 16, pinned=True)
total = int(off[n])
lens8_t = torch.empty(n, dtype=torch.uint8).pin_memory()
lens8 = lens8_t.numpy()
np.copyto(lens8, np.diff(off.view(np.int64)), casting="unsafe")
packed6, dict64 = rf.pack6(chars, pinned=True)
out_t = torch.empty(n, dtype=torch.uint8).pin_memory()
out = out_t.numpy()
ref = None
qa = np.ascontiguousarray(q)


def call(packed, with_create):
    b = None
    def mk():
        h = _ffi.C.c_void_p()
        _ffi.check(L.rf_batch_create_u8(0, qa.ctypes.data, len(qa), 0, _ffi.C.byref(h)))
        return h
    if not with_create:
        b = mk()
    def fn():
        h = b if b is not None else mk()
        if packed:
            _ffi.check(L.rf_batch_stream_u8_len8_packed6(h, packed6.ctypes.data, dict64.ctypes.data, lens8.ctypes.data, n, 0, None, out.ctypes.data))
        else:
            _ffi.check(L.rf_batch_stream_u8_len8(h, chars.ctypes.data, lens8.ctypes.data, n, 0, None, out.ctypes.data))
        if b is None:
            L.rf_batch_destroy(h)
    fn()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    if b is not None:
        L.rf_batch_destroy(b)
    return dt


for mb, kc in [tuple(int(v) for v in s.split(":")) for s in os.environ.get("RF_CHUNKS", "64:2048,32:1024,128:4096,256:8192,512:16384").split(",")]:
    _ffi.check(L.rf_set_option(b"stream_chunk_mb", mb))
    _ffi.check(L.rf_set_option(b"stream_chunk_kcand", kc))
    for packed in (True, False):
        for with_create in (True, False):
            dt = call(packed, with_create)
            if ref is None:
                ref = out.copy()
            same = bool(np.array_equal(ref, out))
            h2d = ((total + 3) // 4 * 3 if packed else total) + n
            print(json.dumps({"chunk_mb": mb, "chunk_kcand": kc, "packed6": packed, "create_destroy_in_step": with_create, "ms": dt * 1e3,
                              "pairs_per_s": n / dt, "h2d_GBps": h2d / dt / 1e9, "same_result": same}), flush=True)
