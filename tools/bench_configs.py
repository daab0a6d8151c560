#!/usr/bin/env python3
"""Secondary configurations of BASELINE.json (configs 3, 4, 5 and the 64-bit-word variant of config 2) on one
GPU: device-resident inputs, CUDA-event timing, parity spot-check against the CPU oracle on a sample.
Prints one JSON line per configuration (kept under profiles/ per round)."""
import ctypes as C
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
from oracle import oracle as orc

L = _ffi.lib()
PEAK = 6547.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
scale = float(os.environ.get("RF_CFG_SCALE", "1.0"))


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def one_vs_many(name, metric, kind, seed, qlen, n, lo, hi, kmax, cutoff, out_f64, bytes_per_pair_fn, steps=20, tol=0.0):
    q = synth.synth_query(seed, qlen)
    chars, offsets = synth.synth_corpus(seed, q, n, lo, hi, kmax)
    corpus = rf.Corpus(chars, offsets)
    cls = type("B", (rf._scorer.BatchComparatorBase,), {"METRIC": metric})
    b = cls(q)
    args = rf.Args() if cutoff is None else rf.Args().score_cutoff(cutoff)
    out = torch.empty(n, dtype=torch.float64 if out_f64 else torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ms = timed(lambda: b.score_into(kind, corpus, out.data_ptr(), args, st), steps)
    m = min(n, 200_000)
    kw = {} if cutoff is None else {"cutoff": cutoff}
    exp = orc.batch(metric, kind, q, chars[: int(offsets[m])], offsets[: m + 1], nthreads=0, **kw)
    got = out[:m].cpu().numpy()
    if out_f64:
        ok = bool(np.all((np.isnan(got) & np.isnan(exp)) | (np.abs(got - exp) <= tol)))
        exact = bool(np.all((np.isnan(got) & np.isnan(exp)) | (got == exp)))
    else:
        ok = exact = bool(np.array_equal(got.view(np.uint32), exp))
    t0 = time.perf_counter()
    orc.batch(metric, kind, q, chars[: int(offsets[min(n, 2_000_000)])], offsets[: min(n, 2_000_000) + 1], nthreads=0, **kw)
    cpu = min(n, 2_000_000) / (time.perf_counter() - t0)
    lens = np.diff(offsets.astype(np.int64))
    alg = float(bytes_per_pair_fn(lens).sum())
    pps = n / (ms * 1e-3)
    print(json.dumps({"config": name, "metric": metric, "kind": kind, "n": n, "query_len": qlen, "cutoff": cutoff,
                      "ms_per_step": ms, "pairs_per_s": pps, "algorithmic_GBps": alg / (ms * 1e-3) / 1e9,
                      "hbm_frac_of_measured_peak": alg / (ms * 1e-3) / 1e9 / PEAK, "matches_oracle_sample": ok,
                      "bit_exact": exact, "some_frac": float(np.mean(~np.isnan(got)) if out_f64 else np.mean(got.view(np.uint32) != 0xFFFFFFFF)),
                      "cpu_oracle_pairs_per_s_all_threads": cpu, "cpu_threads": orc.max_threads()}), flush=True)
    b.close()
    corpus.close()


def cdist(nq, n, k=10, steps=2):
    q0 = synth.synth_query(5, 32)
    qs = [synth.synth_query(5 + i, 32) for i in range(nq)]
    q_chars = np.concatenate(qs)
    q_off = (np.arange(nq + 1, dtype=np.uint64) * 32)
    chars, offsets = synth.synth_corpus(5, q0, n, 8, 64, 16)
    corpus = rf.Corpus(chars, offsets)
    idx = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    dist = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    a = _ffi.RfArgs()
    L.rf_args_default(C.byref(a))
    st = torch.cuda.current_stream().cuda_stream

    def run():
        _ffi.check(L.rf_cdist_topk_u8_device(q_chars.ctypes.data, q_off.ctypes.data, nq, corpus._h, C.byref(a), k,
                                             idx.data_ptr(), dist.data_ptr(), st))
    ms = timed(run, steps, warmup=1)
    # parity on a 100 x 100000 sub-problem against the oracle's full matrix
    sub_q, sub_n = min(nq, 100), min(n, 100_000)
    sc = rf.Corpus(chars[: int(offsets[sub_n])], offsets[: sub_n + 1])
    gi, gd = rf.cdist_topk((q_chars[: sub_q * 32], q_off[: sub_q + 1]), sc, k=k)
    ok = True
    for qi in range(sub_q):
        d = orc.batch("levenshtein", "distance", qs[qi], chars[: int(offsets[sub_n])], offsets[: sub_n + 1], nthreads=0).astype(np.int64)
        keys = np.sort(d * (1 << 32) + np.arange(sub_n))[:k]
        ok = ok and np.array_equal(gi[qi], (keys & 0xFFFFFFFF).astype(np.uint32)) and np.array_equal(gd[qi], (keys >> 32).astype(np.uint32))
    sc.close()
    pairs = nq * n
    print(json.dumps({"config": "C5 cdist top-%d (one GPU's shard)" % k, "nq": nq, "n": n, "ms_per_step": ms,
                      "pairs_per_s": pairs / (ms * 1e-3), "matches_oracle_submatrix_%dx%d" % (sub_q, sub_n): bool(ok)}), flush=True)
    corpus.close()


def extract_filter(n):
    """config-2 shape: scan + on-device top-10 / cutoff compaction, host wall clock (includes the k-entry D2H)."""
    q = synth.synth_query(2, 32)
    chars, offsets = synth.synth_corpus(2, q, n, 8, 64, 16)
    corpus = rf.Corpus(chars, offsets)
    b = type("B", (rf._scorer.BatchComparatorBase,), {"METRIC": "levenshtein"})(q)
    exp = orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=0)
    order = np.lexsort((np.arange(n), exp))[:10]

    def wall(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        return (time.perf_counter() - t0) / reps * 1e3, r
    ms_x, (gi, gs) = wall(lambda: b.extract("distance", corpus, k=10))
    ok_x = bool(np.array_equal(gi, order.astype(np.uint32)) and np.array_equal(gs, exp[order]))
    a = rf.Args().score_cutoff(8)
    ms_f, (fi, fs, tot) = wall(lambda: b.filter("distance", corpus, a, capacity=1 << 22))
    hits = np.nonzero(exp <= 8)[0]
    ok_f = bool(tot == len(hits) and np.array_equal(fi, hits[: 1 << 22].astype(np.uint32)))
    out = np.empty(n, dtype=np.uint32)
    ms_full, _ = wall(lambda: _ffi.check(L.rf_batch_score_u32(b._h, corpus._h, 0, None, out.ctypes.data)))
    print(json.dumps({"config": "C2-shape post-processing", "n": n, "extract_top10_ms": ms_x, "extract_matches_oracle": ok_x,
                      "filter_cutoff8_ms": ms_f, "filter_hits": int(tot), "filter_matches_oracle": ok_f,
                      "full_scores_to_pageable_host_ms": ms_full}), flush=True)
    b.close()
    corpus.close()


def widen(n):
    """The steps either side of the path: u32 elements, packing, corpus files."""
    import tempfile
    q = synth.synth_query(2, 32)
    chars, offsets = synth.synth_corpus(2, q, n, 8, 64, 16)
    # --- u32 elements: same strings as code points (+0x400 so that nothing is a byte), alphabet renaming per call
    elems = chars.astype(np.uint32) + 0x400
    q32 = q.astype(np.uint32) + 0x400
    c32 = rf.Corpus.from_u32(elems, offsets)
    b32 = type("B", (rf._scorer.BatchComparatorBase,), {"METRIC": "levenshtein"})(q32)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ms32 = timed(lambda: b32.score_into("distance", c32, out.data_ptr(), None, st), 10)
    m = min(n, 200_000)
    exp = orc.batch("levenshtein", "distance", q, chars[: int(offsets[m])], offsets[: m + 1], nthreads=0)
    ok32 = bool(np.array_equal(out[:m].cpu().numpy().view(np.uint32), exp))
    b32.close(); c32.close(); del elems
    # --- packing: n Python-side (pointer, length) strings -> CSR (rf_pack_u8), all host threads
    ns = min(n, 20_000_000)
    lens = np.diff(offsets[: ns + 1]).astype(np.uint64)
    base = chars.ctypes.data
    ptrs = (base + offsets[:ns]).astype(np.uint64)
    off2 = np.empty(ns + 1, dtype=np.uint64)
    dst = np.empty(int(offsets[ns]), dtype=np.uint8)
    t0 = time.perf_counter()
    _ffi.check(L.rf_pack_u8(ptrs.ctypes.data, lens.ctypes.data, ns, off2.ctypes.data, dst.ctypes.data, 0))
    t_pack = time.perf_counter() - t0
    okp = bool(np.array_equal(dst, chars[: int(offsets[ns])]) and np.array_equal(off2, offsets[: ns + 1]))
    # --- corpus file: write, map + upload (page cache warm), score from the mapping through the streaming pipeline
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c.rfc")
        t0 = time.perf_counter(); rf.write_corpus_file(path, chars, offsets); t_w = time.perf_counter() - t0
        size = os.path.getsize(path)
        t0 = time.perf_counter(); cf = rf.Corpus.from_file(path); t_load = time.perf_counter() - t0
        cf.close()
        b8 = type("B", (rf._scorer.BatchComparatorBase,), {"METRIC": "levenshtein"})(q)
        with rf.CorpusFile(path) as f:
            t0 = time.perf_counter(); r = b8.stream("distance", f.chars, f.offsets); t_stream = time.perf_counter() - t0
            okf = bool(np.array_equal(r[:m], exp))
        b8.close()
    print(json.dumps({"config": "widen: u32 elements / packing / corpus file", "n": n,
                      "u32_levenshtein_ms_per_step": ms32, "u32_pairs_per_s": n / (ms32 * 1e-3), "u32_matches_oracle_sample": ok32,
                      "pack_strings": ns, "pack_s": t_pack, "pack_GBps": float(offsets[ns]) / t_pack / 1e9, "pack_ok": okp,
                      "file_bytes": size, "file_write_s": t_w, "file_map_upload_build_s": t_load,
                      "file_stream_score_s_pageable_mmap": t_stream, "file_stream_matches_oracle_sample": okf}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2w", "c3", "c4", "c5"]
    for kv in os.environ.get("RF_OPTS", "").split(","):   # A/B runs of the library's tuning knobs: RF_OPTS=cdist_skip=0,...
        if kv:
            _ffi.check(L.rf_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1])))
    if "c2w" in which:  # config 2 with a 64-element query (64-bit words)
        one_vs_many("C2 (query len 64, 64-bit words)", "levenshtein", "distance", 2, 64, int(1e8 * scale), 8, 64, 16, None, False,
                    lambda l: l + 8)
    if "c3" in which:
        one_vs_many("C3 multi-block", "levenshtein", "distance", 3, 256, int(1e7 * scale), 64, 256, 48, 32, False,
                    lambda l: np.where(np.abs(256 - l) <= 32, l + 8, 8))
    if "c4" in which:
        one_vs_many("C4 jaro_winkler", "jaro_winkler", "normalized_similarity", 4, 32, int(1e8 * scale), 8, 64, 16, None, True,
                    lambda l: l + 12, tol=1e-6)
    if "c5" in which:
        cdist(int(1e4 * scale), int(1.25e6))
    if "norm" in which:   # the f64-valued kinds of the edit-distance family (8-byte results, division in the epilogue)
        one_vs_many("C2-shape levenshtein normalized_similarity", "levenshtein", "normalized_similarity", 2, 32, int(1e8 * scale),
                    8, 64, 16, None, True, lambda l: l + 12)
        one_vs_many("C2-shape fuzz::ratio", "ratio", "similarity", 2, 32, int(1e8 * scale), 8, 64, 16, None, True, lambda l: l + 12)
        one_vs_many("C2-shape levenshtein distance cutoff 8", "levenshtein", "distance", 2, 32, int(1e8 * scale), 8, 64, 16, 8,
                    False, lambda l: l + 8)
    if "fam" in which:   # the other bit-parallel metrics of the family on the config-2 shape (u32 results)
        for m, kind in (("indel", "distance"), ("lcs_seq", "similarity"), ("osa", "distance")):
            one_vs_many("C2-shape " + m + " " + kind, m, kind, 2, 32, int(1e8 * scale), 8, 64, 16, None, False, lambda l: l + 8)
        one_vs_many("C2-shape indel distance (query len 64)", "indel", "distance", 2, 64, int(1e8 * scale), 8, 64, 16, None, False,
                    lambda l: l + 8)
    if "mw" in which:   # off-config shapes: longer queries without a (small) cutoff, 64-bit-word variants of the other metrics
        one_vs_many("C3-shape levenshtein, no cutoff", "levenshtein", "distance", 3, 256, int(1e7 * scale), 64, 256, 48, None, False, lambda l: l + 8, steps=5)
        one_vs_many("C3-shape levenshtein, cutoff 100", "levenshtein", "distance", 3, 256, int(1e7 * scale), 64, 256, 48, 100, False, lambda l: l + 8, steps=5)
        one_vs_many("C3-shape indel, no cutoff", "indel", "distance", 3, 256, int(1e7 * scale), 64, 256, 48, None, False, lambda l: l + 8, steps=5)
        one_vs_many("C3-shape jaro_winkler (query 256)", "jaro_winkler", "similarity", 3, 256, int(1e7 * scale), 64, 256, 48, None, True, lambda l: l + 12, steps=3)
        one_vs_many("C2-shape jaro_winkler (query 48)", "jaro_winkler", "similarity", 2, 48, int(1e8 * scale), 8, 64, 16, None, True, lambda l: l + 12, steps=5)
        one_vs_many("C2-shape osa (query 48)", "osa", "distance", 2, 48, int(1e8 * scale), 8, 64, 16, None, False, lambda l: l + 8, steps=5)
    if "simple" in which:   # HBM-bound metrics: bytes = len + 4 + 4 per pair
        for m, kind in (("hamming", "distance"), ("prefix", "similarity"), ("postfix", "similarity")):
            q = synth.synth_query(2, 32)
            n = int(1e8 * scale)
            chars, offsets = synth.synth_corpus(2, q, n, 8, 64, 16)
            corpus = rf.Corpus(chars, offsets)
            b = type("B", (rf._scorer.BatchComparatorBase,), {"METRIC": m})(q)
            out = torch.empty(n, dtype=torch.int32, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            a = rf.Args().pad(True)
            ms = timed(lambda: b.score_into(kind, corpus, out.data_ptr(), a, st), 20)
            mm = min(n, 200_000)
            exp = orc.batch(m, kind, q, chars[: int(offsets[mm])], offsets[: mm + 1], nthreads=0, pad=True)
            ok = bool(np.array_equal(out[:mm].cpu().numpy().view(np.uint32), exp))
            alg = float(offsets[n]) + 8.0 * n
            print(json.dumps({"config": "C2-shape " + m, "n": n, "ms_per_step": ms, "pairs_per_s": n / (ms * 1e-3),
                              "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "hbm_frac_of_measured_peak": alg / (ms * 1e-3) / 1e9 / PEAK,
                              "matches_oracle_sample": ok}), flush=True)
            b.close(); corpus.close()
    if "dp" in which:   # the O(len1*len2) DP metrics: generic Levenshtein weights (Wagner-Fischer) and Damerau-Levenshtein
        for m, w in (("levenshtein", (1, 2, 3)), ("damerau_levenshtein", None)):
            q = synth.synth_query(2, 32)
            n = int(1e7 * scale)
            chars, offsets = synth.synth_corpus(2, q, n, 8, 64, 16)
            corpus = rf.Corpus(chars, offsets)
            b = type("B", (rf._scorer.BatchComparatorBase,), {"METRIC": m})(q)
            out = torch.empty(n, dtype=torch.int32, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            a = rf.Args() if w is None else rf.Args().weights(*w)
            ms = timed(lambda: b.score_into("distance", corpus, out.data_ptr(), a, st), 5, warmup=1)
            mm = min(n, 100_000)
            kw = {} if w is None else {"weights": w}
            exp = orc.batch(m, "distance", q, chars[: int(offsets[mm])], offsets[: mm + 1], nthreads=0, **kw)
            ok = bool(np.array_equal(out[:mm].cpu().numpy().view(np.uint32), exp))
            t0 = time.perf_counter()
            orc.batch(m, "distance", q, chars[: int(offsets[min(n, 1_000_000)])], offsets[: min(n, 1_000_000) + 1], nthreads=0, **kw)
            cpu = min(n, 1_000_000) / (time.perf_counter() - t0)
            print(json.dumps({"config": "C2-shape %s%s" % (m, "" if w is None else " weights %s" % (w,)), "n": n, "ms_per_step": ms,
                              "pairs_per_s": n / (ms * 1e-3), "matches_oracle_sample": ok,
                              "cpu_oracle_pairs_per_s_all_threads": cpu, "cpu_threads": orc.max_threads()}), flush=True)
            b.close(); corpus.close()
    if "widen" in which:
        widen(int(1e8 * scale))
    if "post" in which:
        extract_filter(int(1e8 * scale))
    for extra in which:
        if extra in ("indel", "lcs_seq", "osa", "jaro"):
            kind = "similarity" if extra in ("lcs_seq", "jaro") else "distance"
            one_vs_many("C2-shape " + extra, extra, kind, 2, 32, int(1e8 * scale), 8, 64, 16, None, extra == "jaro", lambda l: l + 8)
