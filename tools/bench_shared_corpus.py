#!/usr/bin/env python3
"""Many metrics / query lengths over ONE resident config-2-shaped corpus (10^8 candidates of 8-64 characters by default):
the corpus is generated, uploaded and laid out once, every case is timed with CUDA events on the resident data and
spot-checked against the CPU oracle on a sample.  One JSON line per case.

  python tools/bench_shared_corpus.py [case,case,...]     RF_CFG_SCALE=0.1 for a 10^7-candidate corpus
cases: lev32 lev64 indel32 osa32 ham pre post jw32 jw48 jaro64 jw32pair jw32r48 jw48pair jw48off levns32 levns32pair ratio32 ratio32pair indel32pair lcs32 (the 48-element Jaro-Winkler query with the row-wise
kernel switched off = the per-lane routine it replaces)"""
import json
import os
import sys

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

CASES = {
    # name: (metric, kind, query length, f64 result, Args builder, comparator options)
    "lev32": ("levenshtein", "distance", 32, False, None, {}),
    "lev64": ("levenshtein", "distance", 64, False, None, {}),
    "indel32": ("indel", "distance", 32, False, None, {}),
    "osa32": ("osa", "distance", 32, False, None, {}),
    "ham": ("hamming", "distance", 32, False, "pad", {}),
    "pre": ("prefix", "similarity", 32, False, None, {}),
    "post": ("postfix", "similarity", 32, False, None, {}),
    "jw32": ("jaro_winkler", "normalized_similarity", 32, True, None, {}),
    "jw48": ("jaro_winkler", "similarity", 48, True, None, {}),
    "jaro64": ("jaro", "similarity", 64, True, None, {}),
    "jw48off": ("jaro_winkler", "similarity", 48, True, None, {"jaro32": 0}),
    "levns32": ("levenshtein", "normalized_similarity", 32, True, None, {}),
    "levns32pair": ("levenshtein", "normalized_similarity", 32, True, None, {"epilogue_table": 0}),
    "ratio32": ("ratio", "similarity", 32, True, None, {}),
    "ratio32pair": ("ratio", "similarity", 32, True, None, {"epilogue_table": 0}),
    "indel32pair": ("indel", "distance", 32, False, None, {"epilogue_table": 0}),
    "lcs32": ("lcs_seq", "similarity", 32, False, None, {}),
    "jw32pair": ("jaro_winkler", "normalized_similarity", 32, True, None, {"jaro32": 2}),   # score algebra per pair instead of the table
    "jw32r48": ("jaro_winkler", "normalized_similarity", 32, True, None, {"jaro32": 3}),    # table, 48-register build (5 CTAs / SM)
    "jw48pair": ("jaro_winkler", "similarity", 48, True, None, {"jaro32": 2}),
}


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


def main():
    which = sys.argv[1].split(",") if len(sys.argv) > 1 else list(CASES)
    for kv in os.environ.get("RF_OPTS", "").split(","):   # process-wide knobs set BEFORE the corpus is created, e.g. build_interleaved_layout=0
        if kv:
            _ffi.check(L.rf_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1])))
    n = int(1e8 * scale)
    q0 = synth.synth_query(2, 32)
    chars, offsets = synth.synth_corpus(2, q0, n, 8, 64, 16)
    corpus = rf.Corpus(chars, offsets)
    lens = np.diff(offsets.astype(np.int64))
    total = float(lens.sum())
    m = min(n, 200_000)
    sub_c, sub_o = chars[: int(offsets[m])], offsets[: m + 1]
    st = torch.cuda.current_stream().cuda_stream
    for name in which:
        metric, kind, qlen, f64, argsk, opts = CASES[name]
        q = q0 if qlen == 32 else synth.synth_query(2, qlen)
        b = type("B", (rf._scorer.BatchComparatorBase,), {"METRIC": metric})(q)
        for k, v in opts.items():
            _ffi.check(L.rf_batch_set_option(b._h, k.encode(), v))
        a = rf.Args().pad(True) if argsk == "pad" else rf.Args()
        out = torch.empty(n, dtype=torch.float64 if f64 else torch.int32, device="cuda")
        ms = timed(lambda: b.score_into(kind, corpus, out.data_ptr(), a, st), 10)
        kw = {"pad": True} if argsk == "pad" else {}
        exp = orc.batch(metric, kind, q, sub_c, sub_o, nthreads=0, **kw)
        got = out[:m].cpu().numpy()
        if f64:
            exact = bool(np.all((np.isnan(got) & np.isnan(exp)) | (got == exp)))
        else:
            exact = bool(np.array_equal(got.view(np.uint32), exp))
        alg = total + (12.0 if f64 else 8.0) * n
        print(json.dumps({"case": name, "metric": metric, "kind": kind, "query_len": qlen, "n": n, "options": opts, "process_options": os.environ.get("RF_OPTS", ""), "ms_per_step": ms,
                          "pairs_per_s": n / (ms * 1e-3), "algorithmic_GBps": alg / (ms * 1e-3) / 1e9,
                          "hbm_frac_of_measured_peak": alg / (ms * 1e-3) / 1e9 / PEAK, "bit_exact_vs_oracle_sample": exact}), flush=True)
        b.close()
        del out
    corpus.close()


if __name__ == "__main__":
    main()
