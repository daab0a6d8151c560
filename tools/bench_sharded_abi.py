#!/usr/bin/env python3
"""ONE process driving N GPUs through the sharded C ABI (rf_corpus_create_sharded_u8 ...): what a Rust / C host without
PyTorch or MPI gets.  Config-2 shape (query len 32, candidates len 8-64), `--per-gpu` candidates per device.

  python tools/bench_sharded_abi.py --gpus 8 [--per-gpu 50000000]

Measures (host wall clock around the blocking C calls, best of 3): rf_sharded_score_u32 (host vector, no collective),
rf_sharded_score_u32_allgather_device with NCCL / copy engines x 1, 4, 8 pieces, rf_sharded_cdist_topk_u8 (config 5:
10^4 x 10^7), rf_sharded_stream_u32 (all PCIe links at once).  Prints one JSON line; results are checked against the
oracle on a sample and against each other."""
import argparse
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
from rapidfuzz_b200 import _ffi, sharding
from oracle import oracle as orc


def best_of(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--per-gpu", type=int, default=50_000_000)
    ap.add_argument("--c5-queries", type=int, default=10_000)
    ap.add_argument("--c5-candidates", type=int, default=10_000_000)
    ap.add_argument("--only-gather", action="store_true")
    ap.add_argument("--skip-c5", action="store_true")
    ap.add_argument("--skip-gather", action="store_true")
    a = ap.parse_args()
    L = _ffi.lib()
    devs = list(range(a.gpus))
    n = a.per_gpu * a.gpus
    q = synth.synth_query(2, 32)
    t0 = time.perf_counter()
    chars, offsets = synth.synth_corpus(2, q, n, 8, 64, 16, pinned=True)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    sc = sharding.ShardedCorpus(chars, offsets, devs)
    t_build = time.perf_counter() - t0
    sb = sharding.ShardedBatchComparator("levenshtein", q, devs)
    m = 200_000
    exp = orc.batch("levenshtein", "distance", q, chars[: int(offsets[m])], offsets[: m + 1], nthreads=0)
    out = {"n_gpus": a.gpus, "candidates": n, "host_gen_s": round(t_gen, 1), "sharded_corpus_build_s": round(t_build, 2),
           "uses_nccl_for_lists": sc.uses_nccl}
    host = torch.empty(n, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    ca = rf.Args()._c(False)

    def score_host():
        _ffi.check(L.rf_sharded_score_u32(sb._h, sc._h, 0, C.byref(ca), host.ctypes.data))
    ms = best_of(score_host)
    ref = host.copy()
    out["score_to_host_vector"] = {"ms": ms, "pairs_per_s": n / (ms * 1e-3), "matches_oracle_sample": bool(np.array_equal(host[:m], exp))}
    bufs = [torch.empty(n, dtype=torch.int32, device="cuda:%d" % d) for d in devs]
    ptrs = (C.c_void_p * a.gpus)(*[b.data_ptr() for b in bufs])
    out["score_allgather_device"] = {}
    for mode, mname in ((1, "copy_engines"), (2, "nccl")):
        if (a.gpus == 1 and mode == 2) or a.skip_gather:
            continue
        for chunks in (1, 4, 8):
            _ffi.check(L.rf_set_option(b"sharded_collective", mode))
            _ffi.check(L.rf_set_option(b"allgather_chunks", chunks))
            for b in bufs:
                b.fill_(-1)

            def gather():
                _ffi.check(L.rf_sharded_score_u32_allgather_device(sb._h, sc._h, 0, C.byref(ca), ptrs))
            ms = best_of(gather)
            ok = all(bool(torch.equal(b.cpu(), torch.from_numpy(ref.view(np.int32)))) for b in (bufs[0], bufs[-1]))
            out["score_allgather_device"]["%s_%d_pieces" % (mname, chunks)] = {"ms": ms, "pairs_per_s": n / (ms * 1e-3), "all_scores_on_every_device": ok}
    _ffi.check(L.rf_set_option(b"sharded_collective", 0))
    _ffi.check(L.rf_set_option(b"allgather_chunks", 0))
    del bufs

    if a.only_gather:
        print(json.dumps(out))
        return

    def stream():
        sb.stream("distance", chars, offsets, out=host)
    ms = best_of(stream, reps=2)
    out["stream_from_pinned_host"] = {"ms": ms, "pairs_per_s": n / (ms * 1e-3), "h2d_GBps_total": (float(offsets[n]) + 8.0 * n) / ms / 1e6,
                                      "equals_resident_scores": bool(np.array_equal(host, ref))}
    # the dynamically balanced forms (shared chunk planner): plain bytes + length bytes, byte results; 6-bit packed characters
    lens8 = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
    np.copyto(lens8, np.diff(offsets.view(np.int64)), casting="unsafe")
    out8 = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
    ms = best_of(lambda: sb.stream_len8("distance", chars, lens8, out=out8, u8_results=True), reps=2)
    out["stream_len8_dynamic"] = {"ms": ms, "pairs_per_s": n / (ms * 1e-3), "h2d_GBps_total": (float(offsets[n]) + n) / ms / 1e6,
                                  "equals_resident_scores": bool(np.array_equal(out8, ref.astype(np.uint8)))}
    packed, d64 = rf.pack6(chars, pinned=True)
    out8[:] = 0
    ms = best_of(lambda: sb.stream_len8("distance", packed, lens8, out=out8, u8_results=True, dict64=d64), reps=2)
    out["stream_len8_packed6_dynamic"] = {"ms": ms, "pairs_per_s": n / (ms * 1e-3), "h2d_GBps_total": (float(offsets[n]) * 0.75 + n) / ms / 1e6,
                                          "equals_resident_scores": bool(np.array_equal(out8, ref.astype(np.uint8)))}
    sb.close()
    sc.close()
    if a.skip_c5:
        print(json.dumps(out))
        return
    # config 5: many-vs-many over the sharded corpus
    nq, n5, k = a.c5_queries, a.c5_candidates, 10
    qs = np.stack([synth.synth_query(5 + i, 32) for i in range(nq)])
    q_chars, q_off = np.ascontiguousarray(qs.reshape(-1)), np.arange(nq + 1, dtype=np.uint64) * 32
    c5, o5 = synth.synth_corpus(5, qs[0], n5, 8, 64, 16)
    sc5 = sharding.ShardedCorpus(c5, o5, devs)
    res = {}

    def cd():
        res["r"] = sharding.sharded_cdist_topk(q_chars, q_off, sc5, k=k)
    ms = best_of(cd, reps=2)
    gi, gd = res["r"]
    ok = True
    for qi in (0, nq // 2, nq - 1):
        dd = orc.batch("levenshtein", "distance", qs[qi], c5, o5, nthreads=0).astype(np.int64)
        keys = np.sort(dd * (1 << 32) + np.arange(n5))[:k]
        ok = ok and bool(np.array_equal(gi[qi], (keys & 0xFFFFFFFF).astype(np.uint64)) and np.array_equal(gd[qi], (keys >> 32).astype(np.uint32)))
    out["config5_cdist_topk"] = {"queries": nq, "candidates": n5, "ms": ms, "pairs_per_s": float(nq) * n5 / (ms * 1e-3),
                                 "global_topk_matches_oracle_sample_queries": ok}
    sc5.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
