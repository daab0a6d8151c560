#!/usr/bin/env python3
"""BASELINE config 5: many-vs-many Levenshtein top-k, 10^4 queries x 10^7 candidates, the corpus sharded by candidate
across the GPUs of one node (one process per GPU).  Per step every rank scans its resident shard for all queries
(rf_cdist_topk_u8_device), the per-shard [nq,k] lists are exchanged with ONE NCCL all-gather and merged on the device
(rf_topk_merge_device).  Strong scaling: the total problem is fixed, value = nq * n / max-over-ranks step time.

  python tools/bench_cdist_sharded.py                      # 1 GPU, whole corpus
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
      tools/bench_cdist_sharded.py --gpus 8
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rapidfuzz-rs_b200"))
import numpy as np
import torch
import torch.distributed as dist
import rapidfuzz_b200 as rf
import synth
from rapidfuzz_b200 import sharding


def sm_clock(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        return pynvml.nvmlDeviceGetClockInfo(pynvml.nvmlDeviceGetHandleByIndex(index), pynvml.NVML_CLOCK_SM)
    except Exception:
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--candidates", type=int, default=10_000_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--check-queries", type=int, default=6)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nq, n, k = a.queries, a.candidates, a.k
    qs = [synth.synth_query(5 + i, 32) for i in range(nq)]
    q_chars = np.concatenate(qs)
    q_off = np.arange(nq + 1, dtype=np.uint64) * 32
    chars, offsets = synth.synth_corpus(5, qs[0], n, 8, 64, 16)          # same corpus on every rank, each keeps its shard
    c_loc, o_loc, lo = sharding.local_shard(chars, offsets, world, rank)
    corpus = rf.Corpus(c_loc, o_loc, device=local)
    t_scan, t_all = [], []

    def step():
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        idx, d = sharding.cdist_topk_device(q_chars, q_off, corpus, k=k, device=dev)
        e1.record()
        if world > 1:
            gi, gd = sharding.all_gather_topk_device(idx, d, lo, k)
        else:
            gi, gd = sharding.merge_topk_device(torch.stack([idx, d], 0).unsqueeze(0).contiguous(),
                                                torch.tensor([lo], dtype=torch.int64, device=dev), k)
        e2.record()
        torch.cuda.synchronize()
        return gi, gd, e0.elapsed_time(e1), e0.elapsed_time(e2)

    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for _ in range(a.steps):
        gi, gd, ms_scan, ms_all = step()
        t_scan.append(ms_scan)
        t_all.append(ms_all)
    ms = torch.tensor([sum(t_all) / len(t_all), sum(t_scan) / len(t_scan)], dtype=torch.float64, device=dev)
    mine = torch.tensor([float(ms[1]), float(sm_clock(local))], dtype=torch.float64, device=dev)
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    else:
        per_rank = [mine]
    ms_all, ms_scan = float(ms[0]), float(ms[1])
    if rank == 0:
        from oracle import oracle as orc   # checker only: global top-k of the first queries over the WHOLE corpus
        ok = True
        for qi in range(min(a.check_queries, nq)):
            dd = orc.batch("levenshtein", "distance", qs[qi], chars, offsets, nthreads=0).astype(np.int64)
            keys = np.sort(dd * (1 << 32) + np.arange(n))[:k]
            ok = ok and np.array_equal(gi[qi].cpu().numpy(), keys & 0xFFFFFFFF) and np.array_equal(gd[qi].cpu().numpy(), keys >> 32)
        print(json.dumps({"metric": "levenshtein_cdist_topk_pairs_per_sec", "value": nq * n / (ms_all * 1e-3), "unit": "pairs/s",
                          "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_all, "scaling": "strong",
                          "config": {"workload": "config5: %d queries len 32 x %d candidates len 8-64, top-%d, corpus sharded by "
                                                 "candidate (byte-balanced), one NCCL all-gather of the per-shard lists + device merge" % (nq, n, k),
                                     "scan_ms_max_over_ranks": ms_scan, "gather_merge_ms": ms_all - ms_scan,
                                     "scan_ms_per_rank": [round(float(t[0]), 2) for t in per_rank],
                                     "sm_mhz_after_last_step_per_rank": [int(t[1]) for t in per_rank],
                                     "global_topk_matches_oracle_first_%d_queries" % min(a.check_queries, nq): bool(ok)}}), flush=True)
    corpus.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
