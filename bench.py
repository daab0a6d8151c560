#!/usr/bin/env python3
"""bench.py -- headline benchmark of the one-vs-many hot path (BASELINE.json: "Levenshtein pairs/sec
(len<=64, one-vs-many) at 1/2/4/8 B200; achieved HBM GB/s").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--configs gather,3,4,5]

Headline (the JSON line's metric / value / e2e / roofline): one step = one pass of
levenshtein::BatchComparator::distance over the rank's resident corpus shard (config 2: 1 ASCII query len 32 vs
10^8 candidates len 8-64 per GPU, synthetic).  N > 1: one process per GPU (torchrun),
candidates sharded by rank, weak scaling (every GPU holds a full 10^8-candidate shard), no collective inside the
headline's timed region (the path has none: independent pairs).

The same line carries a `configs` block with the other BASELINE.json configurations, each timed on the device with
its own roofline:
  c2_gather  config 2 + the path's only exchange step: NCCL all-gather of the per-shard score vectors (N > 1)
  c3         query len 256 vs 10^7 candidates len 64-256, score_cutoff 32 (banded multi-word kernels)
  c4         jaro_winkler normalized_similarity, query len 32 vs 10^8 candidates
  c5         many-vs-many: 10^4 queries x 10^7 candidates, top-10, corpus sharded by candidate over the N ranks
             (STRONG scaling), scan + NCCL all-gather of the per-shard lists + device merge all inside the timed region
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rapidfuzz-rs_b200"))

import numpy as np

METRIC = "levenshtein_pairs_per_sec_one_vs_many_len_le_64"
UNIT = "pairs/s"
QUERY_LEN, MIN_LEN, MAX_LEN, KMAX, SEED = 32, 8, 64, 16, 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--candidates", dest="n", type=int, default=int(os.environ.get("RF_BENCH_N", 100_000_000)),
                    help="candidates per GPU (default 10^8 = BASELINE config 2)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default=os.environ.get("RF_BENCH_CONFIGS", "gather,3,4,5"),
                    help="comma list of the secondary configurations to run (gather,3,4,5; empty = none)")
    ap.add_argument("--c5-queries", type=int, default=10_000)
    ap.add_argument("--c5-candidates", type=int, default=10_000_000)
    return ap.parse_args()


def config2_dict(n):
    """The workload description both arms print verbatim (the driver compares the two `config` objects)."""
    return {"workload": "config2: levenshtein::BatchComparator::distance one-vs-many, 1 ASCII query len %d vs %d candidates "
                        "len %d-%d per GPU (SplitMix64 seed %d, 62-symbol alphabet, 1/64 planted near-matches)"
                        % (QUERY_LEN, n, MIN_LEN, MAX_LEN, SEED),
            "candidates_per_gpu": n, "query_len": QUERY_LEN, "min_len": MIN_LEN, "max_len": MAX_LEN, "seed": SEED,
            "sharding": "candidates by rank (rank r: seed + 7919 r), weak scaling",
            "l2": "inputs (about 4 GB per GPU) are larger than the 126 MB L2",
            "timing": "CUDA events on the launching stream, max over ranks"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/<name>)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        return float(d["dram_bytes_per_launch"]), float(d.get("candidates", 1e8))
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock + clock-event reasons through NVML from a thread while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _reason_names(self, mask):
        nv = self.nv
        table = [("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"),
                 ("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("hw_power_brake", "nvmlClocksThrottleReasonHwPowerBrakeSlowdown"),
                 ("sync_boost", "nvmlClocksThrottleReasonSyncBoost"),
                 ("app_clocks", "nvmlClocksThrottleReasonApplicationsClocksSetting")]
        return [n for n, attr in table if hasattr(nv, attr) and (mask & getattr(nv, attr))]

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.reasons.update(self._reason_names(int(mask)))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


ORIG_AFFINITY = None


def numa_bind_to_gpu(index):
    """Place this rank's host buffers on the NUMA node its GPU hangs off: set_mempolicy(MPOL_PREFERRED, node) for all later
    page allocations (the pinned buffers), and CPU affinity narrowed to that node's cores when the cgroup allows any of
    them.  Round 1's 8-GPU run had every rank's pinned memory on one node (e2e efficiency 0.37).  Best effort: a
    container may refuse the syscall; what happened is reported in the JSON line."""
    info = {"gpu_numa_node": None, "mempolicy": "unchanged", "affinity": "unchanged"}
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())
        node = int(open(path).read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes"] = len(nodes)
        if len(nodes) < 2:
            return info
        libc = C.CDLL(None, use_errno=True)
        mask = C.c_ulong(1 << node)
        rc = libc.syscall(238, 1, C.byref(mask), 64)      # set_mempolicy(MPOL_PREFERRED, &mask, maxnode)
        info["mempolicy"] = "preferred node %d" % node if rc == 0 else "refused (errno %d)" % C.get_errno()
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        both = cpus & allowed
        if both and both != allowed:
            global ORIG_AFFINITY
            ORIG_AFFINITY = allowed
            os.sched_setaffinity(0, both)
            info["affinity"] = "%d of the node's %d cores" % (len(both), len(cpus))
        elif not both:
            info["affinity"] = "none of the node's cores is in this cgroup (allowed: %d cpus)" % len(allowed)
    except Exception as e:
        info["error"] = "%s: %s" % (type(e).__name__, e)
    return info


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm must not obey it)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_oracle_rate(n_sample, threads, reps=1):
    """Times the CPU oracle (port of the reference path) on the first n_sample candidates of the workload."""
    from oracle import oracle as orc
    import synth
    q = synth.synth_query(SEED, QUERY_LEN)
    chars, offsets = synth.synth_corpus(SEED, q, n_sample, MIN_LEN, MAX_LEN, KMAX)
    orc.batch("levenshtein", "distance", q, chars[: int(offsets[1000])], offsets[:1001], nthreads=threads)  # warm
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_sample / best, best


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate cannot be built
    in this image (no cargo/rustc), so this is the oracle port (oracle/rf_oracle.hpp) on all host threads.
    Imports only the oracle and the synthetic generator (its own library): the product library is never mapped."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    import synth
    threads = host_threads()
    q = synth.synth_query(SEED, QUERY_LEN)
    # the GPU arm's workload: args.n candidates per step.  A step is cut to a bounded sample (the first n_sample
    # candidates of the same corpus) only when (steps + warmup) full passes would not fit in a few minutes.
    probe_n = min(args.n, 4_000_000)
    pc, po = synth.synth_corpus(SEED, q, probe_n, MIN_LEN, MAX_LEN, KMAX)
    orc.batch("levenshtein", "distance", q, pc, po, nthreads=threads)
    t0 = time.perf_counter()
    orc.batch("levenshtein", "distance", q, pc, po, nthreads=threads)
    rate = probe_n / (time.perf_counter() - t0)
    budget_s = float(os.environ.get("RF_REF_BUDGET_S", "150"))
    n_sample = int(min(args.n, max(1_000_000, rate * budget_s / max(1, args.steps + args.warmup))))
    chars, offsets = (pc, po) if n_sample == probe_n else synth.synth_corpus(SEED, q, n_sample, MIN_LEN, MAX_LEN, KMAX)
    for _ in range(args.warmup):
        orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.batch("levenshtein", "distance", q, chars, offsets, nthreads=threads)
    dt = time.perf_counter() - t0
    value = n_sample * args.steps / dt
    sample = ("%s %d candidates of the config-2 workload per step (seed %d)"
              % ("all" if n_sample == args.n else "first", n_sample, SEED))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config2_dict(args.n),
        "note": "CPU oracle port of the rapidfuzz-rs BatchComparator path (oracle/rf_oracle.hpp), OpenMP static partition "
                "over candidates on all host threads; the Rust reference itself is single-threaded and cannot be built here",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------- secondary configs
class Ctx:
    pass


def dev_timed(ctx, fn, steps, warmup):
    """CUDA events on the launching stream around `steps` calls of fn, barrier + synchronize on both sides, max over ranks."""
    torch = ctx.torch
    for _ in range(warmup):
        fn()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ctx.L.rf_kernel_launch_count()
    e0.record(ctx.stream)
    for _ in range(steps):
        fn()
    e1.record(ctx.stream)
    ctx.barrier()
    ms = e0.elapsed_time(e1) / steps
    launches = int(ctx.L.rf_kernel_launch_count() - n0) // max(1, steps)
    return ctx.max_over_ranks(ms), launches


def roofline(alg_bytes, ms, kernel, traffic=None, note=None):
    peak, src = measured_peak()
    ach = alg_bytes / (ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
         "peak_source": src, "algorithmic_bytes_per_launch": int(alg_bytes), "kernel": kernel}
    if note:
        r["note"] = note
    return r


def oracle_sample_ok(metric, kind, q, chars, offsets, got, m, cutoff=None, tol=None):
    from oracle import oracle as orc
    kw = {} if cutoff is None else {"cutoff": cutoff}
    exp = orc.batch(metric, kind, q, chars[: int(offsets[m])], offsets[: m + 1], nthreads=0, **kw)
    if got.dtype == np.float64:
        both_nan = np.isnan(got) & np.isnan(exp)
        return bool(np.all(both_nan | (got == exp))) if tol is None else bool(np.all(both_nan | (np.abs(got - exp) <= tol)))
    return bool(np.array_equal(got.view(np.uint32), exp))


def run_c2_gather(ctx, corpus, batch, out_dev, n, steps):
    """config 2 + the path's one exchange step (north_star: 'NCCL all-gather only for the final score vector'): every
    rank ends up with all N shards' scores.  Three ways: (a) scan, then torch.distributed all_gather_into_tensor;
    (b) the library's own rf_batch_score_u32_allgather_device over an rf_comm (ncclCommInitRank inside librfgpu.so):
    scan, then gather; (c) the same with the shard scanned in 4 pieces, piece k on NVLink while piece k+1 is scanned."""
    torch, dist, L, _ffi = ctx.torch, ctx.dist, ctx.L, ctx.ffi
    if dist is None:
        return {"skipped": "single GPU: nothing to gather"}
    full = torch.empty(ctx.world * n, dtype=torch.int32, device="cuda")
    t_scan = []

    def step():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        _ffi.check(L.rf_batch_score_u32_device(batch, corpus, _ffi.KINDS["distance"], None, out_dev.data_ptr(), ctx.sptr))
        e1.record(ctx.stream)
        dist.all_gather_into_tensor(full, out_dev)
        t_scan.append((e0, e1))

    ms, launches = dev_timed(ctx, step, steps, 2)
    scan_ms = ctx.max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in t_scan[-steps:]])))
    ok = bool(torch.equal(full[ctx.rank * n: ctx.rank * n + 4096], out_dev[:4096]))
    ref = full.clone()
    gathered = 4.0 * n * (ctx.world - 1)   # bytes every rank receives
    # the library's communicator: rank 0's id reaches the others through the host's own channel (here torch.distributed)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if ctx.rank == 0:
        buf = C.create_string_buffer(128)
        _ffi.check(L.rf_comm_unique_id(buf))
        uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    comm = C.c_void_p()
    _ffi.check(L.rf_comm_create_rank(bytes(uid.cpu().numpy().tobytes()), ctx.world, ctx.rank, ctx.local_rank, C.byref(comm)))
    lib = {}
    for label, chunks in (("scan_then_gather", 1), ("overlapped_4_pieces", 4), ("overlapped_8_pieces", 8)):
        _ffi.check(L.rf_set_option(b"allgather_chunks", chunks))
        full.fill_(-1)

        def lstep():
            _ffi.check(L.rf_batch_score_u32_allgather_device(batch, corpus, comm, _ffi.KINDS["distance"], None, full.data_ptr(),
                                                             ctx.world * n, None, ctx.sptr))
        lms, _ = dev_timed(ctx, lstep, steps, 2)
        lib[label] = {"ms_per_step": lms, "pairs_per_s": ctx.world * n / (lms * 1e-3), "equals_torch_gather": bool(torch.equal(full, ref))}
    _ffi.check(L.rf_set_option(b"allgather_chunks", 4))
    L.rf_comm_destroy(comm)
    best = min(v["ms_per_step"] for v in lib.values())
    return {"what": "config 2 scan + all-gather of the u32 score vectors over NVLink; every rank holds all %d x %d scores afterwards"
                    % (ctx.world, n),
            "ms_per_step": best, "pairs_per_s": ctx.world * n / (best * 1e-3),
            "torch_all_gather_into_tensor": {"ms_per_step": ms, "scan_ms": scan_ms, "gather_ms": ms - scan_ms,
                                             "gather_share": (ms - scan_ms) / ms, "own_slice_matches": ok,
                                             "gather_recv_GBps_per_rank": gathered / max(ms - scan_ms, 1e-6) / 1e6},
            "rf_batch_score_u32_allgather_device": lib, "recv_bytes_per_rank": int(gathered),
            "overlap_gain_vs_scan_then_gather": lib["scan_then_gather"]["ms_per_step"] / best, "scaling": "weak"}


def run_c3(ctx, steps):
    import synth
    import rapidfuzz_b200 as rf
    torch = ctx.torch
    n, qlen, cutoff = 10_000_000, 256, 32
    q = synth.synth_query(3, qlen)
    chars, offsets = synth.synth_corpus(3 + 7919 * ctx.rank, q, n, 64, 256, 48, nthreads=ctx.gen_threads)
    corpus = rf.Corpus(chars, offsets.astype(np.uint32), device=ctx.local_rank)
    b = rf.distance.levenshtein.BatchComparator(q, device=ctx.local_rank)
    args = rf.Args().score_cutoff(cutoff)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    ms, launches = dev_timed(ctx, lambda: b.score_into("distance", corpus, out.data_ptr(), args, ctx.sptr), steps, 3)
    lens = np.diff(offsets.astype(np.int64))
    surv = np.abs(lens - qlen) <= cutoff
    alg = float((lens[surv] + 8).sum() + 8 * (~surv).sum())   # survivors len+4+4, filtered 4+4
    m = 200_000
    got = out[:m].cpu().numpy().view(np.uint32)
    res = {"workload": "config3: levenshtein distance, query len 256 vs 10^7 candidates len 64-256 per GPU, score_cutoff 32 (seed 3)",
           "ms_per_step": ms, "pairs_per_s": ctx.world * n / (ms * 1e-3), "gpu_launches_per_step": launches,
           "within_cutoff_frac_sample": float(np.mean(got != 0xFFFFFFFF)),
           "results_match_oracle_sample": oracle_sample_ok("levenshtein", "distance", q, chars, offsets, got, m, cutoff) if ctx.rank == 0 else None,
           "roofline": roofline(alg, ms, "band_classify_kernel + band_run_kernel<A> + band_run_kernel<B> (3 launches = one pass)",
                                note="algorithmic bytes per SURVEY 8d: length-filter survivors len+8, others 8 (about 49 B/pair)"),
           "scaling": "weak"}
    b.close()
    corpus.close()
    return res


def run_c4(ctx, corpus_h, q, chars, offsets64, n, total, steps):
    """Jaro-Winkler over the resident config-2 shard (config 4 has the same shape: query len 32, 10^8 candidates len 8-64)."""
    torch, L, _ffi = ctx.torch, ctx.L, ctx.ffi
    h = C.c_void_p()
    _ffi.check(L.rf_batch_create_u8(_ffi.METRICS["jaro_winkler"], q.ctypes.data, len(q), ctx.local_rank, C.byref(h)))
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    kind = _ffi.KINDS["normalized_similarity"]
    ms, launches = dev_timed(ctx, lambda: _ffi.check(L.rf_batch_score_f64_device(h, corpus_h, kind, None, out.data_ptr(), ctx.sptr)), steps, 3)
    m = min(n, 200_000)
    got = out[:m].cpu().numpy()
    ok = exact = None
    if ctx.rank == 0:
        ok = oracle_sample_ok("jaro_winkler", "normalized_similarity", q, chars, offsets64, got, m, tol=1e-6)
        exact = oracle_sample_ok("jaro_winkler", "normalized_similarity", q, chars, offsets64, got, m)
    L.rf_batch_destroy(h)
    return {"workload": "config4: jaro_winkler::BatchComparator::normalized_similarity, query len 32 vs %d candidates len 8-64 per GPU "
                        "(the resident config-2 shard: same shape), prefix_weight 0.1, no cutoff, f64 results" % n,
            "ms_per_step": ms, "pairs_per_s": ctx.world * n / (ms * 1e-3), "gpu_launches_per_step": launches,
            "results_within_1e-6_of_oracle_sample": ok, "results_bit_exact_sample": exact,
            "roofline": roofline(total + 12.0 * n, ms, "scan_jaro32_kernel (+ jaro32_long_kernel, empty)",
                                 note="algorithmic bytes: len + 4 (offset) + 8 (f64 result) per pair"),
            "scaling": "weak"}


def run_c5(ctx, nq, n, steps, k=10):
    """config 5: nq x n many-vs-many, corpus sharded by candidate (byte-balanced) over the ranks -> STRONG scaling.
    Timed region per step: rf_cdist_topk_u8_device on the shard, ONE NCCL all-gather of the stacked per-shard lists
    (+ one of the shard starts), rf_topk_merge_device; every rank ends with the global top-k."""
    import synth
    import rapidfuzz_b200 as rf
    from rapidfuzz_b200 import sharding
    torch, dist = ctx.torch, ctx.dist
    dev = torch.device("cuda", ctx.local_rank)
    qs = np.stack([synth.synth_query(5 + i, 32) for i in range(nq)])
    q_chars = np.ascontiguousarray(qs.reshape(-1))
    q_off = np.arange(nq + 1, dtype=np.uint64) * 32
    chars, offsets = synth.synth_corpus(5, qs[0], n, 8, 64, 16, nthreads=ctx.gen_threads)   # same corpus on every rank
    c_loc, o_loc, lo = sharding.local_shard(chars, offsets, ctx.world, ctx.rank)
    corpus = rf.Corpus(c_loc, o_loc, device=ctx.local_rank)
    scan_ev, last = [], {}

    def step():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        idx, d = sharding.cdist_topk_device(q_chars, q_off, corpus, k=k, device=dev)
        e1.record(ctx.stream)
        if dist is not None:
            gi, gd = sharding.all_gather_topk_device(idx, d, lo, k)
        else:
            gi, gd = sharding.merge_topk_device(torch.stack([idx, d], 0).unsqueeze(0).contiguous(),
                                                torch.tensor([lo], dtype=torch.int64, device=dev), k)
        scan_ev.append((e0, e1))
        last["gi"], last["gd"] = gi, gd

    ms, launches = dev_timed(ctx, step, steps, 1)
    scan_local = float(np.mean([a.elapsed_time(b) for a, b in scan_ev[-steps:]]))
    scan_ms = ctx.max_over_ranks(scan_local)
    per_rank = ctx.gather_floats(scan_local)
    ok = None
    if ctx.rank == 0:   # checker: the oracle's global top-k over the WHOLE corpus for a few queries
        from oracle import oracle as orc
        ok = True
        gi, gd = last["gi"].cpu().numpy(), last["gd"].cpu().numpy()
        for qi in sorted({0, 1 % nq, nq // 2, nq - 1}):
            dd = orc.batch("levenshtein", "distance", qs[qi], chars, offsets, nthreads=0).astype(np.int64)
            keys = np.sort(dd * (1 << 32) + np.arange(n))[:k]
            ok = ok and bool(np.array_equal(gi[qi], keys & 0xFFFFFFFF) and np.array_equal(gd[qi], keys >> 32))
    corpus.close()
    coll = ("one NCCL all_gather_into_tensor of [2][nq][k] i32 per rank (%.2f MB) + shard starts, then rf_topk_merge_device"
            % (8.0 * nq * k / 1e6)) if dist is not None else "none (1 GPU): rf_topk_merge_device only"
    return {"workload": "config5: levenshtein cdist top-%d, %d queries len 32 x %d candidates len 8-64 (seed 5), corpus sharded by "
                        "candidate over %d GPU(s), byte-balanced" % (k, nq, n, ctx.world),
            "ms_per_step": ms, "pairs_per_s": float(nq) * n / (ms * 1e-3), "scan_ms_max_over_ranks": scan_ms,
            "gather_merge_ms": ms - scan_ms, "collective_share": (ms - scan_ms) / ms,
            "scan_ms_per_rank": [round(x, 2) for x in per_rank], "gpu_launches_per_step": launches,
            "collective": coll, "global_topk_matches_oracle_sample_queries": ok, "scaling": "strong",
            "roofline": {"bound": "alu", "note": "compute-bound by construction: the shard (%.0f MB interleaved) is served from L2 after "
                         "the first pass; reported as pairs/s (SURVEY 8d)" % (len(c_loc) * 1.1 / 1e6)}}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # launched without torchrun: re-exec under it, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import synth
    from rapidfuzz_b200 import _ffi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa = numa_bind_to_gpu(local_rank) if os.environ.get("RF_BENCH_NUMA", "1") != "0" else {"disabled": True}
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _ffi.lib()
    if os.environ.get("RF_W1_PATH"):  # dev knob: 1 = CSR/TMA-tile kernel, 2 = interleaved layout through per-warp TMA rings
        _ffi.check(L.rf_set_option(b"single_word_path", int(os.environ["RF_W1_PATH"])))
    n = args.n
    gen_threads = max(1, host_threads() // world)
    configs = [c for c in args.configs.split(",") if c]

    # ---- synthetic shard of this rank, generated into pinned host memory
    q = synth.synth_query(SEED, QUERY_LEN)
    t_gen = time.perf_counter()
    chars, offsets64 = synth.synth_corpus(SEED + 7919 * rank, q, n, MIN_LEN, MAX_LEN, KMAX, nthreads=gen_threads, pinned=True)
    total = int(offsets64[n])
    assert total < 2**32 - 16
    off32_t = torch.empty(n + 1, dtype=torch.int32).pin_memory()
    offsets32 = off32_t.numpy().view(np.uint32)
    np.copyto(offsets32, offsets64, casting="unsafe")
    t_gen = time.perf_counter() - t_gen
    out_host_t = torch.empty(n, dtype=torch.int32).pin_memory()
    out_host = out_host_t.numpy().view(np.uint32)

    def create_corpus():
        h = C.c_void_p()
        _ffi.check(L.rf_corpus_create_u8_off32(chars.ctypes.data, offsets32.ctypes.data, n, local_rank, C.byref(h)))
        return h

    def create_batch():
        h = C.c_void_p()
        _ffi.check(L.rf_batch_create_u8(_ffi.METRICS["levenshtein"], q.ctypes.data, len(q), local_rank, C.byref(h)))
        return h

    corpus = create_corpus()
    batch = create_batch()
    out_dev = torch.empty(n, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def step():
        _ffi.check(L.rf_batch_score_u32_device(batch, corpus, _ffi.KINDS["distance"], None, out_dev.data_ptr(), sptr))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def gather_floats(x):
        if dist is None:
            return [float(x)]
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        outs = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [float(o[0]) for o in outs]

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    launches0 = L.rf_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = int(L.rf_kernel_launch_count() - launches0)

    # ---- correctness spot check of the timed output (oracle as checker, rank 0, first 200k candidates)
    ok = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        m = min(n, 200_000)
        exp = orc.batch("levenshtein", "distance", q, chars[: int(offsets64[m])], offsets64[: m + 1], nthreads=0)
        got = out_dev[:m].cpu().numpy().view(np.uint32)
        ok = bool(np.array_equal(got, exp))

    # ---- the other BASELINE.json configurations (device-timed, each with its own roofline)
    ctx = Ctx()
    ctx.torch, ctx.dist, ctx.L, ctx.ffi = torch, dist, L, _ffi
    ctx.rank, ctx.world, ctx.local_rank, ctx.stream, ctx.sptr = rank, world, local_rank, stream, sptr
    ctx.barrier, ctx.max_over_ranks, ctx.gather_floats, ctx.gen_threads = barrier, max_over_ranks, gather_floats, gen_threads
    sub = {}
    sub_steps = max(3, min(args.steps, 20))

    def run_sub(name, fn):
        t0 = time.perf_counter()
        try:
            r = fn()
        except Exception as e:   # a failing secondary configuration is reported, not hidden
            if dist is not None:
                raise            # ... but a rank that left a collective cannot be papered over
            r = {"error": "%s: %s" % (type(e).__name__, e)}
        r["bench_wall_s"] = round(time.perf_counter() - t0, 2)
        sub[name] = r

    if "gather" in configs:
        run_sub("c2_gather", lambda: run_c2_gather(ctx, corpus, batch, out_dev, n, sub_steps))
    if "4" in configs:
        run_sub("c4", lambda: run_c4(ctx, corpus, q, chars, offsets64, n, total, sub_steps))

    # ---- end to end through the C ABI with HOST buffers (pinned): query tables + chunked H2D / scan / D2H
    #      pipeline (rf_batch_stream_u32_off32), nothing kept on the GPU between steps
    e2e_steps = max(1, args.e2e_steps)
    m_tail = min(n, 200_000)
    step()                                      # re-score: the tail of this vector is compared with the streamed one below
    torch.cuda.synchronize()
    tail_dev = out_dev[n - m_tail:].clone()
    L.rf_corpus_destroy(corpus)   # keep peak device memory low
    corpus = None
    del out_dev

    lens8_t = torch.empty(n, dtype=torch.uint8).pin_memory()
    lens8 = lens8_t.numpy()
    np.copyto(lens8, np.diff(offsets64.view(np.int64)), casting="unsafe")
    out8_t = torch.empty(n, dtype=torch.uint8).pin_memory()
    out8 = out8_t.numpy()

    def e2e_step_csr():
        b2 = create_batch()
        _ffi.check(L.rf_batch_stream_u32_off32(b2, chars.ctypes.data, offsets32.ctypes.data, n, _ffi.KINDS["distance"],
                                               None, out_host.ctypes.data))
        L.rf_batch_destroy(b2)

    def e2e_step_len8():   # one length byte per candidate in, one score byte per candidate out (query len 32, candidates <= 64)
        b2 = create_batch()
        _ffi.check(L.rf_batch_stream_u8_len8(b2, chars.ctypes.data, lens8.ctypes.data, n, _ffi.KINDS["distance"],
                                             None, out8.ctypes.data))
        L.rf_batch_destroy(b2)

    # the 62-symbol corpus packed to 6 bits per character ONCE on the host (rf_pack6_u8, like packing the strings into CSR in
    # the first place: not part of a step); a step sends the packed stream and unpacks each chunk on the device
    import rapidfuzz_b200 as rf
    t_pack = time.perf_counter()
    packed6, dict64 = rf.pack6(chars, nthreads=gen_threads, pinned=True)
    t_pack = time.perf_counter() - t_pack
    out8p_t = torch.empty(n, dtype=torch.uint8).pin_memory()
    out8p = out8p_t.numpy()

    def e2e_step():
        b2 = create_batch()
        _ffi.check(L.rf_batch_stream_u8_len8_packed6(b2, packed6.ctypes.data, dict64.ctypes.data, lens8.ctypes.data, n,
                                                     _ffi.KINDS["distance"], None, out8p.ctypes.data))
        L.rf_batch_destroy(b2)

    def wall(fn):
        fn()   # warm-up: allocates the per-device chunk buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        barrier()
        return (time.perf_counter() - t0) / e2e_steps
    e2e_csr_s = wall(e2e_step_csr)
    e2e_len8_s = wall(e2e_step_len8)
    e2e_s = wall(e2e_step)
    tables = 2 * 256 * 4 + 2 * 256 * 8 + 256 * 8
    h2d_csr, d2h_csr = int(total + 4 * (n + 1) + tables), int(4 * n)
    h2d_len8, d2h_len8 = int(total + n + tables), int(n)
    h2d, d2h = int((total + 3) // 4 * 3 + n + tables), int(n)
    if ok is not None:
        from oracle import oracle as orc
        m = min(n, 200_000)
        exp = orc.batch("levenshtein", "distance", q, chars[: int(offsets64[m])], offsets64[: m + 1], nthreads=0)
        ok = ok and bool(np.array_equal(out_host[:m], exp))
        ok = ok and bool(np.array_equal(out_host[n - m_tail:], tail_dev.cpu().numpy().view(np.uint32)))
        ok = ok and bool(np.array_equal(out8, out_host.astype(np.uint8)) and int(out_host.max()) <= 254)   # byte results == u32 results, all n
        ok = ok and bool(np.array_equal(out8p, out8))                                                        # packed input == plain input, all n
    # secondary: upload + build a RESIDENT corpus (CSR + interleaved layout), score once, download, destroy
    out_host[:] = 0
    barrier()
    t0 = time.perf_counter()
    c2 = create_corpus()
    b2 = create_batch()
    _ffi.check(L.rf_batch_distance_u32(b2, c2, None, out_host.ctypes.data))
    L.rf_batch_destroy(b2)
    L.rf_corpus_destroy(c2)
    barrier()
    resident_s = time.perf_counter() - t0

    if "3" in configs:
        run_sub("c3", lambda: run_c3(ctx, sub_steps))
    if "5" in configs:
        run_sub("c5", lambda: run_c5(ctx, args.c5_queries, args.c5_candidates, max(2, min(args.steps, 3))))

    # ---- aggregate over ranks (device time: max over ranks; pairs: sum over ranks)
    if dist is not None:
        t = torch.tensor([ms, e2e_s, resident_s, e2e_csr_s, e2e_len8_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, resident_s, e2e_csr_s, e2e_len8_s = float(t[0]), float(t[1]), float(t[2]), float(t[3]), float(t[4])
        cnt = torch.tensor([n, total], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        n_all = int(cnt[0])
    else:
        n_all = n

    if rank == 0:
        ms_per_step = ms / args.steps
        value = n_all / (ms_per_step * 1e-3)
        # algorithmic bytes of ONE launch (one GPU's shard): candidate bytes + 4 B offset + 4 B result each
        alg_bytes = total + 8 * n
        traffic = ncu_traffic("scan_lb_traffic.json")
        if traffic is not None:  # the capture was taken at `candidates` per launch; traffic is linear in n
            traffic = traffic[0] * (n / traffic[1])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": config2_dict(n),
            "run": {"mean_len": total / n, "host_gen_s": round(t_gen, 2), "results_match_oracle_sample": ok, "numa_rank0": numa},
            "clocks": clk.summary(),
            "e2e": {"value": n_all / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                    "h2d_gbs": h2d / e2e_s / 1e9,
                    "what": "rf_batch_create_u8 + rf_batch_stream_u8_len8_packed6 (pinned host buffers: the candidates' characters "
                            "6-bit packed once by rf_pack6_u8 + one length byte per candidate -> chunked H2D / unpack / prefix sum / "
                            "scan / narrow / D2H pipeline -> pinned host byte scores) + rf_batch_destroy, per step; PCIe-bound",
                    "one_time_host_pack6_s": round(t_pack, 2),
                    "len8": {"value": n_all / e2e_len8_s, "ms_per_step": e2e_len8_s * 1e3, "h2d_bytes_per_step": h2d_len8,
                             "d2h_bytes_per_step": d2h_len8, "h2d_gbs": h2d_len8 / e2e_len8_s / 1e9,
                             "what": "the same through rf_batch_stream_u8_len8: plain bytes + one length byte in, one score byte out"},
                    "csr_u32": {"value": n_all / e2e_csr_s, "ms_per_step": e2e_csr_s * 1e3, "h2d_bytes_per_step": h2d_csr,
                                "d2h_bytes_per_step": d2h_csr, "h2d_gbs": h2d_csr / e2e_csr_s / 1e9,
                                "what": "the same through rf_batch_stream_u32_off32: u32 CSR starts in, u32 scores out (round 1's e2e)"},
                    "resident_corpus_build_and_score_ms": resident_s * 1e3},
            "gpu_launches": launches,
            "roofline": roofline(alg_bytes, ms_per_step, "scan_lb_kernel<F_LEV,u32,256,RAWDIST>", traffic,
                                 "ALU-pipe bound by design (7 LOP3 per candidate char on a 16-lane/clk pipe): the inner step alone "
                                 "runs at 18.4 SM clocks per warp-character (1.78 ms per 1e8 candidates, profiles/r2_step32_ubench.txt), "
                                 "the kernel at 89 % of that; traffic = ncu DRAM bytes of one launch; see DESIGN.md section 5"),
            "configs": sub,
        }
        if not args.no_cpu_baseline:
            if ORIG_AFFINITY:   # the CPU leg gets every core this job may use, not only the GPU-local ones
                os.sched_setaffinity(0, ORIG_AFFINITY)
            threads = host_threads()
            n_sample = min(n, 2_000_000 * max(1, min(threads, 32)))
            rate, secs = cpu_oracle_rate(n_sample, threads)
            rate1, _ = cpu_oracle_rate(min(n, 2_000_000), 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "first %d candidates of the same workload, %.2f s wall" % (n_sample, secs),
                                    "single_thread_value": rate1}
        print(json.dumps(line))
    L.rf_batch_destroy(batch)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
