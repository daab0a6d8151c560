#!/usr/bin/env python3
"""What the host side of the PCIe links delivers: pinned H2D (and concurrent D2H) copy rates of N GPUs, each alone and
all at once, with the pinned buffers (a) wherever the default policy puts them and (b) on the NUMA node of the GPU
(set_mempolicy before the allocation).  This is the ceiling of the end-to-end (`e2e`) number of bench.py at N GPUs.

  python tools/pcie_peak.py --gpus 8 [--mb 1024] [--reps 6]

One child process per GPU (spawned here, no torchrun needed), synchronised with barriers.  Prints one JSON line."""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_numa_node(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        return int(open("/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower())).read().strip())
    except Exception:
        return -1


def set_mempolicy(mode, node):
    libc = C.CDLL(None, use_errno=True)
    if mode == 0:
        rc = libc.syscall(238, 0, None, 0)
    else:
        mask = C.c_ulong(1 << node)
        rc = libc.syscall(238, mode, C.byref(mask), 64)
    return 0 if rc == 0 else C.get_errno()


def worker(rank, n, mb, reps, barrier, q):
    import torch
    torch.cuda.set_device(rank)
    node = gpu_numa_node(rank)
    res = {"rank": rank, "gpu_numa_node": node, "cpus_allowed": len(os.sched_getaffinity(0))}
    nbytes = mb << 20
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev2 = torch.empty(nbytes // 8, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def rate(h2d, d2h, solo_rank=None):
        """GB/s of `reps` H2D copies (with a concurrent D2H stream of 1/8 the size when d2h) of THIS rank; ranks other than
        solo_rank idle when it is set."""
        barrier.wait()
        if solo_rank is not None and rank != solo_rank:
            barrier.wait()
            return None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s1):
                dev.copy_(h2d, non_blocking=True)
            if d2h is not None:
                with torch.cuda.stream(s2):
                    d2h.copy_(dev2, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier.wait()
        return nbytes * reps / dt / 1e9

    for label, mode in (("default", 0), ("gpu_node", 1)):
        if mode and node >= 0:
            res["set_mempolicy_errno"] = set_mempolicy(1, node)   # MPOL_PREFERRED
        host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        host.fill_(1)
        hout = torch.empty(nbytes // 8, dtype=torch.uint8).pin_memory()
        rate(host, None)  # warm
        res[label + "_all_h2d"] = rate(host, None)
        res[label + "_all_h2d_with_d2h"] = rate(host, hout)
        solo = []
        for r in range(n):
            v = rate(host, None, solo_rank=r)
            if v is not None:
                solo.append(v)
        res[label + "_alone_h2d"] = solo[0] if solo else None
        del host, hout
        if mode:
            set_mempolicy(0, 0)
    q.put(res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=6)
    a = ap.parse_args()
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(a.gpus)
    q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, a.gpus, a.mb, a.reps, barrier, q)) for r in range(a.gpus)]
    for p in ps:
        p.start()
    res = sorted((q.get() for _ in ps), key=lambda r: r["rank"])
    for p in ps:
        p.join()
    nodes = {}
    try:
        for d in sorted(os.listdir("/sys/devices/system/node")):
            if d.startswith("node") and d[4:].isdigit():
                nodes[d] = open("/sys/devices/system/node/%s/cpulist" % d).read().strip()
    except Exception:
        pass
    out = {"what": "pinned host -> device copy rate per GPU in GB/s (%d MiB x %d), ranks concurrently (`all`) and one at a time (`alone`); "
                   "`gpu_node` = buffers allocated under set_mempolicy(MPOL_PREFERRED, the GPU's NUMA node)" % (a.mb, a.reps),
           "n_gpus": a.gpus, "numa_nodes": nodes, "per_rank": res}
    for label in ("default", "gpu_node"):
        for k in ("_all_h2d", "_all_h2d_with_d2h", "_alone_h2d"):
            vals = [r.get(label + k) for r in res if r.get(label + k) is not None]
            out[label + k + "_sum"] = sum(vals) if vals else None
    print(json.dumps(out))


if __name__ == "__main__":
    main()
