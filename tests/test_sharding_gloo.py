"""Host-side logic of the N>1 path on CPU: world_size-2 gloo process group.  Each rank scores its shard (here
with the CPU oracle standing in for the GPU kernels -- this test is about sharding, gathering and top-k
merging), all-gathers, and the result must equal the single-process answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rapidfuzz_b200 as rf
import synth
from rapidfuzz_b200 import sharding
from oracle import oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, chars, offsets, qs, k, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lchars, loffs, lo = sharding.local_shard(chars, offsets, world, rank)
        local = orc.batch("levenshtein", "distance", q, lchars, loffs).astype(np.int64)
        full = sharding.all_gather_scores(torch.from_numpy(local))
        # per-shard top-k of a few queries, merged after an all-gather
        n_local = len(loffs) - 1
        idx = np.full((len(qs), k), 0xFFFFFFFF, dtype=np.uint32)
        dd = np.full((len(qs), k), 0xFFFFFFFF, dtype=np.uint32)
        for qi, qq in enumerate(qs):
            d = orc.batch("levenshtein", "distance", qq, lchars, loffs).astype(np.int64)
            keys = np.sort(d * (1 << 32) + np.arange(n_local))[:k]
            idx[qi, :len(keys)] = keys & 0xFFFFFFFF
            dd[qi, :len(keys)] = keys >> 32
        gi, gd = sharding.all_gather_topk(torch.from_numpy(idx.astype(np.int64)), torch.from_numpy(dd.astype(np.int64)), lo, k)
        ret[rank] = (full.numpy(), gi, gd, lo, n_local)
    finally:
        dist.destroy_process_group()


def test_shard_range_balances_bytes():
    rng = np.random.default_rng(1)
    lens = np.concatenate([rng.integers(1, 10, 1000), rng.integers(500, 1000, 100)])
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for world in (1, 2, 4, 8):
        spans = [sharding.shard_range(offsets, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == len(lens)
        for a, b in zip(spans[:-1], spans[1:]):
            assert a[1] == b[0]
        byts = [int(offsets[hi] - offsets[lo]) for lo, hi in spans]
        assert max(byts) - min(byts) <= 2 * lens.max()
    assert sharding.shard_range(np.zeros(1, np.uint64), 4, 2) == (0, 0)     # empty corpus
    assert sharding.shard_range(np.zeros(9, np.uint64), 4, 3) == (6, 8)     # all-empty candidates: by count


def test_merge_topk_pads_and_orders():
    i0 = np.array([[5, 7, 0xFFFFFFFF]], dtype=np.uint32)
    d0 = np.array([[1, 3, 0xFFFFFFFF]], dtype=np.uint32)
    i1 = np.array([[2, 0xFFFFFFFF, 0xFFFFFFFF]], dtype=np.uint32)
    d1 = np.array([[1, 0xFFFFFFFF, 0xFFFFFFFF]], dtype=np.uint32)
    gi, gd = sharding.merge_topk([i0, i1], [d0, d1], [0, 100], 4)
    assert gi.dtype == np.uint64 and gd.dtype == np.uint32
    assert gi.tolist() == [[5, 102, 7, 0xFFFFFFFFFFFFFFFF]] and gd.tolist() == [[1, 1, 3, 0xFFFFFFFF]]
    # shard starts beyond 2^32 must not spill into the distance (the old 32-bit packed key did)
    gi, gd = sharding.merge_topk([i0, i1], [d0, d1], [0, 2**32 + 100], 4)
    assert gi.tolist() == [[5, 2**32 + 102, 7, 0xFFFFFFFFFFFFFFFF]] and gd.tolist() == [[1, 1, 3, 0xFFFFFFFF]]


@pytest.mark.timeout(300)
def test_world2_gloo_gather_equals_single_process():
    q = synth.synth_query(2, 32)
    chars, offsets = synth.synth_corpus(2, q, 5001, 8, 64, 16)
    qs = [synth.synth_query(100 + i, 32) for i in range(3)]
    k, world = 10, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), q, chars, offsets, qs, k, ret), nprocs=world, join=True)
    exp = orc.batch("levenshtein", "distance", q, chars, offsets).astype(np.int64)
    n = len(offsets) - 1
    assert ret[0][3] == 0 and ret[0][4] + ret[1][4] == n and ret[1][3] == ret[0][4]
    for r in range(world):
        assert np.array_equal(ret[r][0], exp)
    for qi, qq in enumerate(qs):
        d = orc.batch("levenshtein", "distance", qq, chars, offsets).astype(np.int64)
        keys = np.sort(d * (1 << 32) + np.arange(n))[:k]
        for r in range(world):
            assert np.array_equal(ret[r][1][qi], (keys & 0xFFFFFFFF).astype(np.uint64))
            assert np.array_equal(ret[r][2][qi], (keys >> 32).astype(np.uint32))
