"""Multi-GPU plumbing for the one-vs-many / many-vs-many path: the candidate corpus shards by candidate
index across the ranks of one node (one process per GPU, torch.distributed); the query and its 2-8 KB match
table are replicated; the only collective is the final all-gather of the score vector (or of the per-shard
top-k lists, merged locally).  Nothing here is on the scan's critical path.

The reference has no counterpart (single-threaded library); this is new design."""
import numpy as np


def shard_range(offsets, world_size, rank):
    """Contiguous candidate range [lo, hi) of `rank`, balanced by bytes (sum of lengths), not by count
    (SURVEY section 8e).  offsets: CSR starts (n+1)."""
    offsets = np.asarray(offsets)
    n = len(offsets) - 1
    total = int(offsets[n])
    if world_size <= 1:
        return 0, n
    # boundary b = first candidate whose start offset >= total * b / world_size; ties by count for empty corpora
    bounds = [0]
    for b in range(1, world_size):
        if total == 0:
            bounds.append(n * b // world_size)
        else:
            bounds.append(int(np.searchsorted(offsets[:n], total * b // world_size, side="left")))
    bounds.append(n)
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds[rank], bounds[rank + 1]


def local_shard(chars, offsets, world_size, rank):
    """(chars, offsets rebased to 0, lo) of this rank's shard."""
    lo, hi = shard_range(offsets, world_size, rank)
    offsets = np.asarray(offsets)
    c0, c1 = int(offsets[lo]), int(offsets[hi])
    return np.asarray(chars)[c0:c1], (offsets[lo:hi + 1] - offsets[lo]).astype(np.uint64), lo


def all_gather_scores(local_scores, group=None):
    """All-gather of per-shard score vectors of unequal length (NCCL on GPUs, gloo on CPU): shards are
    padded to the longest, gathered with ONE all_gather, and trimmed.  Returns the full vector in
    candidate order on every rank (same device/dtype as the input tensor)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_local = torch.tensor([local_scores.numel()], dtype=torch.int64, device=local_scores.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    padded = torch.zeros(m, dtype=local_scores.dtype, device=local_scores.device)
    padded[: local_scores.numel()] = local_scores
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)])


def merge_topk(idx_parts, dist_parts, shard_starts, k):
    """Merge per-shard top-k lists ([nq,k] each, shard-local indices, 0xFFFFFFFF padding) into the global
    top-k by (distance, global index).  Returns (idx uint64 [nq,k] GLOBAL indices, UINT64_MAX = none; dist uint32,
    0xFFFFFFFF = none) -- the same contract as rf_topk_merge_device, valid for sharded corpora of 2^32 candidates and more."""
    NONE32, NONE64 = np.uint32(0xFFFFFFFF), np.uint64(0xFFFFFFFFFFFFFFFF)
    gi, gd = [], []
    for idx, dist, lo in zip(idx_parts, dist_parts, shard_starts):
        idx = np.asarray(idx).astype(np.uint32)
        valid = idx != NONE32
        gi.append(np.where(valid, idx.astype(np.uint64) + np.uint64(lo), NONE64))
        gd.append(np.where(valid, np.asarray(dist).astype(np.uint32), NONE32))
    gi, gd = np.concatenate(gi, axis=1), np.concatenate(gd, axis=1)
    order = np.lexsort((gi, gd), axis=1)[:, :k]      # primary key distance, ties by global index; padding sorts last
    return np.take_along_axis(gi, order, axis=1), np.take_along_axis(gd, order, axis=1)


def cdist_topk_device(q_chars, q_offsets, corpus, k=10, score_cutoff=None, device=None):
    """rf_cdist_topk_u8_device on this rank's shard: (idx, dist) int32 CUDA tensors [nq,k] (shard-local indices,
    -1 = none), left on the device for the gather.  q_chars u8 / q_offsets u64: host CSR of the queries."""
    import ctypes as C
    import torch
    from . import _ffi
    q_chars = np.ascontiguousarray(q_chars, dtype=np.uint8)
    q_offsets = np.ascontiguousarray(q_offsets, dtype=np.uint64)
    nq = len(q_offsets) - 1
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
    dist = torch.empty((nq, k), dtype=torch.int32, device=dev)
    a = _ffi.RfArgs()
    _ffi.lib().rf_args_default(C.byref(a))
    if score_cutoff is not None:
        a.has_cutoff = 1
        a.cutoff_u = int(score_cutoff)
    _ffi.check(_ffi.lib().rf_cdist_topk_u8_device(q_chars.ctypes.data, q_offsets.ctypes.data, nq, corpus._h, C.byref(a), k,
                                                 idx.data_ptr(), dist.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return idx, dist


def merge_topk_device(parts_idx_dist, starts, k, stream=None):
    """Device-side merge (rf_topk_merge_device): parts_idx_dist int32 CUDA tensor [parts, 2, nq, k] (what one
    all_gather_into_tensor of the stacked per-shard (idx, dist) lists produces), starts int64 CUDA tensor [parts].
    Returns (idx int64 [nq,k] global indices, -1 = none; dist int32 [nq,k], -1 = none) on the device."""
    import torch
    from . import _ffi
    parts, two, nq, kk = parts_idx_dist.shape
    assert two == 2 and kk == k and parts_idx_dist.is_contiguous() and parts_idx_dist.dtype == torch.int32
    dev = parts_idx_dist.device
    out_idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
    out_dist = torch.empty((nq, k), dtype=torch.int32, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    base = parts_idx_dist.data_ptr()
    _ffi.check(_ffi.lib().rf_topk_merge_device(base, base + 4 * nq * k, 2 * nq * k, starts.data_ptr(), parts, nq, k,
                                               out_idx.data_ptr(), out_dist.data_ptr(), dev.index or 0, st))
    return out_idx, out_dist


def all_gather_topk_device(idx_local, dist_local, shard_start, k, group=None):
    """GPU path of all_gather_topk: ONE all-gather (NCCL over NVLink) of the stacked per-shard lists + one of the
    shard starts, merged on the device; nothing touches the host.  idx_local / dist_local: int32 CUDA [nq,k]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = idx_local.device
    packed = torch.stack([idx_local, dist_local], dim=0).contiguous()
    parts = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=dev)
    start = torch.tensor([shard_start], dtype=torch.int64, device=dev)
    starts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(parts, packed, group=group)
    dist.all_gather_into_tensor(starts, start, group=group)
    return merge_topk_device(parts, starts, k)


def all_gather_topk(idx_local, dist_local, shard_start, k, group=None):
    """All-gather of the per-shard [nq,k] (index, distance) lists followed by a local merge (host tensors / gloo:
    the plumbing test; on GPUs use all_gather_topk_device)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = idx_local.device
    packed = torch.stack([idx_local.to(torch.int64), dist_local.to(torch.int64)], dim=0).contiguous()
    start = torch.tensor([shard_start], dtype=torch.int64, device=dev)
    parts = [torch.empty_like(packed) for _ in range(world)]
    starts = [torch.empty_like(start) for _ in range(world)]
    dist.all_gather(parts, packed, group=group)
    dist.all_gather(starts, start, group=group)
    idx_parts = [p[0].cpu().numpy().astype(np.uint32) for p in parts]
    dist_parts = [p[1].cpu().numpy().astype(np.uint32) for p in parts]
    return merge_topk(idx_parts, dist_parts, [int(s.item()) for s in starts], k)


# ---------------------------------------------------------------------------------------------------------------------
# One process, several GPUs: mirrors of the C ABI's sharded handles (include/rfgpu.h, rf_sharded.cu).  Unlike the
# functions above (one process per GPU, torch.distributed), these need no PyTorch: the library owns the split, the
# per-device streams and the NCCL communicator.
class ShardedCorpus:
    """rf_corpus_create_sharded_u8: the candidates split by bytes over `devices` (one resident shard per entry)."""

    def __init__(self, chars, offsets, devices):
        import ctypes as C
        from . import _ffi
        chars = np.ascontiguousarray(chars, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.devices = [int(d) for d in devices]
        dv = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        _ffi.check(_ffi.lib().rf_corpus_create_sharded_u8(chars.ctypes.data, offsets.ctypes.data, len(offsets) - 1, dv,
                                                         len(self.devices), C.byref(h)))
        self._h = h

    def __len__(self):
        from . import _ffi
        return int(_ffi.lib().rf_sharded_corpus_size(self._h))

    @property
    def uses_nccl(self):
        from . import _ffi
        return bool(_ffi.lib().rf_sharded_corpus_uses_nccl(self._h))

    def shard_ranges(self):
        import ctypes as C
        from . import _ffi
        out = []
        for i in range(len(self.devices)):
            a, b = C.c_uint64(), C.c_uint64()
            _ffi.check(_ffi.lib().rf_sharded_corpus_shard_range(self._h, i, C.byref(a), C.byref(b)))
            out.append((a.value, b.value))
        return out

    def close(self):
        from . import _ffi
        if getattr(self, "_h", None):
            _ffi.lib().rf_sharded_corpus_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedBatchComparator:
    """rf_sharded_batch_create_*: one BatchComparator replicated on every device of the list."""

    def __init__(self, metric, query, devices):
        import ctypes as C
        from . import _ffi
        from ._scorer import _as_query
        q = _as_query(query)
        self.metric, self.query, self.devices = metric, q, [int(d) for d in devices]
        dv = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        fn = _ffi.lib().rf_sharded_batch_create_u32 if q.dtype == np.uint32 else _ffi.lib().rf_sharded_batch_create_u8
        _ffi.check(fn(_ffi.METRICS[metric], q.ctypes.data, len(q), dv, len(self.devices), C.byref(h)))
        self._h = h

    def _is_f(self, kind):
        from . import _ffi
        return bool(_ffi.lib().rf_result_is_float(_ffi.METRICS[self.metric], _ffi.KINDS[kind]))

    def score(self, kind, corpus, args=None):
        """Raw sentinel-carrying scores of the whole sharded corpus (host vector; no collective)."""
        import ctypes as C
        from . import _ffi
        from ._scorer import Args
        is_f = self._is_f(kind)
        ca = (args if args is not None else Args())._c(is_f)
        out = np.empty(len(corpus), dtype=np.float64 if is_f else np.uint32)
        fn = _ffi.lib().rf_sharded_score_f64 if is_f else _ffi.lib().rf_sharded_score_u32
        _ffi.check(fn(self._h, corpus._h, _ffi.KINDS[kind], C.byref(ca), out.ctypes.data))
        return out

    def score_allgather(self, kind, corpus, out_ptrs, args=None):
        """rf_sharded_score_*_allgather_device: out_ptrs[i] = device pointer of an n-element buffer on devices[i]."""
        import ctypes as C
        from . import _ffi
        from ._scorer import Args
        is_f = self._is_f(kind)
        ca = (args if args is not None else Args())._c(is_f)
        arr = (C.c_void_p * len(out_ptrs))(*[int(p) for p in out_ptrs])
        fn = _ffi.lib().rf_sharded_score_f64_allgather_device if is_f else _ffi.lib().rf_sharded_score_u32_allgather_device
        _ffi.check(fn(self._h, corpus._h, _ffi.KINDS[kind], C.byref(ca), arr))

    def extract(self, kind, corpus, k=5, args=None):
        import ctypes as C
        from . import _ffi
        from ._scorer import Args
        is_f = self._is_f(kind)
        ca = (args if args is not None else Args())._c(is_f)
        idx = np.empty(k, dtype=np.uint64)
        score = np.empty(k, dtype=np.float64 if is_f else np.uint32)
        m = C.c_uint32(0)
        fn = _ffi.lib().rf_sharded_extract_f64 if is_f else _ffi.lib().rf_sharded_extract_u32
        _ffi.check(fn(self._h, corpus._h, _ffi.KINDS[kind], C.byref(ca), k, idx.ctypes.data, score.ctypes.data, C.byref(m)))
        return idx[: m.value], score[: m.value]

    def stream(self, kind, chars, offsets, args=None, out=None):
        """rf_sharded_stream_*: host-resident candidates, split by bytes over the devices, one PCIe pipeline each."""
        import ctypes as C
        from . import _ffi
        from ._scorer import Args
        chars = np.ascontiguousarray(chars, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        is_f = self._is_f(kind)
        ca = (args if args is not None else Args())._c(is_f)
        if out is None:
            out = np.empty(n, dtype=np.float64 if is_f else np.uint32)
        fn = _ffi.lib().rf_sharded_stream_f64 if is_f else _ffi.lib().rf_sharded_stream_u32
        _ffi.check(fn(self._h, chars.ctypes.data, offsets.ctypes.data, n, _ffi.KINDS[kind], C.byref(ca), out.ctypes.data))
        return out

    def stream_len8(self, kind, chars, lens, args=None, out=None, u8_results=False, dict64=None):
        """rf_sharded_stream_{u32,u8}_len8[_packed6]: one length byte per candidate on the wire, chunks handed to the devices
        dynamically (no static split).  dict64 given: `chars` is the 6-bit packed stream of corpus.pack6 (u8 results only)."""
        import ctypes as C
        from . import _ffi
        from ._scorer import Args
        chars = np.ascontiguousarray(chars, dtype=np.uint8)
        lens = np.ascontiguousarray(lens, dtype=np.uint8)
        n = len(lens)
        ca = (args if args is not None else Args())._c(False)
        dt = np.uint8 if u8_results else np.uint32
        if out is None:
            out = np.empty(n, dtype=dt)
        assert out.dtype == dt and len(out) >= n and out.flags.c_contiguous
        L = _ffi.lib()
        if dict64 is not None:
            assert u8_results, "the packed form returns byte scores"
            d = np.ascontiguousarray(dict64, dtype=np.uint8)
            _ffi.check(L.rf_sharded_stream_u8_len8_packed6(self._h, chars.ctypes.data, d.ctypes.data, lens.ctypes.data, n,
                                                           _ffi.KINDS[kind], C.byref(ca), out.ctypes.data))
        else:
            fn = L.rf_sharded_stream_u8_len8 if u8_results else L.rf_sharded_stream_u32_len8
            _ffi.check(fn(self._h, chars.ctypes.data, lens.ctypes.data, n, _ffi.KINDS[kind], C.byref(ca), out.ctypes.data))
        return out

    def close(self):
        from . import _ffi
        if getattr(self, "_h", None):
            _ffi.lib().rf_sharded_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sharded_cdist_topk(q_chars, q_offsets, corpus, k=10, score_cutoff=None):
    """rf_sharded_cdist_topk_u8: (idx uint64 [nq,k] global indices, dist uint32 [nq,k]); padding = all ones."""
    import ctypes as C
    from . import _ffi
    q_chars = np.ascontiguousarray(q_chars, dtype=np.uint8)
    q_offsets = np.ascontiguousarray(q_offsets, dtype=np.uint64)
    nq = len(q_offsets) - 1
    a = _ffi.RfArgs()
    _ffi.lib().rf_args_default(C.byref(a))
    if score_cutoff is not None:
        a.has_cutoff = 1
        a.cutoff_u = int(score_cutoff)
    idx = np.empty((nq, k), dtype=np.uint64)
    dist = np.empty((nq, k), dtype=np.uint32)
    _ffi.check(_ffi.lib().rf_sharded_cdist_topk_u8(q_chars.ctypes.data, q_offsets.ctypes.data, nq, corpus._h, C.byref(a), k,
                                                  idx.ctypes.data, dist.ctypes.data))
    return idx, dist
