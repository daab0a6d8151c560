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
    top-k by (distance, global index)."""
    NONE = np.uint64(0xFFFFFFFF)
    keys = []
    for idx, dist, lo in zip(idx_parts, dist_parts, shard_starts):
        idx = np.asarray(idx).astype(np.uint64)
        dist = np.asarray(dist).astype(np.uint64)
        valid = idx != NONE
        key = (dist << np.uint64(32)) | (idx + np.uint64(lo))
        keys.append(np.where(valid, key, np.uint64(0xFFFFFFFFFFFFFFFF)))
    allk = np.sort(np.concatenate(keys, axis=1), axis=1)[:, :k]
    none = allk == np.uint64(0xFFFFFFFFFFFFFFFF)
    out_idx = np.where(none, NONE, allk & NONE).astype(np.uint32)
    out_dist = np.where(none, NONE, allk >> np.uint64(32)).astype(np.uint32)
    return out_idx, out_dist


def all_gather_topk(idx_local, dist_local, shard_start, k, group=None):
    """All-gather of the per-shard [nq,k] (index, distance) lists followed by a local merge."""
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
