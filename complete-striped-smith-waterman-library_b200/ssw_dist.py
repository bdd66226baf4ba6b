"""Multi-GPU plumbing of the batched aligner: independent (query, reference) pairs are the unit of data
parallelism (the reference's CLI loops over them serially, main.c:462-532).  Ranks take contiguous blocks of the
pair list balanced by DP cells, every rank aligns its block on its own GPU, and the fixed-size result records
are gathered to rank 0 -- one collective per batch, nothing is exchanged while the matrices are filled.
Works with any torch.distributed backend: "nccl" on the GPUs, "gloo" in the CPU tests."""
import numpy as np


def shard_bounds(cells, world):
    """Split a list of per-pair cell counts into `world` contiguous blocks of nearly equal total cells.
    Returns world+1 boundaries."""
    cells = np.asarray(cells, dtype=np.float64)
    n = len(cells)
    if n == 0:
        return [0] * (world + 1)
    csum = np.concatenate([[0.0], np.cumsum(cells)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(csum, target, side="left"))
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return bounds


def shard_range(cells, rank, world):
    b = shard_bounds(cells, world)
    return b[rank], b[rank + 1]


def gather_records(local, rank, world, device=None):
    """Gather structured numpy records (e.g. ssw_lib.RESULT_DTYPE) from all ranks to rank 0, in rank order.
    Shards may have different lengths.  Returns the concatenated array on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    recs, _ = gather_batch(local, np.zeros(0, dtype=np.uint32), rank, world, device)
    return recs


_pinned = {}


def _pinned_host(nbytes, key):
    """A cached pinned host buffer (page-locked allocations cost ~0.3 s per GB: made once per size class, not per step)."""
    import torch
    t = _pinned.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(int(nbytes), dtype=torch.uint8, pin_memory=True)
        _pinned[key] = t
    return t


def _gather_bytes(payload_u8, sizes_bytes, rank, world, dev, key):
    """Gather variable-length byte strings to rank 0.  Device side: every rank copies its bytes into a buffer padded to the
    longest shard, one dist.gather fills the rows of a [world, cap] tensor on rank 0, one copy brings it to a cached pinned
    host buffer.  Returns (host uint8 array [world, cap], cap) on rank 0, (None, cap) elsewhere."""
    import torch
    import torch.distributed as dist
    cap = max(max(sizes_bytes), 16)
    cap = (cap + 15) // 16 * 16
    on_gpu = dev.type == "cuda"
    mine = torch.empty(cap, dtype=torch.uint8, device=dev)
    n = int(payload_u8.size)
    if n:
        mine[:n].copy_(torch.from_numpy(payload_u8))
    if rank == 0:
        out = torch.empty((world, cap), dtype=torch.uint8, device=dev)
        dist.gather(mine, list(out.unbind(0)), dst=0)
        if on_gpu:
            host = _pinned_host(world * cap, key)[: world * cap].view(world, cap)
            host.copy_(out)
            return host.numpy(), cap
        return out.numpy(), cap
    dist.gather(mine, None, dst=0)
    return None, cap


def gather_batch(records, pool, rank, world, device=None):
    """Gather a rank's result records AND its CIGAR pool to rank 0 (lengths first, then the payloads).
    Phase 1: every rank announces (records, CIGAR words).  Phase 2: one gather of the fixed-size records and one of
    the variable-length CIGAR words (padded to the longest shard).  On rank 0 the records come back in rank order with
    `cigar_off` re-based into the concatenated pool; other ranks get (None, None).  The arrays returned on rank 0 may be
    views of a cached pinned buffer: they are valid until the next gather of the same kind."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return records, pool
    dev = torch.device(device) if device is not None else torch.device("cpu")
    item = records.dtype.itemsize
    mine = torch.tensor([len(records), len(pool)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, mine)                                            # phase 1: lengths
    sizes = [[int(x) for x in s.tolist()] for s in sizes]
    rec_u8 = np.ascontiguousarray(records).view(np.uint8).reshape(-1) if len(records) else np.zeros(0, dtype=np.uint8)
    host_r, cap_r = _gather_bytes(rec_u8, [s[0] * item for s in sizes], rank, world, dev, "records")     # phase 2: payloads
    any_words = any(s[1] > 0 for s in sizes)
    host_p = None
    if any_words:
        pool_u8 = np.ascontiguousarray(np.asarray(pool, dtype=np.uint32)).view(np.uint8).reshape(-1)
        host_p, cap_p = _gather_bytes(pool_u8, [s[1] * 4 for s in sizes], rank, world, dev, "cigars")
    if rank != 0:
        return None, None
    if all(s[0] * item == cap_r for s in sizes):
        recs = host_r.reshape(-1).view(records.dtype)                        # equal shards: no compaction copy
    else:
        recs = np.concatenate([host_r[r, : sizes[r][0] * item].view(records.dtype) for r in range(world)])
    pools, base = [], 0
    if any_words:
        lo = 0
        for r in range(world):
            n_r, n_p = sizes[r]
            if n_p:
                part = recs[lo: lo + n_r]
                has = part["cigar_off"] >= 0
                part["cigar_off"][has] += base
                pools.append(host_p[r, : n_p * 4].view(np.uint32))
            base += n_p
            lo += n_r
    return recs, (np.concatenate(pools) if pools else np.zeros(0, dtype=np.uint32))


def split_even(n, rank, world):
    """Contiguous block [lo, hi) of n equal-cost units (e.g. the queries of a full queries x targets grid) for `rank`."""
    return n * rank // world, n * (rank + 1) // world
