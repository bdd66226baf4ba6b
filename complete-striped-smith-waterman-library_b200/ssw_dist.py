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


def gather_batch(records, pool, rank, world, device=None):
    """Gather a rank's result records AND its CIGAR pool to rank 0 (lengths first, then the payloads).
    Phase 1: every rank announces (records, CIGAR words).  Phase 2: one gather of the fixed-size records and one of
    the variable-length CIGAR words (padded to the longest shard).  On rank 0 the records come back in rank order with
    `cigar_off` re-based into the concatenated pool; other ranks get (None, None)."""
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
    cap_r = max(max(s[0] for s in sizes), 1) * item
    cap_p = max(max(s[1] for s in sizes), 1)
    rb = np.zeros(cap_r, dtype=np.uint8)
    rb[: len(records) * item] = records.view(np.uint8).reshape(-1) if len(records) else rb[:0]
    pb = np.zeros(cap_p, dtype=np.int32)
    pb[: len(pool)] = np.asarray(pool, dtype=np.uint32).view(np.int32)
    tr, tp = torch.from_numpy(rb).to(dev), torch.from_numpy(pb).to(dev)
    out_r = [torch.empty_like(tr) for _ in range(world)] if rank == 0 else None
    out_p = [torch.empty_like(tp) for _ in range(world)] if rank == 0 else None
    dist.gather(tr, out_r, dst=0)                                           # phase 2: payloads
    any_words = any(s[1] > 0 for s in sizes)
    if any_words:
        dist.gather(tp, out_p, dst=0)
    if rank != 0:
        return None, None
    recs, pools, base = [], [], 0
    for r in range(world):
        n_r, n_p = sizes[r]
        part = out_r[r].cpu().numpy()[: n_r * item].view(records.dtype).copy()
        if n_p:
            has = part["cigar_off"] >= 0
            part["cigar_off"][has] += base
            pools.append(out_p[r].cpu().numpy()[:n_p].view(np.uint32))
        base += n_p
        recs.append(part)
    return np.concatenate(recs), (np.concatenate(pools) if pools else np.zeros(0, dtype=np.uint32))


def split_even(n, rank, world):
    """Contiguous block [lo, hi) of n equal-cost units (e.g. the queries of a full queries x targets grid) for `rank`."""
    return n * rank // world, n * (rank + 1) // world
