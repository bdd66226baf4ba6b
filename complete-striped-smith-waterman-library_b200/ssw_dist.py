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
    item = local.dtype.itemsize
    dev = torch.device(device) if device is not None else torch.device("cpu")
    n_local = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = np.zeros(cap * item, dtype=np.uint8)
    buf[: len(local) * item] = np.frombuffer(local.tobytes(), dtype=np.uint8)
    t = torch.from_numpy(buf).to(dev)
    out = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, out, dst=0)
    if rank != 0:
        return None
    parts = [np.frombuffer(o.cpu().numpy().tobytes()[: sizes[r] * item], dtype=local.dtype) for r, o in enumerate(out)]
    return np.concatenate(parts) if parts else local[:0]
