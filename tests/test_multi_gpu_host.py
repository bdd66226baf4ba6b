"""CPU test of the N>1 host logic (world_size 2, gloo): cell-balanced sharding of the pair list and the gather of
result records to rank 0.  The per-rank alignment itself is stood in for by the oracle (test infrastructure);
on the GPU box the same code path runs with backend nccl inside bench.py."""
import importlib.util
import os
import socket
import sys

import numpy as np
import pytest

import common as C


def _mod(name):
    spec = importlib.util.spec_from_file_location("ssw_b200_" + name, os.path.join(C.PKG, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_shard_bounds_balance_cells():
    D = _mod("ssw_dist")
    cells = [150 * 5000] * 10 + [10000 * 5000] * 2 + [150 * 5000] * 10
    b = D.shard_bounds(cells, 4)
    assert b[0] == 0 and b[-1] == len(cells) and all(b[i] <= b[i + 1] for i in range(4))
    assert D.shard_bounds([], 3) == [0, 0, 0, 0]
    tot = [sum(cells[b[i]:b[i + 1]]) for i in range(4)]
    assert max(tot) <= 2.2 * (sum(cells) / 4)          # two huge pairs dominate; no rank gets both


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(C.ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = _mod("ssw_dist")
    L = _mod("ssw_lib")
    ref, reads = C.make_dna_workload(3000, 13, 60, seed_ref=5, seed_reads=6)
    reads[3] = reads[3][:25]                              # ragged
    mat = C.dna_matrix(2, 2)
    cells = [len(q) * len(ref) for q in reads]
    lo, hi = D.shard_range(cells, rank, world)
    oracle = C.load_oracle()
    local = np.zeros(hi - lo, dtype=L.RESULT_DTYPE)
    for i in range(lo, hi):
        r = oracle.align(reads[i], ref, mat, 5, 3, 1, 8, 0, 0, 30, 2)
        for f in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "flag"):
            local[i - lo][f] = r[f]
        local[i - lo]["cigar_off"] = -1
    merged = D.gather_records(local, rank, world)
    if rank == 0:
        full = []
        for i in range(len(reads)):
            r = oracle.align(reads[i], ref, mat, 5, 3, 1, 8, 0, 0, 30, 2)
            full.append(tuple(r[f] for f in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2")))
        got = [tuple(int(m[f]) for f in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2")) for m in merged]
        q.put(got == full and len(merged) == len(reads))
    dist.destroy_process_group()


def test_gather_to_rank0_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
