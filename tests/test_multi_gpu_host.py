"""CPU tests of the N>1 host logic (world_size 2, gloo): cell-balanced sharding of the pair list and the gather of
result records (and, two-phase, of the variable-length CIGAR words) to rank 0.  In the first test the per-rank
alignment is stood in for by the oracle; in the second every rank runs the PRODUCT's engine and kernel sources
(emulator build, tests/cuda_emu) on its shard, as bench.py does with backend nccl on the GPU box."""
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


def _worker_emu(rank, world, port, q):
    """every rank: product engine (emulator build) on its cell-balanced shard, flag 2 (CIGARs) -> gather_batch"""
    import subprocess
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(C.ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = _mod("ssw_dist")
    L = _mod("ssw_lib")
    emu_dir = os.path.join(C.ROOT, "tests", "cuda_emu")
    ref, reads = C.make_dna_workload(2500, 11, 70, seed_ref=15, seed_reads=16, p_ins=0.03, p_del=0.03)
    reads[2] = reads[2][:31]
    reads[7] = np.random.default_rng(3).integers(0, 4, size=70, dtype=np.int8)       # decoy: tiny or no CIGAR
    mat = C.dna_matrix(2, 2)
    cells = [len(x) * len(ref) for x in reads]
    lo, hi = D.shard_range(cells, rank, world)
    eng = L.BatchAligner(lib_dir=emu_dir, lib_name="libssw_emu.so")
    eng.set_sequences(reads[lo:hi], [ref])
    res, pool = eng.align(mat, 5, 3, 1, flag=2, filters=0, filterd=32767, mask_len=35, score_size=2)
    eng.close()
    recs, words = D.gather_batch(res, pool, rank, world)
    ok = None
    if rank == 0:
        exp, exp_pool, _, _, _ = C.cpu_batch(reads, [ref], np.arange(len(reads)), np.zeros(len(reads)), mat, 5, 3, 1, flag=2,
                                            filters=0, filterd=32767, mask_len=35, score_size=2, impl="port", threads=2)
        ok = len(recs) == len(reads) and C.compare_records(recs, words, exp, exp_pool) == [] and int((recs["cigar_len"] > 0).sum()) >= 8
        q.put(bool(ok))
    else:
        assert recs is None and words is None
    dist.destroy_process_group()


def test_product_engine_per_rank_with_cigar_gather_world2_gloo():
    import subprocess
    import torch.multiprocessing as mp
    subprocess.run(["make", "-s", "-C", os.path.join(C.ROOT, "tests", "cuda_emu")], check=True)
    C.build_oracle()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_emu, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=400)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
