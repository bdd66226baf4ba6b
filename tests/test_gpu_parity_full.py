"""Parity at the sizes SURVEY 8(d) asks for, against the COMPILED reference (oracle/_ref/libssw_ref.so) run on all host
cores of the GPU box:
  config 2: all 1,000 reads x 5 Mbp, flag 0x0f (scores, ends, begins, flag, CIGAR) and flag 0;
  config 3: a fixed 2,000-read subsample of the 100,000-read set;
  config 4: a 200 x 1,000 sub-grid of the protein grid, word scores;
  config 5: 1,000 x 10 kbp long reads, flags 0x0f and 2, every field and every CIGAR word.
The reference side is CPU-bound (~1e12 cells per case, tens of seconds on the box's usable cores); EVERY pair of every case is
checked (oracle/ssw_harness.c), the coverage is printed to stdout and recorded in gpurun_out/parity_full.json."""
import json
import os
import time

import numpy as np
import pytest

import common as C
from test_gpu_parity import FIELDS8, batch_dict, engine  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

THREADS = C.effective_cores()[0]


_cache = {}


def _workload(ref_len, n_reads, read_len, seed_ref, seed_reads, **kw):
    key = (ref_len, n_reads, read_len, seed_ref, seed_reads, tuple(sorted(kw.items())))
    if key not in _cache:
        _cache[key] = C.make_dna_workload(ref_len, n_reads, read_len, seed_ref=seed_ref, seed_reads=seed_reads, **kw)
    return _cache[key]


def _record(name, info):
    path = os.path.join(C.ROOT, "gpurun_out", "parity_full.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = {}
    if os.path.exists(path):
        with open(path) as f:
            data = json.load(f)
    data[name] = info
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


_expected = {}


def check_all(name, capfd, res, pool, queries, refs, pair_q, pair_r, mat, n, expect_key=None, **kw):
    """EVERY pair is re-computed by the compiled reference on all usable host cores (pthread harness) and compared field by
    field and CIGAR word by CIGAR word; the coverage goes to stdout (so that it lands in the driver's test log) and to
    gpurun_out/parity_full.json."""
    t0 = time.time()
    if expect_key is None or expect_key not in _expected:
        _expected[expect_key] = C.cpu_batch(queries, refs, pair_q, pair_r, mat, n, impl="reference", threads=THREADS, **kw)
    exp, exp_pool, secs, cells, kind = _expected[expect_key]
    if expect_key is None:
        del _expected[None]
    bad = C.compare_records(res, pool, exp, exp_pool)
    info = {"pairs_total": int(len(pair_q)), "pairs_checked": int(len(exp)), "mismatches": len(bad), "threads": THREADS,
            "seconds": round(time.time() - t0, 1), "checker": kind, "checker_gcups": round(cells / secs / 1e9, 1)}
    _record(name, info)
    with capfd.disabled():
        print("\n[parity %s] %s" % (name, json.dumps(info)))
    assert len(exp) == len(pair_q) == len(res)
    assert bad == [], (bad[:3], [(res[i], exp[i]) for i in bad[:2]])


@pytest.fixture(scope="module")
def ref_lib():
    if not C.have_ref():
        pytest.skip("compiled reference (oracle/_ref/libssw_ref.so) not present")
    return C.load_ref()


@pytest.mark.parametrize("flag", [0x0f, 0])
def test_config2_all_reads(engine, ref_lib, flag, capfd):
    ref, reads = _workload(5_000_000, 1000, 150, 1001, 2002)
    mat = C.dna_matrix(2, 2)
    engine.set_sequences(reads, [ref])
    res, pool = engine.align(mat, 5, 3, 1, flag=flag, filters=0, filterd=32767, mask_len=75, score_size=2)
    n = len(reads)
    check_all("config2_flag%d" % flag, capfd, res, pool, reads, [ref], np.arange(n), np.zeros(n), mat, 5,
              gapO=3, gapE=1, flag=flag, filters=0, filterd=32767, mask_len=75, score_size=2)


@pytest.mark.parametrize("mode", ["auto", "blocks_3_launches", "columns_3_launches"])
def test_config3_subsample(engine, ref_lib, mode, capfd):
    """the first 2,000 reads of the 100,000-read set of config 3 (seed 3003; the generator is sequential): in one launch
    (automatic: block column maxima), and forced through >= 3 launches by a capped column-maximum budget in both storage
    modes -- the multi-launch path the full 100,000-read batch takes."""
    ref, sub = _workload(5_000_000, 2000, 150, 1001, 3003)
    mat = C.dna_matrix(2, 2)
    engine.set_sequences(sub, [ref])
    if mode == "blocks_3_launches":
        engine.set_option("cm_block", 1)
        engine.set_option("cm_budget_mb", 110)           # 1,000 pair-tasks x 312 KB of block maxima = 312 MB
    elif mode == "columns_3_launches":
        engine.set_option("cm_block", 0)
        engine.set_option("cm_budget_mb", 7000)          # 1,000 pair-tasks x 20 MB of column maxima = 20 GB
    try:
        res, pool = engine.align(mat, 5, 3, 1, flag=0, mask_len=75, score_size=2)
        launches = engine.timing()["fill_forward_launches"]
    finally:
        engine.set_option("cm_block", -1)
        engine.set_option("cm_budget_mb", 0)
    if mode != "auto":
        assert launches >= 4, launches                   # >= 3 byte-semantics launches + the word re-fill of the overflows
    n = len(sub)
    check_all("config3_subsample_" + mode, capfd, res, pool, sub, [ref], np.arange(n), np.zeros(n), mat, 5, expect_key="config3",
              gapO=3, gapE=1, flag=0, filters=0, filterd=0, mask_len=75, score_size=2)


def test_config4_subgrid(engine, ref_lib, capfd):
    """200 queries x 1,000 targets of the BLOSUM50 grid (seeds 4004 / 4005), word scores, through the grid path"""
    W = C.config_workload(4, n_queries=200, n_targets=1000)
    queries, targets = W["queries"], W["refs"]
    engine.set_sequences(queries, targets)
    res, pool = engine.align(C.BLOSUM50, 24, 3, 1, flag=0, mask_len=150, score_size=1)
    pq = np.repeat(np.arange(200), 1000)
    pr = np.tile(np.arange(1000), 200)
    check_all("config4_subgrid", capfd, res, pool, queries, targets, pq, pr, C.BLOSUM50, 24,
              gapO=3, gapE=1, flag=0, filters=0, filterd=0, mask_len=150, score_size=1)


def test_config4_grid_in_launch_groups(engine, ref_lib, capfd):
    """96 queries x 50,000 targets = 4.8 M pairs: large enough for the device-planned grid to run as several launch groups whose
    records are copied back while the next group computes; every pair against the compiled reference."""
    W = C.config_workload(4, n_queries=96, n_targets=50_000)
    engine.set_sequences(W["queries"], W["refs"])
    res, pool = engine.align(C.BLOSUM50, 24, 3, 1, flag=0, mask_len=150, score_size=1)
    assert engine.timing()["fill_forward_launches"] >= 3
    pq = np.repeat(np.arange(96), 50_000)
    pr = np.tile(np.arange(50_000), 96)
    check_all("config4_launch_groups", capfd, res, pool, W["queries"], W["refs"], pq, pr, C.BLOSUM50, 24,
              gapO=3, gapE=1, flag=0, filters=0, filterd=0, mask_len=150, score_size=1)


@pytest.mark.parametrize("flag", [0x0f, 2])
def test_config5_all_long_reads(engine, ref_lib, flag, capfd):
    ref, reads = _workload(100_000, 1000, 10_000, 5005, 5006, decoy_frac=0.0, p_sub=0.05, p_ins=0.02, p_del=0.02)
    mat = C.dna_matrix(2, 2)
    engine.set_sequences(reads, [ref])
    res, pool = engine.align(mat, 5, 3, 1, flag=flag, filters=0, filterd=32767, mask_len=5000, score_size=2)
    assert int((res["cigar_len"] > 0).sum()) == len(reads)
    n = len(reads)
    check_all("config5_flag%d" % flag, capfd, res, pool, reads, [ref], np.arange(n), np.zeros(n), mat, 5,
              gapO=3, gapE=1, flag=flag, filters=0, filterd=32767, mask_len=5000, score_size=2)
