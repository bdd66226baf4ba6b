"""Parity at the sizes SURVEY 8(d) asks for, against the COMPILED reference (oracle/_ref/libssw_ref.so) run on all host
cores of the GPU box:
  config 2: all 1,000 reads x 5 Mbp, flag 0x0f (scores, ends, begins, flag, CIGAR) and flag 0;
  config 3: a fixed 2,000-read subsample of the 100,000-read set;
  config 4: a 200 x 1,000 sub-grid of the protein grid, word scores;
  config 5: 1,000 x 10 kbp long reads, flags 0x0f and 2, every field and every CIGAR word.
The reference side is CPU-bound (~1e12 cells per case); each case checks its pairs in blocks until a time budget is
used up, requires a minimum number checked, and records the coverage in gpurun_out/parity_full.json."""
import json
import os
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import common as C
from test_gpu_parity import FIELDS8, batch_dict, engine  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

BUDGET_S = float(os.environ.get("SSW_FULL_PARITY_BUDGET", "75"))
THREADS = max(1, (os.cpu_count() or 2) - 2)


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


def check_pairs(name, res, pool, pairs, ref_call, minimum):
    """pairs: list of (result index, thunk args); ref_call(args) -> expected dict.  Blocks of THREADS*2 pairs."""
    t0 = time.time()
    done = 0
    bad = []
    with ThreadPoolExecutor(THREADS) as ex:
        step = THREADS * 2
        for lo in range(0, len(pairs), step):
            if done >= minimum and time.time() - t0 > BUDGET_S:
                break
            block = pairs[lo: lo + step]
            for (idx, _), exp in zip(block, ex.map(ref_call, [a for _, a in block])):
                d = C.diff_results(batch_dict(res, pool, idx), exp)
                if d:
                    bad.append((idx, d))
            done += len(block)
    _record(name, {"pairs_total": len(pairs), "pairs_checked": done, "mismatches": len(bad), "threads": THREADS,
                   "seconds": round(time.time() - t0, 1)})
    assert not bad, bad[:3]
    assert done >= minimum
    return done


@pytest.fixture(scope="module")
def ref_lib():
    if not C.have_ref():
        pytest.skip("compiled reference (oracle/_ref/libssw_ref.so) not present")
    return C.load_ref()


@pytest.mark.parametrize("flag", [0x0f, 0])
def test_config2_all_reads(engine, ref_lib, flag):
    ref, reads = _workload(5_000_000, 1000, 150, 1001, 2002)
    mat = C.dna_matrix(2, 2)
    engine.set_sequences(reads, [ref])
    res, pool = engine.align(mat, 5, 3, 1, flag=flag, filters=0, filterd=32767, mask_len=75, score_size=2)
    pairs = [(i, i) for i in range(len(reads))]
    check_pairs("config2_flag%d" % flag, res, pool, pairs,
                lambda i: ref_lib.align(reads[i], ref, mat, 5, 3, 1, flag, 0, 32767, 75, 2), minimum=256)


def test_config3_subsample(engine, ref_lib):
    """the first 2,000 reads of the 100,000-read set of config 3 (seed 3003; the generator is sequential)"""
    ref, sub = _workload(5_000_000, 2000, 150, 1001, 3003)
    mat = C.dna_matrix(2, 2)
    engine.set_sequences(sub, [ref])
    res, pool = engine.align(mat, 5, 3, 1, flag=0, mask_len=75, score_size=2)
    pairs = [(i, i) for i in range(len(sub))]
    check_pairs("config3_subsample", res, pool, pairs,
                lambda i: ref_lib.align(sub[i], ref, mat, 5, 3, 1, 0, 0, 0, 75, 2), minimum=256)


def test_config4_subgrid(engine, ref_lib):
    """200 queries x 1,000 targets of the BLOSUM50 grid (seeds 4004 / 4005), word scores, through the grid path"""
    rq, rt = np.random.default_rng(4004), np.random.default_rng(4005)
    queries = [rq.integers(0, 20, size=300).astype(np.int8) for _ in range(200)]
    targets = []
    for t in range(1000):
        s = rt.integers(0, 20, size=400).astype(np.int8)
        if t % 10 == 0:
            q = queries[int(rt.integers(0, len(queries)))]
            a = int(rt.integers(0, 100))
            seg = q[a: a + 200].copy()
            m = rt.random(len(seg)) < 0.2
            seg[m] = rt.integers(0, 20, size=int(m.sum()))
            b = int(rt.integers(0, 200))
            s[b: b + 200] = seg
        targets.append(s)
    engine.set_sequences(queries, targets)
    res, pool = engine.align(C.BLOSUM50, 24, 3, 1, flag=0, mask_len=150, score_size=1)
    pairs = [(q * 1000 + t, (q, t)) for q in range(200) for t in range(1000)]
    check_pairs("config4_subgrid", res, pool, pairs,
                lambda a: ref_lib.align(queries[a[0]], targets[a[1]], C.BLOSUM50, 24, 3, 1, 0, 0, 0, 150, 1), minimum=20_000)


@pytest.mark.parametrize("flag", [0x0f, 2])
def test_config5_all_long_reads(engine, ref_lib, flag):
    ref, reads = _workload(100_000, 1000, 10_000, 5005, 5006, decoy_frac=0.0, p_sub=0.05, p_ins=0.02, p_del=0.02)
    mat = C.dna_matrix(2, 2)
    engine.set_sequences(reads, [ref])
    res, pool = engine.align(mat, 5, 3, 1, flag=flag, filters=0, filterd=32767, mask_len=5000, score_size=2)
    assert int((res["cigar_len"] > 0).sum()) == len(reads)
    pairs = [(i, i) for i in range(len(reads))]
    check_pairs("config5_flag%d" % flag, res, pool, pairs,
                lambda i: ref_lib.align(reads[i], ref, mat, 5, 3, 1, flag, 0, 32767, 5000, 2), minimum=128)
