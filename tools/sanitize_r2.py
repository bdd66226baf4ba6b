#!/usr/bin/env python
"""Small workloads that walk every kernel path added in round 2, for `compute-sanitizer --tool memcheck python tools/sanitize_r2.py`:
block column maxima + re-fill items (forced on short references), eight-warp CTAs with a shared protein profile, the device-planned
grid in launch groups, the speculative traceback kernel, the engine pool behind ssw_align, the resident-reference cache.
Every result is compared with the CPU checker, so a clean sanitizer log also means correct output."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C
from __graft_entry__ import load_package
L = load_package()
eng = L.BatchAligner(device=0)
rng = np.random.default_rng(7)
checked = 0

def check(res, pool, queries, refs, pq, pr, mat, n, **kw):
    global checked
    exp, ep, _, _, _ = C.cpu_batch(queries, refs, pq, pr, mat, n, threads=8, **kw)
    bad = C.compare_records(res, pool, exp, ep)
    assert bad == [], bad[:3]
    checked += len(exp)

# 1. block column maxima on a 40 kbp reference, several chunk lengths, byte and word semantics
mat = C.dna_matrix(2, 2)
ref, reads = C.make_dna_workload(40_000, 24, 150, seed_ref=3, seed_reads=4)
eng.set_option("latency_cols", 0); eng.set_option("cm_block", 1)
eng.set_sequences(reads, [ref])
for chunk in (0, 4096, 1024):
    eng.set_option("chunk", chunk)
    res, pool = eng.align(mat, 5, 3, 1, flag=0, mask_len=75, score_size=2)
    check(res, pool, reads, [ref], np.arange(24), np.zeros(24), mat, 5, gapO=3, gapE=1, flag=0, mask_len=75, score_size=2)
eng.set_option("chunk", 0); eng.set_option("cm_block", -1)
# 2. protein grid: eight-warp CTAs, launch groups with per-group copy back
W = C.config_workload(4, n_queries=8, n_targets=300)
eng.set_option("grid_min", 1); eng.set_option("grid_split", 1); eng.set_option("grid_group", 1)
eng.set_sequences(W["queries"], W["refs"])
res, pool = eng.align(C.BLOSUM50, 24, 3, 1, flag=0, mask_len=150, score_size=1)
check(res, pool, W["queries"], W["refs"], np.repeat(np.arange(8), 300), np.tile(np.arange(300), 8), C.BLOSUM50, 24, gapO=3, gapE=1, flag=0, mask_len=150, score_size=1)
eng.set_option("grid_min", -1); eng.set_option("grid_split", -1); eng.set_option("grid_group", -1)
# 3. long reads with CIGARs: strip pipeline, reverse pass, speculative traceback (small batch -> automatic) and the one-after-the-other variant
refL, readsL = C.make_dna_workload(12_000, 6, 1500, seed_ref=5, seed_reads=6, decoy_frac=0.0, p_sub=0.05, p_ins=0.02, p_del=0.02)
eng.set_sequences(readsL, [refL])
for spec in (-1, 0):
    eng.set_option("tb_spec", spec)
    res, pool = eng.align(mat, 5, 3, 1, flag=2, filters=0, filterd=32767, mask_len=750, score_size=2)
    check(res, pool, readsL, [refL], np.arange(6), np.zeros(6), mat, 5, gapO=3, gapE=1, flag=2, filters=0, filterd=32767, mask_len=750, score_size=2)
eng.set_option("tb_spec", -1)
eng.close()
# 4. the drop-in call: engine pool from four threads, resident-reference cache with an in-place rewrite
from concurrent.futures import ThreadPoolExecutor
ours, ref_lib = C.load_ours(), (C.load_ref() if C.have_ref() else C.load_oracle())
r2 = rng.integers(0, 4, size=9000, dtype=np.int8)
qs = [C.mutate_read(rng, r2, int(rng.integers(0, 8800)), 90, 0.05, 0.01, 0.01) for _ in range(12)]
def one(q):
    return ours.align(q, r2, mat, 5, 3, 1, 0x0f, 0, 32767, 45, 2)
for rnd in range(2):
    with ThreadPoolExecutor(4) as ex:
        got = list(ex.map(one, qs))
    for q, g in zip(qs, got):
        assert C.diff_results(g, ref_lib.align(q, r2, mat, 5, 3, 1, 0x0f, 0, 32767, 45, 2)) == []
        checked += 1
    r2[:] = np.roll(r2, 333)
print("sanitize_r2: %d results equal to the checker" % checked)
