#!/usr/bin/env python
"""Config 5 driven through two engines from two host threads (each takes half of the reads): the latency-bound traceback
of one half overlaps the fills of the other.  python tools/two_engines.py [reads]"""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C
from __graft_entry__ import load_package

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
L = load_package()
ref, reads = C.make_dna_workload(100_000, n_reads, 10_000, seed_ref=5005, seed_reads=5006, decoy_frac=0.0, p_sub=0.05, p_ins=0.02, p_del=0.02)
mat = C.dna_matrix(2, 2)


def run(eng, part):
    eng.set_sequences(part, [ref])
    return eng.align(mat, 5, 3, 1, flag=2, filters=0, filterd=32767, mask_len=5000, score_size=2)


one = L.BatchAligner(device=0)
run(one, reads)
t0 = time.perf_counter(); res1, pool1 = run(one, reads); t_one = time.perf_counter() - t0
for k in (2, 3):
    engs = [L.BatchAligner(device=0) for _ in range(k)]
    parts = [reads[i * n_reads // k: (i + 1) * n_reads // k] for i in range(k)]
    with ThreadPoolExecutor(k) as ex:
        list(ex.map(lambda a: run(*a), zip(engs, parts)))
        t0 = time.perf_counter()
        out = list(ex.map(lambda a: run(*a), zip(engs, parts)))
        t_k = time.perf_counter() - t0
    ok = True
    off = 0
    for r, _ in out:
        ok = ok and (r["score1"] == res1["score1"][off: off + len(r)]).all() and (r["cigar_len"] == res1["cigar_len"][off: off + len(r)]).all()
        off += len(r)
    print({"reads": n_reads, "one_engine_ms": round(t_one * 1e3, 1), "engines": k, "ms": round(t_k * 1e3, 1), "same_results": bool(ok)})
    for e in engs:
        e.close()
