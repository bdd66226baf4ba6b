#!/usr/bin/env python
"""Time one BASELINE.json configuration on one GPU through the batch ABI and print the engine's breakdown.
  python tools/run_config.py 5 [--reads N] [--flag F]     config 5: 10 kbp reads x 100 kbp reference (word path, CIGAR)
  python tools/run_config.py 4 [--queries Q --targets T]  config 4: protein BLOSUM50 300 aa x 400 aa (word path)
  python tools/run_config.py 2                            config 2 (1,000 x 150 bp vs 5 Mbp)
  python tools/run_config.py 3 --reads 100000             config 3 (bench.py's headline batch on one GPU)
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C
from __graft_entry__ import load_package

ap = argparse.ArgumentParser()
ap.add_argument("config", type=int)
ap.add_argument("--reads", type=int, default=1000)
ap.add_argument("--flag", type=int, default=-1)
ap.add_argument("--queries", type=int, default=64)
ap.add_argument("--targets", type=int, default=50000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--lib", default="libssw.so")
ap.add_argument("--opt", action="append", default=[], help="engine option name=value (repeatable)")
ap.add_argument("--check", type=int, default=0, help="compare this many pairs with the reference/oracle")
a = ap.parse_args()
L = load_package()
eng = L.BatchAligner(device=0, lib_name=a.lib)
for kv in a.opt:
    eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
if a.config == 5:
    ref, reads = C.make_dna_workload(100_000, a.reads, 10_000, seed_ref=5005, seed_reads=5006, decoy_frac=0.0, p_sub=0.05, p_ins=0.02, p_del=0.02)
    mat, n, flag, ml, ss = C.dna_matrix(2, 2), 5, (2 if a.flag < 0 else a.flag), 5000, 2
    qs, rs = reads, [ref]
elif a.config == 4:
    rq, rt = np.random.default_rng(4004), np.random.default_rng(4005)
    qs = [rq.integers(0, 20, size=300).astype(np.int8) for _ in range(a.queries)]
    rs = []
    for t in range(a.targets):
        s = rt.integers(0, 20, size=400).astype(np.int8)
        if t % 10 == 0:
            q = qs[int(rt.integers(0, len(qs)))]
            seg = q[50:250].copy(); m = rt.random(200) < 0.2; seg[m] = rt.integers(0, 20, size=int(m.sum())); s[100:300] = seg
        rs.append(s)
    mat, n, flag, ml, ss = C.BLOSUM50, 24, (0 if a.flag < 0 else a.flag), 150, 1
elif a.config == 3:
    ref, reads = C.make_dna_workload(5_000_000, a.reads, 150, seed_ref=1001, seed_reads=3003)
    mat, n, flag, ml, ss = C.dna_matrix(2, 2), 5, (0 if a.flag < 0 else a.flag), 75, 2
    qs, rs = reads, [ref]
else:
    ref, reads = C.make_dna_workload(5_000_000, a.reads, 150, seed_ref=1001, seed_reads=2002)
    mat, n, flag, ml, ss = C.dna_matrix(2, 2), 5, (0 if a.flag < 0 else a.flag), 75, 2
    qs, rs = reads, [ref]
cells = float(sum(len(q) for q in qs)) * float(sum(len(r) for r in rs))
eng.set_sequences(qs, rs)
best = None
for rep in range(a.reps):
    t0 = time.perf_counter()
    res, pool = eng.align(mat, n, 3, 1, flag=flag, filters=0, filterd=32767, mask_len=ml, score_size=ss)
    dt = time.perf_counter() - t0
    tm = eng.timing()
    print(json.dumps({"config": a.config, "rep": rep, "pairs": len(res), "wall_ms": dt * 1e3, "gcups_wall": cells / dt / 1e9,
                      "gcups_fill": cells / (tm["fill_forward_ms"] * 1e-3 + 1e-12) / 1e9, **{k: (round(v, 2) if isinstance(v, float) else v) for k, v in tm.items()}}))
if a.check:
    chk = C.load_ref() if C.have_ref() else C.load_oracle()
    bad = 0
    idx = np.linspace(0, len(res) - 1, a.check).astype(int)
    for p in idx:
        q, r = qs[p // len(rs)], rs[p % len(rs)]
        exp = chk.align(q, r, mat, n, 3, 1, flag, 0, 32767, ml, ss)
        got = {k: int(res[p][k]) for k in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "flag")}
        got["cigar"] = [int(x) for x in pool[res[p]["cigar_off"]: res[p]["cigar_off"] + res[p]["cigar_len"]]] if res[p]["cigar_off"] >= 0 else []
        if C.diff_results(got, exp): bad += 1
    print(json.dumps({"checked": len(idx), "mismatches": bad}))
eng.close()
