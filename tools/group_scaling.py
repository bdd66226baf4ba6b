#!/usr/bin/env python
"""One batch over 1 .. N GPUs of ONE process through the device groups (ssw_group_align, include/ssw_batch.h): config 2
(1,000 x 150 bp vs 5 Mbp; host buffers in, records out: the wall time of the call incl. the uploads) at every group size, the
records of every size compared with those of size 1 and a sample with the CPU checker.
   python tools/group_scaling.py [--reads N] [--ref-len L] [--sizes 1,2] [--reps 3] [--lib-dir D --lib-name F]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C
from __graft_entry__ import load_package

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1000)
ap.add_argument("--ref-len", type=int, default=5_000_000)
ap.add_argument("--sizes", default="1,2")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", type=int, default=32)
ap.add_argument("--lib-dir", default=None)
ap.add_argument("--lib-name", default="libssw.so")
a = ap.parse_args()
L = load_package()
ref, reads = C.make_dna_workload(a.ref_len, a.reads, 150, seed_ref=1001, seed_reads=2002)
mat = C.dna_matrix(2, 2)
cells = float(sum(len(q) for q in reads)) * float(len(ref))
base = None
for size in [int(x) for x in a.sizes.split(",")]:
    grp = L.GroupAligner(n_devices=size, lib_dir=a.lib_dir, lib_name=a.lib_name)
    times = []
    for rep in range(a.reps + 1):                                   # the first call is the warm-up (allocations, contexts)
        t0 = time.perf_counter()
        res, pool = grp.align(reads, [ref], mat, 5, 3, 1, flag=0, mask_len=75, score_size=2)
        times.append(time.perf_counter() - t0)
    tm = grp.timing()
    same = True
    if base is None:
        base = res.copy()
    else:
        same = all(bool(np.array_equal(res[f], base[f])) for f in ("score1", "score2", "ref_end1", "read_end1", "ref_end2", "status"))
    best = min(times[1:])
    print(json.dumps({"devices": size, "pairs": len(res), "best_ms": round(best * 1e3, 2), "median_ms": round(float(np.median(times[1:])) * 1e3, 2),
                      "gcups_e2e": round(cells / best / 1e9, 1), "same_as_one_device": same,
                      "fill_ms_per_device": [round(t["fill_forward_ms"], 2) for t in tm], "cells_per_device": [int(t["cells_forward"]) for t in tm]}), flush=True)
    grp.close()
if a.check:
    idx = np.linspace(0, len(base) - 1, min(a.check, len(base))).astype(int)
    exp, exp_pool, _, _, kind = C.cpu_batch(reads, [ref], idx, np.zeros(len(idx), dtype=np.int64), mat, 5, 3, 1, flag=0, mask_len=75, score_size=2)
    bad = C.compare_records(base, np.zeros(0, np.uint32), exp, exp_pool, idx=idx)
    print(json.dumps({"checked": len(idx), "checker": kind, "mismatches": len(bad)}))
