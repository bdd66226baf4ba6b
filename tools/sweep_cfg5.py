#!/usr/bin/env python
"""Config 5 (1,000 x 10 kbp reads vs 100 kbp, flag 2) under different engine options in ONE process: the read set is
generated once, every option set gets `--reps` calls, the best and median wall times are printed.
  python tools/sweep_cfg5.py "slices=4" "slices=5,slice_taper=70" ...        (an empty string = defaults)"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C
from __graft_entry__ import load_package

ap = argparse.ArgumentParser()
ap.add_argument("sets", nargs="*", default=[""])
ap.add_argument("--reads", type=int, default=1000)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--flag", type=int, default=2)
ap.add_argument("--lib", default="libssw.so")
a = ap.parse_args()
L = load_package()
ref, reads = C.make_dna_workload(100_000, a.reads, 10_000, seed_ref=5005, seed_reads=5006, decoy_frac=0.0, p_sub=0.05, p_ins=0.02, p_del=0.02)
mat = C.dna_matrix(2, 2)
cells = float(sum(len(q) for q in reads)) * float(len(ref))
ALL = ("carve", "slices", "slice_taper", "super", "parts", "tb_spec")
DEFAULTS = {"carve": 1, "slices": 0, "slice_taper": 0, "super": 0, "parts": 0, "tb_spec": -1}
eng = L.BatchAligner(device=0, lib_name=a.lib)
eng.set_sequences(reads, [ref])
eng.align(mat, 5, 3, 1, flag=a.flag, filters=0, filterd=32767, mask_len=5000, score_size=2)      # warm-up (allocations)
first = None
for spec in a.sets:
    opts = dict(DEFAULTS)
    for kv in [x for x in spec.split(",") if x]:
        opts[kv.split("=")[0]] = int(kv.split("=")[1])
    for k in ALL:
        eng.set_option(k, opts[k])
    eng.align(mat, 5, 3, 1, flag=a.flag, filters=0, filterd=32767, mask_len=5000, score_size=2)  # helper engines / streams of this shape
    walls, tms = [], []
    for rep in range(a.reps):
        t0 = time.perf_counter()
        res, pool = eng.align(mat, 5, 3, 1, flag=a.flag, filters=0, filterd=32767, mask_len=5000, score_size=2)
        walls.append((time.perf_counter() - t0) * 1e3)
        tms.append(eng.timing())
    sig = (int(np.sum(res["score1"].astype(np.int64))), int(np.sum(res["cigar_len"].astype(np.int64))), int(np.sum(np.asarray(pool, dtype=np.int64))))
    if first is None:
        first = sig
    b = int(np.argmin(walls))
    print(json.dumps({"lib": a.lib, "opts": spec or "defaults", "best_ms": round(min(walls), 2), "median_ms": round(float(np.median(walls)), 2),
                      "gcups_best": round(cells / min(walls) / 1e6, 1), "same_results": sig == first,
                      "fwd": round(tms[b]["fill_forward_ms"], 1), "rev": round(tms[b]["fill_reverse_ms"], 1), "tb": round(tms[b]["traceback_ms"], 1)}), flush=True)
eng.close()
