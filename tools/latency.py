#!/usr/bin/env python
"""Latency of the one-pair drop-in call (ssw_init + ssw_align + align_destroy through libssw.so) for small inputs:
what a legacy caller that loops over pairs pays per call.  python tools/latency.py [n_calls]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
lib = C.load_ours()
rng = np.random.default_rng(5)
mat = C.dna_matrix(2, 2)
out = {}
for name, rl, ql, flag in (("150bp_vs_1kbp_flag0", 1000, 150, 0), ("150bp_vs_1kbp_flag2", 1000, 150, 2), ("150bp_vs_100kbp_flag0", 100_000, 150, 0)):
    ref = rng.integers(0, 4, size=rl, dtype=np.int8)
    q = C.mutate_read(rng, ref, rl // 3, ql, 0.05, 0.01, 0.01)
    lib.align(q, ref, mat, 5, 3, 1, flag, 0, 32767, 75, 2)
    t0 = time.perf_counter()
    for _ in range(n):
        lib.align(q, ref, mat, 5, 3, 1, flag, 0, 32767, 75, 2)
    out[name] = round((time.perf_counter() - t0) / n * 1e6, 1)
print({"us_per_call": out, "calls": n})
if os.environ.get("SSW_TRACE"):
    for rl in (1000, 100_000):
        sys.stderr.write("---- one traced call (150 bp vs %d bp, flag 0) ----\n" % rl)
        ref = rng.integers(0, 4, size=rl, dtype=np.int8)
        q = C.mutate_read(rng, ref, rl // 3, 150, 0.05, 0.01, 0.01)
        lib.align(q, ref, mat, 5, 3, 1, 0, 0, 32767, 75, 2)
        t0 = time.perf_counter()
        lib.align(q, ref, mat, 5, 3, 1, 0, 0, 32767, 75, 2)
        sys.stderr.write("whole call %.1f us\n" % ((time.perf_counter() - t0) * 1e6))
