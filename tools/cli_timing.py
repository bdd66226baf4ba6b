#!/usr/bin/env python
"""Wall time of the command-line front ends on the reference's demo workload demo/1M.fa x demo/54mer_hap1_1.100.fa
(100 reads of 54 bp against a 1 Mbp reference; the sequences are rebuilt from the frozen goldens, the demo files
themselves do not travel to the GPU box):
  ssw_test_ref    the unmodified reference CLI on its own ssw.c (CPU, one blocking call per pair, main.c:462-532)
  ssw_test_b200   the SAME unmodified main.c linked against our libssw.so (drop-in: one GPU call per pair)
  ssw_batch_cli   our batched front end (both files parsed once, every pair in one ssw_align_batch call)
Options as in the README of the reference: default (scores/ends) and -c (CIGAR).   python tools/cli_timing.py [reps]"""
import json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = [c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "goldens.json"))) if c["name"] == "demo_old_txt_1M_x_54mer"][0]
ref = np.load(os.path.join(ROOT, "tests", "golden", g["refs_npz"]))["ref"] if "refs_npz" in g else None
L = "ACGTN"
tmp = tempfile.mkdtemp()
fa, fq = os.path.join(tmp, "1M.fa"), os.path.join(tmp, "reads.fa")
with open(fa, "w") as f:
    s = "".join(L[c] for c in ref)
    f.write(">chr1\n" + "\n".join(s[i:i + 70] for i in range(0, len(s), 70)) + "\n")
with open(fq, "w") as f:
    for i, r in enumerate(g["reads"]):
        f.write(">read%d\n%s\n" % (i, r if isinstance(r, str) else "".join(L[c] for c in r)))
exes = {"ssw_test_ref": os.path.join(ROOT, "oracle", "_ref", "ssw_test_ref"), "ssw_test_b200": os.path.join(ROOT, "oracle", "_ref", "ssw_test_b200"),
        "ssw_batch_cli": os.path.join(ROOT, "complete-striped-smith-waterman-library_b200", "ssw_batch_cli")}
out = {}
base = {}
for opts in ([], ["-c"]):
    for name, exe in exes.items():
        if not os.path.exists(exe):
            continue
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            r = subprocess.run([exe] + opts + [fa, fq], capture_output=True, text=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        key = " ".join(opts) or "default"
        body = "\n".join(l for l in r.stdout.splitlines() if not l.startswith("CPU time"))
        base.setdefault(key, body)
        out["%s [%s]" % (name, key)] = {"wall_ms": round(best * 1e3, 1), "rc": r.returncode, "same_output_as_reference": body == base[key]}
print(json.dumps(out, indent=1))
