#!/usr/bin/env python
"""Long random campaign of the product's kernels on the CPU emulator (tests/cuda_emu) against the oracle.
  python tools/fuzz_emulator.py [n_cases] [seed]     prints the mismatching cases (none expected)"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C
from test_oracle import random_case
from test_emulated_kernels import EMU_DIR, _latency_instances

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
subprocess.run(["make", "-s", "-C", EMU_DIR], check=True, stdout=sys.stderr)
emu = C.SswLib(os.path.join(EMU_DIR, "libssw_emu.so"))
oracle = C.load_oracle()
rng = np.random.default_rng(seed)
bad = 0
for k in range(n):
    if k % 250 == 0:
        _latency_instances((k // 250) % 2 == 0)
    c = random_case(rng)
    d = C.diff_results(emu.align(**c), oracle.align(**c))
    if d:
        bad += 1
        print("MISMATCH case", k, "seed", seed, d, {x: (c[x] if not hasattr(c[x], "shape") else c[x].tolist()) for x in c}, flush=True)
print({"cases": n, "seed": seed, "mismatches": bad})
