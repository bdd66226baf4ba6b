#!/usr/bin/env python
"""Random score-only full grids through the device-planned path (ssw_grid.cuh + late arming + launch groups) on the CPU
emulator (tests/cuda_emu) against the CPU checker: protein (BLOSUM50) and DNA, equal and ragged query lengths around the
kernel-instance boundaries (… 152/160, 256, 289-304/320 rows …), random arming tails, group and split sizes, byte / word /
byte-then-word.   python tools/fuzz_grid_emulator.py [n_grids] [seed]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C
from test_emulated_kernels import EMU_DIR, _pkg

n_grids = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
LIB = os.environ.get("SSW_FUZZ_LIB")        # another build of the emulator library, e.g. one made with -fsanitize=address
if not LIB:
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True, stdout=sys.stderr)
    LIB = os.path.join(EMU_DIR, "libssw_emu.so")
L = _pkg()
eng = L.BatchAligner(lib_dir=os.path.dirname(LIB), lib_name=os.path.basename(LIB))
eng.set_option("latency_cols", 0)
eng.set_option("grid_min", 1)
rng = np.random.default_rng(seed)
bad = 0
pairs_checked = 0
redone = 0
LENS = [8, 31, 64, 100, 128, 150, 152, 160, 161, 200, 255, 256, 257, 288, 289, 300, 304, 305, 320]
for g in range(n_grids):
    prot = rng.random() < 0.6
    n = 24 if prot else 5
    alpha = 20 if prot else 4
    mat = C.BLOSUM50.copy() if prot else C.dna_matrix(int(rng.integers(1, 4)), int(rng.integers(1, 4)))
    gapE = int(rng.integers(1, 3)); gapO = gapE + int(rng.integers(1, 11))
    nq, nt = int(rng.integers(1, 7)), int(rng.integers(1, 40))
    if rng.random() < 0.5:
        ql = [int(rng.choice(LENS))] * nq                       # equal lengths: grouped launches
    else:
        ql = [int(rng.choice(LENS)) if rng.random() < 0.7 else int(rng.integers(1, 321)) for _ in range(nq)]
    qs = [rng.integers(0, alpha, size=k).astype(np.int8) for k in ql]
    ts = []
    for _ in range(nt):
        t = rng.integers(0, alpha, size=int(rng.integers(10, 500))).astype(np.int8)
        if rng.random() < 0.5:                                 # embed a mutated piece of a query: a real maximum somewhere in the target
            q = qs[int(rng.integers(0, nq))]
            k = int(rng.integers(1, len(q) + 1)); a = int(rng.integers(0, len(q) - k + 1))
            piece = q[a:a + k].copy()
            m = rng.random(k) < 0.15
            piece[m] = rng.integers(0, alpha, size=int(m.sum()))
            if k <= len(t):
                b = int(rng.integers(0, len(t) - k + 1))
                t[b:b + k] = piece
        ts.append(t)
    ss = int(rng.choice([0, 1, 1, 2, 2]))
    mask = int(rng.choice([15, 20, 75, 150]))
    eng.set_option("grid_arm", int(rng.choice([-1, -1, 0, 10, 60, 200])))
    eng.set_option("grid_split", int(rng.choice([-1, 1, 7, 64])))
    eng.set_option("grid_group", int(rng.choice([-1, 1, 2])))
    eng.set_option("chunk", int(rng.choice([0, 0, 64])))
    eng.set_sequences(qs, ts)
    res, pool = eng.align(mat, n, gapO, gapE, flag=0, mask_len=mask, score_size=ss)
    redone += eng.timing()["byte_overflows"]
    pq, pr = np.repeat(np.arange(nq), nt), np.tile(np.arange(nt), nq)
    exp, exp_pool, _, _, _ = C.cpu_batch(qs, ts, pq, pr, mat, n, gapO, gapE, flag=0, mask_len=mask, score_size=ss, threads=2)
    d = C.compare_records(res, pool, exp, exp_pool)
    pairs_checked += len(pq)
    if d:
        bad += len(d)
        print("MISMATCH grid", g, "seed", seed, dict(prot=prot, ql=ql, nt=nt, ss=ss, gapO=gapO, gapE=gapE, mask=mask), d[:3], flush=True)
print({"grids": n_grids, "pairs": pairs_checked, "seed": seed, "redone_by_general_path": redone, "mismatches": bad})
