#!/usr/bin/env python
"""Random batches through the batch ABI on the CPU emulator (tests/cuda_emu) against the oracle: mixed query / reference
lengths, full grids and explicit pair lists, all flags, byte / word / byte-then-word, forced small chunks, the
device-planned grid and the sliced path.   python tools/fuzz_batch_emulator.py [n_batches] [seed]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C
from test_emulated_kernels import EMU_DIR, _pkg

n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
LIB = os.environ.get("SSW_FUZZ_LIB")        # another build of the emulator library, e.g. one made with -fsanitize=address
if not LIB:
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True, stdout=sys.stderr)
    LIB = os.path.join(EMU_DIR, "libssw_emu.so")
L = _pkg()
eng = L.BatchAligner(lib_dir=os.path.dirname(LIB), lib_name=os.path.basename(LIB))
oracle = C.load_oracle()
rng = np.random.default_rng(seed)
bad = 0
pairs_checked = 0
for b in range(n_batches):
    prot = rng.random() < 0.25
    n = 24 if prot else 5
    alpha = 20 if prot else 4
    mat = C.BLOSUM50.copy() if prot else C.dna_matrix(int(rng.integers(1, 4)), int(rng.integers(1, 4)))
    gapE = int(rng.integers(1, 3)); gapO = gapE + int(rng.integers(1, 6))
    refs = [rng.integers(0, alpha, size=int(rng.integers(20, 900))).astype(np.int8) for _ in range(int(rng.integers(1, 5)))]
    qs = []
    for _ in range(int(rng.integers(1, 9))):
        r = refs[int(rng.integers(0, len(refs)))]
        ql = int(rng.integers(5, 400))
        if rng.random() < 0.7 and len(r) > ql + 3:
            qs.append(C.mutate_read(rng, r, int(rng.integers(0, len(r) - ql)), ql, 0.08, 0.02, 0.02, alpha))
        else:
            qs.append(rng.integers(0, alpha, size=ql).astype(np.int8))
    flag = int(rng.choice([0, 0, 1, 2, 8, 0x0f]))
    ss = int(rng.choice([0, 1, 2, 2]))
    mask = int(rng.choice([15, 20, 40]))
    filters, filterd = int(rng.integers(0, 40)), int(rng.choice([30, 200, 32767]))
    eng.set_option("chunk", int(rng.choice([0, 0, 64, 128])))
    eng.set_option("grid_min", int(rng.choice([1, 32768])))
    eng.set_option("latency_cols", int(rng.choice([0, 1 << 20])))
    eng.set_option("slices", int(rng.choice([0, 2, 3])))
    eng.set_option("parts", int(rng.choice([0, 2])))
    eng.set_sequences(qs, refs)
    if rng.random() < 0.5:
        pq = pr = None
        idx = [(i, j) for i in range(len(qs)) for j in range(len(refs))]
    else:
        m = int(rng.integers(1, 12))
        pq = rng.integers(0, len(qs), size=m).astype(np.int32)
        pr = rng.integers(0, len(refs), size=m).astype(np.int32)
        idx = list(zip(pq.tolist(), pr.tolist()))
    res, pool = eng.align(mat, n, gapO, gapE, flag=flag, filters=filters, filterd=filterd, mask_len=mask, score_size=ss, pair_query=pq, pair_ref=pr)
    for k, (i, j) in enumerate(idx):
        exp = oracle.align(qs[i], refs[j], mat, n, gapO, gapE, flag, filters, filterd, mask, ss)
        r = res[k]
        pairs_checked += 1
        if exp is None:
            if int(r["status"]) != 1:
                bad += 1; print("MISMATCH batch", b, "pair", k, "expected NULL", flush=True)
            continue
        got = {f: int(r[f]) for f in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "flag")}
        got["cigar"] = [int(x) for x in pool[r["cigar_off"]: r["cigar_off"] + r["cigar_len"]]] if r["cigar_off"] >= 0 else []
        d = C.diff_results(got, exp)
        if d:
            bad += 1
            print("MISMATCH batch", b, "seed", seed, "pair", k, d, dict(flag=flag, ss=ss, gapO=gapO, gapE=gapE, n=n, mask=mask), flush=True)
print({"batches": n_batches, "pairs": pairs_checked, "seed": seed, "mismatches": bad})
