#!/usr/bin/env python
"""Random long-read batches (queries of several strips, structural indels that force band doublings over several 32-column
tiles) through the strip pipeline, the reverse pass and the banded traceback on the CPU emulator (tests/cuda_emu) against
the CPU checker, every CIGAR word: random super-block size, strip splits, slices, traceback kernel choice and speculation.
   python tools/fuzz_long_emulator.py [n_batches] [seed]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C
from test_emulated_kernels import EMU_DIR, _pkg

n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
LIB = os.environ.get("SSW_FUZZ_LIB")        # another build of the emulator library, e.g. one made with -fsanitize=address
if not LIB:
    subprocess.run(["make", "-s", "-C", EMU_DIR], check=True, stdout=sys.stderr)
    LIB = os.path.join(EMU_DIR, "libssw_emu.so")
L = _pkg()
eng = L.BatchAligner(lib_dir=os.path.dirname(LIB), lib_name=os.path.basename(LIB))
eng.set_option("latency_cols", 0)
rng = np.random.default_rng(seed)
bad = 0
pairs_checked = 0
cigar_words = 0


def structural(rng, ref, start, length):
    """A read copied from ref[start:], with point noise plus a few long deletions / insertions (bands of several tiles)."""
    out = []
    pos = start
    while sum(len(x) for x in out) < length and pos < len(ref) - 5:
        seg = int(rng.integers(30, 400))
        piece = ref[pos:pos + seg].copy()
        m = rng.random(len(piece)) < rng.choice([0.0, 0.03, 0.1])
        piece[m] = rng.integers(0, 4, size=int(m.sum()))
        out.append(piece)
        pos += seg
        ev = rng.random()
        if ev < 0.25:
            pos += int(rng.integers(1, 150))                              # deletion in the read
        elif ev < 0.5:
            out.append(rng.integers(0, 4, size=int(rng.integers(1, 90))).astype(np.int8))      # insertion
    r = np.concatenate(out)[:length].astype(np.int8)
    return r if len(r) else ref[start:start + 5].copy()


for b in range(n_batches):
    mat = C.dna_matrix(int(rng.integers(1, 4)), int(rng.integers(1, 5)))
    gapE = int(rng.integers(1, 3)); gapO = gapE + int(rng.integers(0, 6))
    refs = [rng.integers(0, 4, size=int(rng.integers(600, 5000))).astype(np.int8) for _ in range(int(rng.integers(1, 4)))]
    qs = []
    for _ in range(int(rng.integers(1, 6))):
        r = refs[int(rng.integers(0, len(refs)))]
        ql = int(rng.choice([int(rng.integers(330, 700)), int(rng.integers(700, 2600)), int(rng.integers(20, 330))]))
        ql = min(ql, len(r) - 50)
        st = int(rng.integers(0, max(1, len(r) - ql)))
        if rng.random() < 0.6:
            qs.append(structural(rng, r, st, ql))
        else:
            qs.append(C.mutate_read(rng, r, st, ql, float(rng.choice([0.02, 0.08])), 0.02, 0.02))
    flag = int(rng.choice([2, 0x0f, 0x0f, 1, 8, 0]))
    ss = int(rng.choice([1, 2, 2]))
    mask = int(rng.choice([15, 50, 300]))
    filterd = int(rng.choice([32767, 32767, 500]))
    eng.set_option("super", int(rng.choice([0, 64, 256, 1024])))
    eng.set_option("parts", int(rng.choice([0, 1, 2, 4])))
    eng.set_option("slices", int(rng.choice([0, 0, 2, 4])))
    eng.set_option("slice_taper", int(rng.choice([0, 60, 80])))
    eng.set_option("tb_maxbw", int(rng.choice([-1, -1, 0, 64])))
    eng.set_option("tb_spec", int(rng.choice([-1, 0, 0, 4096])))
    eng.set_option("chunk", int(rng.choice([0, 0, 256])))
    eng.set_sequences(qs, refs)
    if rng.random() < 0.5:
        pq = np.repeat(np.arange(len(qs)), len(refs)).astype(np.int32); pr = np.tile(np.arange(len(refs)), len(qs)).astype(np.int32)
        res, pool = eng.align(mat, 5, gapO, gapE, flag=flag, filters=0, filterd=filterd, mask_len=mask, score_size=ss)
    else:
        m = int(rng.integers(1, 8))
        pq = rng.integers(0, len(qs), size=m).astype(np.int32)
        pr = rng.integers(0, len(refs), size=m).astype(np.int32)
        res, pool = eng.align(mat, 5, gapO, gapE, flag=flag, filters=0, filterd=filterd, mask_len=mask, score_size=ss, pair_query=pq, pair_ref=pr)
    exp, exp_pool, _, _, _ = C.cpu_batch(qs, refs, pq, pr, mat, 5, gapO, gapE, flag=flag, filters=0, filterd=filterd, mask_len=mask, score_size=ss, threads=2)
    d = C.compare_records(res, pool, exp, exp_pool)
    pairs_checked += len(pq)
    cigar_words += int(len(exp_pool))
    if d:
        bad += len(d)
        print("MISMATCH batch", b, "seed", seed, dict(flag=flag, ss=ss, gapO=gapO, gapE=gapE, mask=mask, filterd=filterd, ql=[len(q) for q in qs]), d[:3], flush=True)
print({"batches": n_batches, "pairs": pairs_checked, "cigar_words": cigar_words, "seed": seed, "mismatches": bad})
