#!/usr/bin/env python
"""Random batches through the device groups (ssw_group_align, csrc/ssw_group.cpp) on the CPU emulator with 1 .. 5 emulated devices
against ONE engine's answer for the same pairs: codes or text (+ reverse complements), full grids, grid prefixes and explicit pair
lists (incl. minus-strand indices), all flags, marked CIGARs + NM.   python tools/fuzz_group_emulator.py [n_cases] [seed]"""
import os, subprocess, sys
import numpy as np
os.environ["SSW_EMU_DEVICES"] = "5"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C
from test_emulated_kernels import EMU_DIR, _pkg

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
subprocess.run(["make", "-s", "-C", EMU_DIR], check=True, stdout=sys.stderr)
L = _pkg()
one = L.BatchAligner(lib_dir=EMU_DIR, lib_name="libssw_emu.so")
one.set_option("latency_cols", 0)
groups = {}
rng = np.random.default_rng(seed)
table = np.full(128, 4, dtype=np.int8)
for i, c in enumerate("ACGT"):
    table[ord(c)] = i
    table[ord(c.lower())] = i
FIELDS = ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "flag", "cigar_len", "status")
bad = 0
pairs = 0
for case in range(n_cases):
    world = int(rng.integers(1, 6))
    if world not in groups:
        groups[world] = L.GroupAligner(n_devices=world, lib_dir=EMU_DIR, lib_name="libssw_emu.so")
        groups[world].set_option("latency_cols", 0)
    grp = groups[world]
    text = rng.random() < 0.5
    rc = text and rng.random() < 0.6
    refs = [rng.integers(0, 4, size=int(rng.integers(40, 500))).astype(np.int8) for _ in range(int(rng.integers(1, 4)))]
    qs = []
    for _ in range(int(rng.integers(1, 8))):
        r = refs[int(rng.integers(0, len(refs)))]
        ql = int(min(rng.integers(8, 200), len(r) - 5))
        q = C.mutate_read(rng, r, int(rng.integers(0, len(r) - ql)), ql, 0.08, 0.02, 0.02)
        if rc and rng.random() < 0.5:
            q = (3 - q[::-1]).astype(np.int8)                   # a minus-strand read (codes 0..3)
        qs.append(q)
    flag = int(rng.choice([0, 1, 2, 8, 0x0f]))
    marked = bool(flag & 7) and rng.random() < 0.5
    ss = int(rng.choice([0, 1, 2, 2]))
    mask = int(rng.choice([15, 30]))
    mat = C.dna_matrix(2, int(rng.integers(1, 4)))
    nq_all = len(qs) * (2 if rc else 1)
    mode = rng.random()
    pq = pr = None
    n_pairs = None
    if mode < 0.4:
        m = int(rng.integers(1, 14))
        pq = rng.integers(0, nq_all, size=m).astype(np.int32)
        pr = rng.integers(0, len(refs), size=m).astype(np.int32)
    elif mode < 0.55:
        n_pairs = int(rng.integers(1, nq_all * len(refs) + 1))       # a prefix of the grid
    if text:
        tq = ["".join("ACGT"[c] for c in q) for q in qs]
        tr = ["".join("ACGT"[c] for c in r) for r in refs]
        one.set_sequences_text(tq, tr, table, 5, add_reverse_complement=rc)
        kw = dict(table=table, add_reverse_complement=rc)
        gq, gr = tq, tr
    else:
        one.set_sequences(qs, refs)
        kw = {}
        gq, gr = qs, refs
    if pq is None and n_pairs is not None:
        bq = (np.arange(n_pairs) // len(refs)).astype(np.int32); br = (np.arange(n_pairs) % len(refs)).astype(np.int32)
    else:
        bq, br = pq, pr
    base, bpool = one.align(mat, 5, 3, 1, flag=flag, filters=int(rng.integers(0, 30)) if False else 0, filterd=32767, mask_len=mask, score_size=ss, pair_query=bq, pair_ref=br)
    bnm = None
    if marked:
        base, bpool, bnm = one.mark_mismatch(base, bpool, pair_query=bq, pair_ref=br)
    out = grp.align(gq, gr, mat, 5, 3, 1, flag=flag, filterd=32767, mask_len=mask, score_size=ss, pair_query=pq, pair_ref=pr, marked=marked, n_pairs=n_pairs, **kw)
    res, pool = out[0], out[1]
    ok = len(res) == len(base)
    for i in range(len(base)):
        if not ok:
            break
        a, b = res[i], base[i]
        ok = all(int(a[k]) == int(b[k]) for k in FIELDS)
        if ok and a["cigar_len"] > 0:
            ok = list(pool[a["cigar_off"]: a["cigar_off"] + a["cigar_len"]]) == list(bpool[b["cigar_off"]: b["cigar_off"] + b["cigar_len"]])
    if ok and marked:
        ok = list(out[2]) == list(bnm)
    pairs += len(base)
    if not ok:
        bad += 1
        print("MISMATCH case", case, "seed", seed, dict(world=world, text=text, rc=rc, flag=flag, marked=marked, ss=ss, mode=round(mode, 2), nq=len(qs), nr=len(refs)), flush=True)
print({"cases": n_cases, "pairs": pairs, "seed": seed, "mismatches": bad})
