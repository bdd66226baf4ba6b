#!/usr/bin/env python
"""Regenerate the golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and the checkers built
by `make -C oracle`).  It freezes, as small fixtures that travel to the GPU box:

  goldens.json        inputs (sequence strings / codes, scoring, flags) and the
                      reference's answers for
                        * the reference's shipped regression golden demo/old.txt
                          (== demo/new.txt): 100 x 54 bp vs demo/1M.fa, flag 0;
                        * the README sample (-c -s on demo/1k.fa);
                        * BASELINE config 1 (demo/r1.fa x demo/r1_query.fq), flags 0 and 2;
                        * the protein / repeat / fastq demo pairs;
                        * the example.c / example.cpp 15-mer known answer.
  demo_1M_ref.npz     the encoded 1,000,001-code reference of demo/1M.fa
                      (codes 0..4, compressed) -- the only large input.

Expected values come from two independent sources that must agree here:
the text the reference ships in demo/old.txt, and a live run of
oracle/_ref/libssw_ref.so on the same encoded inputs.
"""
import gzip
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from common import (BLOSUM50, cigar_string, dna_matrix, dna_matrix_cpp, encode_aa, encode_dna,  # noqa: E402
                    load_ref)

DEMO = "/root/reference/demo"


def read_fastx(path):
    """Minimal FASTA/FASTQ reader -> list of (name, sequence)."""
    op = gzip.open if path.endswith(".gz") else open
    out = []
    with op(path, "rt") as f:
        lines = [l.rstrip("\n") for l in f]
    i = 0
    while i < len(lines):
        l = lines[i]
        if l.startswith(">"):
            name = l[1:].split()[0]
            seq = []
            i += 1
            while i < len(lines) and not lines[i].startswith(">"):
                seq.append(lines[i].strip())
                i += 1
            out.append((name, "".join(seq)))
        elif l.startswith("@"):
            name = l[1:].split()[0]
            out.append((name, lines[i + 1].strip()))
            i += 4
        else:
            i += 1
    return out


def run_cases(lib, enc, mat, n, refs, reads, flag, gapO=3, gapE=1, filters=0, filterd=0, score_size=2,
              maskLen=None, mark=False):
    """The CLI's double loop (main.c:462-532): every read against every reference."""
    exp = []
    for _, q in reads:
        qn = enc(q)
        for _, r in refs:
            rn = enc(r)
            ml = len(qn) // 2 if maskLen is None else maskLen        # main.c:465
            res = lib.align(qn, rn, mat, n, gapO, gapE, flag, filters, filterd, ml, score_size, mark=mark)
            exp.append(res)
    return exp


def main():
    lib = load_ref()
    cases = []
    dna = dna_matrix(2, 2)

    # ---- 1. the reference's shipped golden: demo/old.txt ----------------------------------
    refs = read_fastx(f"{DEMO}/1M.fa")
    reads = read_fastx(f"{DEMO}/54mer_hap1_1.100.fa")
    txt = open(f"{DEMO}/old.txt").read()
    assert txt == open(f"{DEMO}/new.txt").read()
    shipped = []
    for blk in txt.strip().split("\n\n"):
        m = re.search(r"optimal_alignment_score: (\d+)\t(?:suboptimal_alignment_score: (\d+)\t)?strand: \+\t"
                      r"target_end: (\d+)\tquery_end: (\d+)", blk)
        shipped.append((int(m.group(1)), int(m.group(2) or 0), int(m.group(3)) - 1, int(m.group(4)) - 1))
    assert len(shipped) == 100
    live = run_cases(lib, encode_dna, dna, 5, refs, reads, flag=0)
    for s, l in zip(shipped, live):
        assert s == (l["score1"], l["score2"], l["ref_end1"], l["read_end1"]), (s, l)
    np.savez_compressed(os.path.join(HERE, "demo_1M_ref.npz"), ref=encode_dna(refs[0][1]))
    cases.append(dict(name="demo_old_txt_1M_x_54mer", source="demo/old.txt == demo/new.txt + live libssw_ref.so",
                      alphabet="dna", matrix="dna_2_2", n=5, gapO=3, gapE=1, flag=0, filters=0, filterd=0,
                      score_size=2, maskLen=None, refs_npz="demo_1M_ref.npz",
                      reads=[q for _, q in reads], expected=live))

    # ---- 2. README sample + small demo pairs ---------------------------------------------
    def add(name, ref_file, read_file, flag, alphabet="dna", mat=dna, n=5, matrix="dna_2_2", mark=True):
        enc = encode_dna if alphabet == "dna" else encode_aa
        rf = read_fastx(f"{DEMO}/{ref_file}")
        rd = read_fastx(f"{DEMO}/{read_file}")
        exp = run_cases(lib, enc, mat, n, rf, rd, flag=flag, mark=mark and flag != 0)
        cases.append(dict(name=name, source=f"live libssw_ref.so on demo/{ref_file} x demo/{read_file}",
                          alphabet=alphabet, matrix=matrix, n=n, gapO=3, gapE=1, flag=flag, filters=0, filterd=0,
                          score_size=2, maskLen=None, refs=[r for _, r in rf], reads=[q for _, q in rd],
                          expected=exp))
        return exp

    e = add("readme_1k_x_54mer_c", "1k.fa", "54mer_hap1_1.100.fa", 2)
    # README.md:113-121 first record: POS 453, AS 37, NM 11, ZS 28, 453-492 / 17-51
    assert (e[0]["ref_begin1"] + 1, e[0]["score1"], e[0]["nm"], e[0]["score2"]) == (453, 37, 11, 28), e[0]
    assert (e[0]["ref_end1"] + 1, e[0]["read_begin1"] + 1, e[0]["read_end1"] + 1) == (492, 17, 51)
    e = add("config1_r1_flag0", "r1.fa", "r1_query.fq", 0)
    assert (e[0]["score1"], e[0]["score2"], e[0]["ref_end1"] + 1, e[0]["read_end1"] + 1) == (52, 40, 192, 94)
    e = add("config1_r1_c", "r1.fa", "r1_query.fq", 2)
    assert (e[0]["ref_begin1"] + 1, e[0]["read_begin1"] + 1) == (103, 1) and e[0]["nm"] == 38
    e = add("protein2_x_protein1_c", "protein2.fa", "protein1.fa", 2, alphabet="protein", mat=BLOSUM50, n=24,
            matrix="blosum50")
    assert (e[0]["score1"], e[0]["score2"]) == (218, 120)
    e = add("pRef_x_pRead_c", "pRef.fa", "pRead.fa", 2)
    assert (e[0]["score1"], e[0]["score2"], e[0]["ref_begin1"] + 1, e[0]["ref_end1"] + 1) == (40, 18, 4, 41)
    e = add("target_fastq_x_query_fastq_c", "target.fastq", "query.fastq", 2)
    assert cigar_string(e[0]["cigar_marked"]) == "37S8=9S"

    # ---- 3. example.c / example.cpp known answer ---------------------------------------------
    ref_s = "CAGCCTTTCTGACCCGGAAATCAAAATAGGCACAACAAA"       # example.c:41 / example.cpp:20
    read_s = "CTGAGCCGGTAAATC"
    for nm, flag, mat, mname in (("example_c_flag1", 1, dna, "dna_2_2"),
                                 ("example_cpp_flag0f", 0x0f, dna_matrix_cpp(2, 2), "dna_cpp_2_2")):
        res = lib.align(encode_dna(read_s), encode_dna(ref_s), mat, 5, 3, 1, flag, 0, 32767, 15, 2, mark=True)
        assert (res["score1"], res["score2"], res["ref_begin1"], res["ref_end1"], res["read_begin1"],
                res["read_end1"], res["ref_end2"], res["nm"]) == (21, 8, 8, 21, 0, 14, 4, 2), res
        assert cigar_string(res["cigar_marked"]) == "4=1X4=1I5="
        cases.append(dict(name=nm, source="example.c:41-147 / example.cpp:20-36 known answer + live libssw_ref.so",
                          alphabet="dna", matrix=mname, n=5, gapO=3, gapE=1, flag=flag, filters=0, filterd=32767,
                          score_size=2, maskLen=15, refs=[ref_s], reads=[read_s], expected=[res]))

    with open(os.path.join(HERE, "goldens.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print("wrote", len(cases), "cases,", sum(len(c["expected"]) for c in cases), "alignments")


if __name__ == "__main__":
    main()
