#!/usr/bin/env python
"""Freeze the stdout of the reference's own drivers (ssw_test, example_c, example_cpp built from the UNMODIFIED
reference against its own ssw.c: oracle/_ref/*_ref) on small demo inputs.  The same unmodified drivers linked
against OUR libssw.so (oracle/_ref/*_b200) must print the same text on the GPU box (test_gpu_parity.py).
Build container only: reads /root/reference/demo."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref")
DEMO = "/root/reference/demo"

files = {}
for name in ("r1.fa", "r1_query.fq", "1k.fa", "54mer_hap1_1.100.fa", "protein1.fa", "protein2.fa", "pRef.fa", "pRead.fa",
             "target.fastq", "query.fastq"):
    files[name] = open(os.path.join(DEMO, name)).read()

runs = []
for exe, args in [
    ("ssw_test", ["r1.fa", "r1_query.fq"]),                      # BASELINE config 1
    ("ssw_test", ["-c", "r1.fa", "r1_query.fq"]),
    ("ssw_test", ["-c", "-s", "r1.fa", "r1_query.fq"]),
    ("ssw_test", ["-c", "-s", "-h", "1k.fa", "54mer_hap1_1.100.fa"]),   # README sample
    ("ssw_test", ["-c", "-r", "1k.fa", "54mer_hap1_1.100.fa"]),
    ("ssw_test", ["-p", "-c", "protein2.fa", "protein1.fa"]),
    ("ssw_test", ["-c", "pRef.fa", "pRead.fa"]),
    ("ssw_test", ["-c", "-s", "target.fastq", "query.fastq"]),
    ("ssw_test", ["-m", "1", "-x", "3", "-o", "5", "-e", "2", "-c", "1k.fa", "54mer_hap1_1.100.fa"]),
    ("example_c", []),
    ("example_cpp", []),
]:
    a = [x if not x.endswith((".fa", ".fq", ".fastq")) else os.path.join(DEMO, x) for x in args]
    out = subprocess.run([os.path.join(BIN, exe + "_ref")] + a, capture_output=True, text=True, check=True).stdout
    out = "\n".join(l for l in out.splitlines() if not l.startswith("CPU time"))
    runs.append({"exe": exe, "args": args, "stdout": out})

with open(os.path.join(HERE, "consumer_outputs.json"), "w") as f:
    json.dump({"files": files, "runs": runs}, f, indent=0)
print("froze", len(runs), "driver runs,", sum(len(r["stdout"]) for r in runs), "bytes of stdout")
