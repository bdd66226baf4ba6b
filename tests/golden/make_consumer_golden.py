#!/usr/bin/env python
"""Freeze the stdout of the reference's own drivers (ssw_test, example_c, example_cpp built from the UNMODIFIED
reference against its own ssw.c: oracle/_ref/*_ref) on small demo inputs.  The same unmodified drivers linked
against OUR libssw.so (oracle/_ref/*_b200) must print the same text on the GPU box (test_gpu_parity.py).
Build container only: reads /root/reference/demo."""
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref")
DEMO = "/root/reference/demo"

files = {}
for name in ("r1.fa", "r1_query.fq", "1k.fa", "54mer_hap1_1.100.fa", "protein1.fa", "protein2.fa", "pRef.fa", "pRead.fa",
             "target.fastq", "query.fastq", "54mer_hap1_1.100.fastq"):
    files[name] = open(os.path.join(DEMO, name)).read()
# short name: the reference's option parser also scans the characters of an option VALUE as flags (main.c:253-303)
files["b62.txt"] = open(os.path.join(DEMO, "blosum62.txt")).read()

TMP = tempfile.mkdtemp()
for name, text in files.items():
    with open(os.path.join(TMP, name), "w") as f:
        f.write(text)

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
    ("ssw_test", ["-a", "b62.txt", "-c", "protein2.fa", "protein1.fa"]),        # weight matrix file
    ("ssw_test", ["-c", "-s", "-r", "-h", "1k.fa", "54mer_hap1_1.100.fastq"]),        # SAM, both strands, qualities
    ("ssw_test", ["-c", "-r", "-f", "60", "1k.fa", "54mer_hap1_1.100.fa"]),           # score filter
    ("ssw_test", ["-r", "1k.fa", "54mer_hap1_1.100.fa"]),                             # scores only, both strands
    ("example_c", []),
    ("example_cpp", []),
]:
    out = subprocess.run([os.path.join(BIN, exe + "_ref")] + args, capture_output=True, text=True, check=True, cwd=TMP).stdout
    out = "\n".join(l for l in out.splitlines() if not l.startswith("CPU time"))
    runs.append({"exe": exe, "args": args, "stdout": out})

with open(os.path.join(HERE, "consumer_outputs.json"), "w") as f:
    json.dump({"files": files, "runs": runs}, f, indent=0)
print("froze", len(runs), "driver runs,", sum(len(r["stdout"]) for r in runs), "bytes of stdout")
