"""Host logic of ssw_batch_cli (FASTA/FASTQ parsing, options, strand choice, BLAST-like and SAM formatting) checked on
CPU: the driver source is linked against tests/cli_shim/shim.c, which answers the batch API with the oracle, and its
stdout is compared with the frozen outputs of the reference's ssw_test (tests/golden/consumer_outputs.json).
The GPU suite runs the same comparison with the real binary over libssw.so (test_gpu_parity.py)."""
import json
import os
import subprocess

import pytest

import common as C


@pytest.fixture(scope="module")
def cli(tmp_path_factory):
    C.build_oracle() if hasattr(C, "build_oracle") else subprocess.run(["make", "-s", "-C", C.ORACLE_DIR, "all"], check=True)
    out = str(tmp_path_factory.mktemp("cli") / "ssw_batch_cli_cpu")
    src = os.path.join(C.PKG, "csrc", "ssw_batch_cli.cpp")
    shim = os.path.join(C.ROOT, "tests", "cli_shim", "shim.c")
    obj = out + "_shim.o"
    subprocess.run(["gcc", "-O2", "-c", "-o", obj, shim], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", out, src, obj, "-L" + C.ORACLE_DIR, "-lssw_oracle",
                    "-Wl,-rpath," + C.ORACLE_DIR, "-lz", "-lm"], check=True)
    return out


def test_cli_reproduces_reference_driver_output(cli, tmp_path):
    with open(os.path.join(C.GOLDEN, "consumer_outputs.json")) as f:
        G = json.load(f)
    for name, seqs in G["files"].items():
        (tmp_path / name).write_text(seqs)
    n = 0
    for run in G["runs"]:
        if run["exe"] != "ssw_test":
            continue
        out = subprocess.run([cli] + run["args"], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
        assert out.returncode == 0, out.stderr[-500:]
        assert "\n".join(out.stdout.splitlines()) == run["stdout"], run["args"]
        n += 1
    assert n >= 10


def test_cli_gzip_and_multiline_input(cli, tmp_path):
    """gzip-compressed and line-wrapped inputs parse to the same records as the plain files"""
    import gzip
    with open(os.path.join(C.GOLDEN, "consumer_outputs.json")) as f:
        G = json.load(f)
    ref = G["files"]["1k.fa"]
    name, seq = ref.split("\n", 1)
    seq = seq.replace("\n", "")
    wrapped = name + "\n" + "\n".join(seq[i:i + 70] for i in range(0, len(seq), 70)) + "\n"
    (tmp_path / "1k.fa").write_text(ref)
    (tmp_path / "wrapped.fa").write_text(wrapped)
    (tmp_path / "q.fa").write_text(G["files"]["54mer_hap1_1.100.fa"])
    with gzip.open(tmp_path / "q.fa.gz", "wt") as f:
        f.write(G["files"]["54mer_hap1_1.100.fa"])
    a = subprocess.run([cli, "-c", "1k.fa", "q.fa"], capture_output=True, text=True, cwd=str(tmp_path)).stdout
    b = subprocess.run([cli, "-c", "wrapped.fa", "q.fa.gz"], capture_output=True, text=True, cwd=str(tmp_path)).stdout
    assert a == b and len(a) > 1000


def test_cli_usage_and_missing_file(cli, tmp_path):
    assert subprocess.run([cli], capture_output=True).returncode == 1
    assert subprocess.run([cli, "nope.fa", "nope2.fa"], capture_output=True, cwd=str(tmp_path)).returncode == 1
