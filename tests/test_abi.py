"""CPU test: the shipped CUDA library loads and exports every symbol that include/*.h
declares (plus the two helpers the reference's libssw.so leaks).  No compute calls."""
import ctypes as ct
import os
import re
import subprocess

import numpy as np
import pytest

import common as C

DECLARED = ["ssw_init", "init_destroy", "ssw_align", "align_destroy", "mark_mismatch", "encoded_ops",
            "ssw_engine_create", "ssw_engine_destroy", "ssw_engine_device_name", "ssw_engine_set_sequences",
            "ssw_engine_align", "ssw_align_batch", "ssw_engine_last_timing", "ssw_engine_set_option",
            "ssw_engine_set_sequences_text", "ssw_align_batch_text", "ssw_engine_mark_mismatch", "ssw_align_batch_marked", "ssw_engine_set_sequences_packed",
            "ssw_device_count", "ssw_group_create", "ssw_group_destroy", "ssw_group_size", "ssw_group_engine", "ssw_group_align", "ssw_group_align_batch"]
LEAKED_BY_REFERENCE = ["add_cigar", "store_previous_m"]       # non-static in ssw.c:984,994


def _declared_in_headers():
    names = set()
    for h in ("ssw.h", "ssw_batch.h"):
        txt = open(os.path.join(C.ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        for m in re.finditer(r"\b([a-z_]+[a-z0-9_]*)\s*\(", txt):
            if not re.search(r"static\s+inline[^;{]*\b%s\s*\($" % m.group(1), txt[: m.end()].split(";")[-1]):
                names.add(m.group(1))
    return names


def test_headers_and_list_agree():
    found = _declared_in_headers()
    for s in DECLARED:
        if s != "encoded_ops":
            assert s in found, s


@pytest.mark.skipif(not os.path.exists(C.LIB_OURS), reason="libssw.so not built yet (python __graft_entry__.py)")
def test_library_exports_every_declared_symbol():
    lib = ct.CDLL(C.LIB_OURS)
    for s in DECLARED + LEAKED_BY_REFERENCE:
        assert hasattr(lib, s), "libssw.so does not export %s" % s
    ops = (ct.c_uint8 * 128).in_dll(lib, "encoded_ops")
    assert [ops[ord(c)] for c in "MIDNSHP=X"] == list(range(9))
    out = subprocess.run(["nm", "-D", "--defined-only", C.LIB_OURS], capture_output=True, text=True).stdout
    assert " T ssw_align" in out and " T ssw_engine_align" in out


@pytest.mark.skipif(not os.path.exists(C.LIB_OURS), reason="libssw.so not built yet")
def test_product_does_not_link_the_oracle():
    out = subprocess.run(["nm", "-D", C.LIB_OURS], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "cuemu" not in out


@pytest.mark.skipif(not os.path.exists(C.LIB_OURS), reason="libssw.so not built yet")
def test_library_exports_the_cpp_wrapper():
    """include/ssw_cpp.h: every public member of StripedSmithWaterman::Aligner is defined in libssw.so"""
    out = subprocess.run(["nm", "-DC", "--defined-only", C.LIB_OURS], capture_output=True, text=True).stdout
    for member in ("Aligner::Aligner()", "Aligner::SetReferenceSequence(char const*, unsigned long)",
                   "Aligner::SetReferenceSequence(char const*)", "Aligner::ClearReferenceSequence()",
                   "Aligner::SetGapPenalty(unsigned char, unsigned char)", "Aligner::Clear()", "Aligner::ReBuild()",
                   "Aligner::AlignBatch(", "Aligner::Align(char const*, unsigned long, char const*, unsigned long,",
                   "Aligner::Align(char const*, char const*,", "Aligner::Align(char const*, unsigned long, StripedSmithWaterman::Filter const&",
                   "Aligner::Align(char const*, StripedSmithWaterman::Filter const&"):
        assert "StripedSmithWaterman::" + member in out, member


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(not os.path.exists(C.LIB_OURS), reason="libssw.so not built yet")
@pytest.mark.skipif(_have_gpu(), reason="needs a host without a GPU")
def test_no_cpu_fallback_without_a_gpu(capfd, tmp_path):
    """Without a CUDA device the product refuses to work, loudly: no engine, ssw_align -> NULL, the CLI exits non-zero."""
    lib = ct.CDLL(C.LIB_OURS)
    lib.ssw_engine_create.restype = ct.c_void_p
    assert not lib.ssw_engine_create(-1)
    assert "no CPU compute path" in capfd.readouterr().err
    ours = C.load_ours()
    q = np.array([0, 1, 2, 3] * 10, dtype=np.int8)
    r = np.array([0, 1, 2, 3] * 30, dtype=np.int8)
    assert ours.align(q, r, C.dna_matrix(2, 2), 5, 3, 1, 0, 0, 0, 15, 2) is None
    cli = os.path.join(C.PKG, "ssw_batch_cli")
    if os.path.exists(cli):
        import json
        with open(os.path.join(C.GOLDEN, "consumer_outputs.json")) as f:
            files = json.load(f)["files"]
        (tmp_path / "r1.fa").write_text(files["r1.fa"])
        (tmp_path / "q.fq").write_text(files["r1_query.fq"])
        out = subprocess.run([cli, "r1.fa", "q.fq"], capture_output=True, text=True, cwd=str(tmp_path))
        assert out.returncode != 0 and out.stdout == ""
