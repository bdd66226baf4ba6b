"""GPU parity of the device groups (include/ssw_batch.h: ssw_group_*): one batch cut over several engines, each driven by its
own host thread.  On a one-GPU box the group holds two engines on device 0 (the threads, the cut, the scatter of the
records and the CIGAR concatenation are the same code); with more GPUs visible the group spans them.  Every record is
compared with the CPU checker and with one engine's answer.  Nothing here reads /root/reference."""
import os
import subprocess

import numpy as np
import pytest

import common as C
from test_gpu_parity import _pkg
from test_device_group_host import run_group_cases

pytestmark = pytest.mark.gpu


def _devices(L):
    import ctypes as ct
    lib = ct.CDLL(C.LIB_OURS)
    lib.ssw_device_count.restype = ct.c_int32
    n = int(lib.ssw_device_count())
    assert n >= 1
    return list(range(min(n, 4))) if n > 1 else [0, 0]


def test_group_config2_slice_and_protein_grid():
    L = _pkg()
    devs = _devices(L)
    grp = L.GroupAligner(devices=devs)
    assert grp.size == len(devs)
    one = L.BatchAligner(device=0)
    run_group_cases(grp, one, ref_len=300_000, n_reads=96, n_queries=24, n_targets=500, n_check=600, long_ref=40_000, long_len=3000)
    grp.close()
    one.close()


def test_batch_cli_over_a_group(tmp_path):
    """ssw_batch_cli -g N (N engines' worth of GPUs; here what the box has) prints what -g 1 prints."""
    import json
    cli = os.path.join(C.PKG, "ssw_batch_cli")
    assert os.path.exists(cli)
    with open(os.path.join(C.GOLDEN, "consumer_outputs.json")) as f:
        G = json.load(f)
    for name, text in G["files"].items():
        (tmp_path / name).write_text(text)
    for run in G["runs"]:
        if run["exe"] != "ssw_test" or "1k.fa" not in run["args"] or "-r" not in run["args"]:
            continue
        out = subprocess.run([cli, "-g", "0"] + run["args"], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
        assert out.returncode == 0, out.stderr[-500:]
        assert "\n".join(out.stdout.splitlines()) == run["stdout"], run["args"]
