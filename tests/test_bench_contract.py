"""bench.py's output contract, checked on the CPU-only arm (`--impl reference` times the reference's own SSE2 path on
the host cores): exactly one line on stdout, a JSON object with the keys the driver reads."""
import json
import os
import subprocess
import sys

import common as C


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(C.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "2", "--ref-len", "400000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-800:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "GCUPS" and d["value"] > 0 and d["higher_is_better"] is True
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert "workload" in d["config"]
