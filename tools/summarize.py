#!/usr/bin/env python
"""Print the key numbers of bench.py JSON lines (one file per run)."""
import json, sys
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable:", e); continue
    r = d.get("roofline", {}); a = d.get("alu_roofline", {})
    print("%-28s value %7.0f GCUPS  step %7.1f ms  fill %7.1f ms  kernel-only %7.0f GCUPS  e2e %7.0f  launches %s  clocks %s" % (
        p.split("/")[-1], d["value"], d["ms_per_step"], r.get("kernel_ms_per_step", 0), a.get("achieved_gcups", 0),
        d.get("e2e", {}).get("value", 0), d.get("gpu_launches"), (d.get("clocks") or {}).get("sm_mhz")))
