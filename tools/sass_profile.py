#!/usr/bin/env python
"""Where do the executed instructions of a profiled kernel go?  Reads `ncu -i X.ncu-rep --page source --csv` (SASS
view) and prints (a) the opcode histogram weighted by executed warp instructions, (b) the share of the hottest
contiguous region (the unrolled step bodies: instructions executed >= 50 % of the most executed one), (c) the rest
split into 'per-step overhead' and 'rare paths'.   python tools/sass_profile.py report.ncu-rep [top-N]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ci, cs, cn = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
ins = [(r[cs].strip(), int(r[ci]), int(r[cn])) for r in rows[2:] if len(r) > ci and r[ci].isdigit()]
tot = sum(x[1] for x in ins)
mx = max(x[1] for x in ins)
def opcode(s):
    p = s.split()
    if p and p[0].startswith("@"):
        p = p[1:]
    return p[0].rstrip(";") if p else "?"
hist = Counter()
for s, n, _ in ins:
    hist[opcode(s)] += n
print("kernel:", rows[0][1][:100])
print("executed warp instructions: %.3e, static instructions %d" % (tot, len(ins)))
print("opcode shares:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in hist.most_common(top)))
hot = sum(n for _, n, _ in ins if n >= mx * 0.5)
warm = sum(n for _, n, _ in ins if mx * 0.02 <= n < mx * 0.5)
cold = tot - hot - warm
print("executed >= 50%% of the hottest instruction (step bodies): %.1f%%; 2-50%% (per-item / per-event code): %.1f%%; < 2%%: %.1f%%"
      % (100.0 * hot / tot, 100.0 * warm / tot, 100.0 * cold / tot))
dpx = sum(v for k, v in hist.items() if k.startswith(("VIADDMNMX", "VIMNMX", "VIADD")))
print("DPX/packed-16 instructions: %.1f%% of all executed" % (100.0 * dpx / tot))
