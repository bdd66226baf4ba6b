#!/usr/bin/env python
"""Print the opcode histogram (and the first lines) of one unrolled step of a fill kernel from `cuobjdump -sass` output."""
import re, sys
from collections import Counter
txt = open(sys.argv[1]).read()
key = sys.argv[2]
i = txt.index("Function : " + key)
j = txt.find("Function :", i + 10)
body = txt[i:j if j > 0 else None]
lines = [l for l in body.splitlines() if re.match(r'\s+/\*[0-9a-f]{4}\*/', l)]
ins = [re.sub(r'^\s+/\*[0-9a-f]{4}\*/\s+', '', l).split('/*')[0].strip() for l in lines]
idx = [k for k, x in enumerate(ins) if x.startswith('SHFL.UP')]
seg = ins[idx[3] - 2: idx[6] - 2]
hot = []
for x in seg:
    hot.append(x)
    if x.startswith('@') and 'BRA' in x:
        break
c = Counter(x.split()[0] if not x.startswith('@') else x.split()[1] for x in hot)
print(len(hot), dict(c))
n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
print('\n'.join(hot[:n]))
