#!/bin/bash
mkdir -p gpurun_out
SSW_TRACE=1 python tools/sweep_cfg5.py --reps 4 "slices=1" "slices=1,tb_spec=1" "slices=1,tb_spec=33" "slices=1,tb_spec=65" "slices=1,tb_spec=257" "slices=1,tb_spec=513" "" "tb_spec=1" "tb_spec=65" "tb_spec=257" 2>&1 | grep -i "opts\|round: 1000" | cut -c1-420 > gpurun_out/sweep5h.txt
awk '/opts/ || !seen[$0]++' gpurun_out/sweep5h.txt | grep -v "^\[libssw" ; grep "^\[libssw" gpurun_out/sweep5h.txt | sort | uniq -c | sort -rn | head -12
