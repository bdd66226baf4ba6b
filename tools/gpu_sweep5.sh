#!/bin/bash
mkdir -p gpurun_out
python tools/sweep_cfg5.py --reps 4 "" "strip_warps=8" "strip_warps=8,parts=1" "strip_warps=8,parts=4" "strip_warps=12" "strip_warps=8,slices=6,slice_taper=80" "strip_warps=12,slices=6,slice_taper=80" "slices=6,slice_taper=80" "slices=1" "slices=1,strip_warps=8" "slices=1,strip_warps=8,parts=4" 2>&1 | grep -i "opts" > gpurun_out/sweep5g.txt
for lib in libssw_tbw1.so libssw_tbw2.so; do python tools/sweep_cfg5.py --lib $lib --reps 4 "" "slices=1" "slices=6,slice_taper=80" "strip_warps=8" 2>&1 | grep -i "opts"; done >> gpurun_out/sweep5g.txt
cat gpurun_out/sweep5g.txt
