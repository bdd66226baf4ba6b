#!/bin/bash
mkdir -p gpurun_out
python tools/sweep_cfg5.py "" "slices=1" "slices=2" "slices=4" "slices=5" "slices=6" "slices=8" \
  "slices=3,slice_taper=70" "slices=4,slice_taper=70" "slices=5,slice_taper=70" "slices=6,slice_taper=70" \
  "slices=4,slice_taper=50" "slices=5,slice_taper=50" "slices=6,slice_taper=80" "slices=8,slice_taper=80" \
  "slices=3,slice_prio=1" "slices=5,slice_prio=1" "slices=5,slice_taper=70,slice_prio=1" \
  "slices=3,tail_spec=1" "slices=5,slice_taper=70,tail_spec=1" "slices=5,slice_taper=50,tail_spec=1,slice_prio=1" \
  "super=1024" "super=2048" "super=8192" "slices=1,tb_spec=0" "" > gpurun_out/sweep5.txt 2> gpurun_out/sweep5.err
cat gpurun_out/sweep5.txt
SSW_TRACE=1 python tools/sweep_cfg5.py --reps 1 "slices=1" 2>&1 | grep -i "traceback round\|opts" | tail -30 > gpurun_out/sweep5_trace.txt
cat gpurun_out/sweep5_trace.txt
timeout 600 python -m pytest tests/test_gpu_parity_full.py -x -q -s -k "config5" 2>&1 | grep -v "^When maskLen\|^Please set\|^Warning: The align" | tail -4
