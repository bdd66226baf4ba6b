#!/bin/bash
# Device groups on hardware (one GPU: two engines on device 0), smoke, then as much of the parity suite as the time allows.
mkdir -p gpurun_out
echo "== group"; timeout 120 python -m pytest tests/test_gpu_parity_group.py -x -q -s 2>&1 | grep -v "^When maskLen\|^Please set\|^Warning: The align" | tail -6 | tee gpurun_out/group_check.log
echo "== smoke"; timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/group_check.log
echo "== parity"; timeout 100 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 | tee -a gpurun_out/group_check.log
