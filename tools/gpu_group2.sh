#!/bin/bash
# Device groups over two real GPUs of one process: config 2 at group sizes 1 and 2, then the group parity test.
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 40 python tools/group_scaling.py --sizes 1,2 --reps 3 2>&1 | grep -v "^reference tree" | tee gpurun_out/group_scaling_n2.txt
timeout 30 python -m pytest tests/test_gpu_parity_group.py -x -q -s 2>&1 | grep -v "^When maskLen\|^Please set\|^Warning: The align" | tail -4 | tee -a gpurun_out/group_scaling_n2.txt
