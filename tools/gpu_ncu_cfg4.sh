#!/bin/bash
# ncu evidence of the (16,19) config-4 fill kernel (one GPU).
mkdir -p gpurun_out
timeout 400 ncu --clock-control none --set full --import-source on -k regex:ssw_fill_kernel -s 1 -c 1 -f -o gpurun_out/ncu_fill_cfg4_16x19_r2 python tools/run_config.py 4 --reps 1 > gpurun_out/ncu_b2.log 2>&1; echo rc=$?
(python tools/ncu_summary.py gpurun_out/ncu_fill_cfg4_16x19_r2.ncu-rep; python tools/sass_profile.py gpurun_out/ncu_fill_cfg4_16x19_r2.ncu-rep) > gpurun_out/ncu_fill_cfg4_16x19_r2.txt 2>&1
cut -c1-160 gpurun_out/ncu_fill_cfg4_16x19_r2.txt | sed -n 1,14p; tail -4 gpurun_out/ncu_fill_cfg4_16x19_r2.txt | cut -c1-300
