#!/bin/bash
# Round-2 ncu evidence (one GPU; everything lands in gpurun_out/, summaries are copied to profiles/ by hand).
mkdir -p gpurun_out
NCU="ncu --clock-control none"
echo "== full capture: config-2 fill (block column maxima)"
$NCU --set full --import-source on -k regex:ssw_fill_kernel -c 1 -f -o gpurun_out/ncu_fill_cfg2_r2 python tools/run_config.py 2 --reps 1 > gpurun_out/ncu_a.log 2>&1
echo "== full capture: config-4 fill (16,20), eight warps per CTA"
$NCU --set full --import-source on -k regex:ssw_fill_kernel -c 1 -f -o gpurun_out/ncu_fill_cfg4_r2 python tools/run_config.py 4 --reps 1 > gpurun_out/ncu_b.log 2>&1
echo "== full capture: config-5 forward strips"
$NCU --set full --import-source on -k regex:ssw_fill_strips_kernel -c 1 -f -o gpurun_out/ncu_strips_cfg5_r2 python tools/run_config.py 5 --reps 1 --reads 592 --opt slices=1 > gpurun_out/ncu_c.log 2>&1
for r in ncu_fill_cfg2_r2 ncu_fill_cfg4_r2 ncu_strips_cfg5_r2; do
  (python tools/ncu_summary.py gpurun_out/$r.ncu-rep; python tools/sass_profile.py gpurun_out/$r.ncu-rep) > gpurun_out/$r.txt 2>&1
done
echo "== dram traffic of the 100,000-read launch (config 3)"
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:ssw_fill_kernel -c 1 --csv --log-file gpurun_out/traffic_fill_cfg3_r2.csv python tools/run_config.py 3 --reads 100000 --reps 1 > gpurun_out/ncu_d.log 2>&1
echo "== launch list of the bench command"
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --e2e-reps 1 --no-cpu-baseline > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep | head
