#!/bin/bash
# Round-2 measurement bundle for one gpurun call (everything lands in gpurun_out/).
mkdir -p gpurun_out
L=complete-striped-smith-waterman-library_b200/libssw.so
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^When maskLen\|^Please set\|^Warning: The align" > gpurun_out/gputest_r2.log; tail -4 gpurun_out/gputest_r2.log
echo "== bench (short)"; timeout 500 python bench.py --steps 1 --warmup 1 --e2e-reps 1 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo rc=$?
echo "== config 4 A/B"; (python tools/run_config.py 4 --reps 3; python tools/run_config.py 4 --reps 3 --opt inst=13) > gpurun_out/cfg4_ab.txt 2>&1
echo "== config 5"; python tools/run_config.py 5 --reps 3 > gpurun_out/cfg5.txt 2>&1
echo "== call bench"; (for args in "1000 150 0" "1000 150 2" "100000 150 0" "2000000 150 0"; do ./tools/call_bench $L 400 8 $args; ./tools/call_bench oracle/_ref/libssw_ref.so 400 8 $args; done) > gpurun_out/call_bench.txt 2>&1
echo "== cli timing"; python tools/cli_timing.py 3 > gpurun_out/cli_timing.json 2> gpurun_out/cli_timing.err
echo "== ncu traffic config 2"; timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ssw_fill_kernel -c 2 --csv --log-file gpurun_out/traffic_fill_r2.csv python tools/run_config.py 2 --reps 1 > gpurun_out/ncu_traffic.log 2>&1; echo rc=$?
