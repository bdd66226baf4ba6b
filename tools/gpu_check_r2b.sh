#!/bin/bash
# Check of the (16,19) forward instance (one GPU): the driver's test command, config 4 A/B, a short bench line.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -v "^When maskLen\|^Please set\|^Warning: The align" > gpurun_out/gputest_r2c.log; tail -3 gpurun_out/gputest_r2c.log; grep "parity config4" gpurun_out/gputest_r2c.log | cut -c1-200
echo "== config 4: (16,19) automatic, then (16,20) forced"; (python tools/run_config.py 4 --reps 3; python tools/run_config.py 4 --reps 3 --opt inst=7) > gpurun_out/cfg4_r2c.txt 2>&1; grep -o '"rep": [0-9], "pairs": [0-9]*, "wall_ms": [0-9.]*\|"fill_forward_ms": [0-9.]*\|"byte_overflows": [0-9]*' gpurun_out/cfg4_r2c.txt | paste - - - 
echo "== short bench"; timeout 600 python bench.py --steps 2 --warmup 3 --e2e-reps 1 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo rc=$?; python tools/summarize.py gpurun_out/bench_r2c.json | tail -3
