#!/bin/bash
# Re-entry check of round 2 (one GPU): the driver's test command, smoke, and the config 4 / 5 phase timings.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -v "^When maskLen\|^Please set\|^Warning: The align" > gpurun_out/gputest_r2b.log; tail -3 gpurun_out/gputest_r2b.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== config 5"; python tools/run_config.py 5 --reps 3 > gpurun_out/cfg5_r2b.txt 2>&1; tail -5 gpurun_out/cfg5_r2b.txt
echo "== config 4"; python tools/run_config.py 4 --reps 3 > gpurun_out/cfg4_r2b.txt 2>&1; tail -5 gpurun_out/cfg4_r2b.txt
