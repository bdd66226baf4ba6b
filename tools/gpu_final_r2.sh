#!/bin/bash
# Final verification bundle of round 2 (one GPU): the driver's own commands.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -v "^When maskLen\|^Please set\|^Warning: The align" > gpurun_out/gputest_final.log; tail -3 gpurun_out/gputest_final.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== config 5 (defaults)"; python tools/run_config.py 5 --reps 4 > gpurun_out/cfg5_final.txt 2>&1; tail -2 gpurun_out/cfg5_final.txt | cut -c1-400
echo "== latency"; python tools/latency.py 2>&1 | tail -2
echo "== bench (driver arguments)"; /usr/bin/time -v python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo rc=$?; grep "Elapsed (wall" gpurun_out/bench_final_n1.err
echo "== reference arm (driver arguments)"; /usr/bin/time -v python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo rc=$?; grep "Elapsed (wall" gpurun_out/bench_final_ref.err
