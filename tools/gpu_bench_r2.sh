#!/bin/bash
# The driver's bench commands on one GPU (both arms), plus the long-read one-pair call.
mkdir -p gpurun_out
L=complete-striped-smith-waterman-library_b200/libssw.so
echo "== long-read call"; ./tools/call_bench $L 20 1 100000 10000 2 2>&1 | tail -1; ./tools/call_bench $L 20 1 100000 10000 2 tb_spec=0 2>&1 | tail -1
echo "== bench (driver arguments)"; time (python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err); echo rc=$?
echo "== reference arm (driver arguments)"; time (python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err); echo rc=$?
python tools/summarize.py gpurun_out/bench_final_n1.json gpurun_out/bench_final_ref.json 2>&1 | tail -5
