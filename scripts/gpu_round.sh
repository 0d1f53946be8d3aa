#!/bin/bash
# Run on the GPU box (under gpurun): GPU tests, bench (both arms), launch list of one C3 sweep, full ncu captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c3.csv python scripts/prof_c3.py > gpurun_out/launches_c3.log 2>&1
for k in search_fused poisson_full match_seg density_update potential_energy; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -f -o gpurun_out/prof_$k python scripts/prof_c3.py > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out
