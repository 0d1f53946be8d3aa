#!/bin/bash
# Run on the GPU box (under gpurun): GPU tests, bench (both arms), launch list of the bench command, full ncu captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; tail -c 6000 gpurun_out/bench_n.json; tail -5 gpurun_out/bench_n.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_n.json 2> gpurun_out/bench_ref_n.err; cat gpurun_out/bench_ref_n.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_n.csv python bench.py --steps 1 --warmup 0 --no-rn --no-batch --no-micro --no-cpu-baseline > gpurun_out/launches_n.log 2>&1
for k in search_seg match_cta poisson_full density_update potential_energy; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o gpurun_out/prof_n_$k python scripts/prof_c3.py > gpurun_out/prof_n_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream_visit -s 0 -c 1 -f -o gpurun_out/prof_n_stream_visit python scripts/gpu_micro.py 64 0 > gpurun_out/prof_n_stream_visit.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream_visit -s 12 -c 1 -f -o gpurun_out/prof_n_stream_visit_up python scripts/gpu_micro.py 64 0 > gpurun_out/prof_n_stream_visit_up.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_n_c5a.csv python scripts/gpu_micro.py 64 0 > gpurun_out/launches_n_c5a.log 2>&1
ls -la gpurun_out | tail -30
timeout 300 ncu --set full --clock-control none --import-source on -k regex:numerov_lanes_seg -s 1 -c 1 -f -o gpurun_out/prof_n_lanes_seg python scripts/prof_c5b.py 16 > gpurun_out/prof_n_lanes_seg.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:match_win -s 3 -c 1 -f -o gpurun_out/prof_n_match_win python scripts/gpu_rn.py > gpurun_out/prof_n_match_win.log 2>&1
