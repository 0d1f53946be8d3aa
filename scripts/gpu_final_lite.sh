#!/bin/bash
# refresh of the bench line and the launch list of the same command (kernels unchanged since the last full pass: scripts/gpu_final_r2.sh)
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 300 gpurun_out/bench_r2d.json; tail -3 gpurun_out/bench_r2d.err
DFTATOM_OPTIONS="use_graph=0 stream_groups=1" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 0 --no-rn --no-batch --no-micro --no-cpu-baseline --no-parity > gpurun_out/launches_r2.log 2>&1
