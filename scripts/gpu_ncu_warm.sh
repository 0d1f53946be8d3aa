#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:poisson_warm -s 20 -c 1 -f -o gpurun_out/prof_r2_poisson_warm python scripts/gpu_tail3.py warm_until_step=0 > gpurun_out/prof_r2_poisson_warm.log 2>&1
tail -3 gpurun_out/prof_r2_poisson_warm.log
python scripts/gpu_warm1.py 2>&1 | grep -v "^L="
