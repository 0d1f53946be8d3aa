#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one C3 sweep + full ncu captures of the three heaviest kernels.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_c3.csv python scripts/prof_c3.py > gpurun_out/launches_c3.log 2>&1
for k in search_fused poisson_full match_seg; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k python scripts/prof_c3.py > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out
