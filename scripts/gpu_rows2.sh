#!/bin/bash
mkdir -p gpurun_out
echo "== tail3"; python scripts/gpu_tail3.py 2>&1 | tail -3
echo "== C3 per-step"; DFTATOM_DEBUG_STEPS=1 DFTATOM_DEBUG_ROUNDS=1 python scripts/gpu_steps_c3.py $@ 2> gpurun_out/steps_c3_rows.txt; grep -A3 "device ms" gpurun_out/steps_c3_rows.txt; awk 'NR<=12 || NR%8==0' gpurun_out/steps_c3_rows.txt | cut -c1-48
