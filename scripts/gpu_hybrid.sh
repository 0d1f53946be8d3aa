#!/bin/bash
mkdir -p gpurun_out
for T in 24 28 32 36 44; do
  echo "== warm_until_step=$T"; python scripts/gpu_steps_c3.py warm_until_step=$T 2>&1 | grep "device ms"
done
DFTATOM_DEBUG_STEPS=1 python scripts/gpu_steps_c3.py 2> gpurun_out/steps_c3_hybrid.txt; awk 'NR<=8 || NR%6==0' gpurun_out/steps_c3_hybrid.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
