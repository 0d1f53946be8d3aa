#!/bin/bash
# first GPU run of the rows search kernel: component parity, SCF tests, C3 timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_components.py -x -q -k "lanes_match or level_search or hydrogenic" 2>&1 | tail -15
echo "== tail3"; python scripts/gpu_tail3.py 2>&1 | tail -3
echo "== tail3 old kernel"; python scripts/gpu_tail3.py search_kernel=1 2>&1 | tail -3
echo "== C3 per-step"; DFTATOM_DEBUG_STEPS=1 DFTATOM_DEBUG_ROUNDS=1 python scripts/gpu_steps_c3.py 2> gpurun_out/steps_c3_rows.txt; tail -4 gpurun_out/steps_c3_rows.txt; head -12 gpurun_out/steps_c3_rows.txt
