#!/bin/bash
mkdir -p gpurun_out
python scripts/gpu_rounds_by_step.py 2>&1 | tail -17
python scripts/gpu_rounds_by_step.py search_predict=0 2>&1 | head -4
bash scripts/gpu_quick.sh
