#!/bin/bash
for cfg in 1042 273 529 274 530; do
  echo "== rows_cfg=$cfg"; python scripts/gpu_tail3.py rows_cfg=$cfg 2>&1 | tail -1
  DFTATOM_DEBUG_ROUNDS=1 python scripts/gpu_steps_c3.py rows_cfg=$cfg 2>&1 | grep -A1 "device ms\|histogram" | grep -v "^--" | cut -c1-200
done
