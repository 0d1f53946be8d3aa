#!/bin/bash
mkdir -p gpurun_out
{
for tool in memcheck racecheck; do
  for g in 0 1; do
    echo "== $tool use_graph=$g"
    SAN_GRAPH=$g timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_r2e.py 2>&1 | grep -v "Host Frame" | tail -14
  done
done
} > gpurun_out/sanitizer_r2.txt 2>&1
tail -60 gpurun_out/sanitizer_r2.txt
