#!/bin/bash
# compute-sanitizer pass over the round-2 kernels (host-driven SCF loop: see DESIGN.md section 4.4 for the graph loop under the tool)
mkdir -p gpurun_out
{
for tool in memcheck racecheck; do
  echo "== $tool (use_graph=0)"
  SAN_GRAPH=0 timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_r2e.py 2>&1 | grep -v "Host Frame" | tail -24
done
} > gpurun_out/sanitizer_r2.txt 2>&1
tail -50 gpurun_out/sanitizer_r2.txt
