#!/bin/bash
# the C3 bench line without tests and side legs (A/B of a kernel change)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-rn --no-batch --no-micro --no-cpu-baseline --no-parity > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_quick.json") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 4),
      "kernels ms/sweep", {k: round(v["ms"] / d["steps"], 2) for k, v in d["kernels"].items()}, "rounds", round(d["search"]["rounds_per_solve"], 3))
PY
