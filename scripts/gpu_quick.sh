#!/bin/bash
# quick check of a kernel change: GPU tests, the tail regime (Z = 68-70), the C3 bench line without the side legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
python scripts/gpu_tail3.py use_graph=0 stream_groups=1 2>&1 | tail -1
python scripts/gpu_tail3.py 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-rn --no-batch --no-micro --no-cpu-baseline --no-parity > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_quick.json") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 4),
      "full_load", d["roofline"].get("full_load"), "kernels ms/sweep", {k: round(v["ms"] / d["steps"], 2) for k, v in d["kernels"].items()}, "rounds", round(d["search"]["rounds_per_solve"], 3))
PY
