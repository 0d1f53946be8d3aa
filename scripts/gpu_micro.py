"""GPU experiment: the C5a / C5b kernel micro-benchmarks of bench.py for every stream-visit window shape."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import dftatom_b200 as D

nd = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2]
ctx = D.Context(0)
peak = ctx.measure_fp64_peak()
print("fp64 peak", peak, flush=True)
print(json.dumps(bench.micro_c5b(ctx, peak)), flush=True)
for v in variants:
    ctx.set_option("stream_variant", v % 10)
    ctx.set_option("stream_mid_levels", 14 if v >= 10 else 11)
    r = bench.micro_c5a(ctx, torch, bench._hbm_peak(), n_dens=nd, cpu_baseline=(v == variants[0]))
    print("variant", v, json.dumps(r), flush=True)
