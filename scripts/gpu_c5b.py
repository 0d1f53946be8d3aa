"""GPU experiment: C5b lanes micro-benchmark for several segment counts."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dftatom_b200 as D
ctx = D.Context(0)
peak = ctx.measure_fp64_peak()
for segs in (32, 16, 8):
    ctx.set_option("r_segments", segs)
    r = bench.micro_c5b(ctx, peak, cpu_baseline=False)
    print("r_segments", segs, json.dumps(r["kernels"]), r["known_answer_ok"], flush=True)
