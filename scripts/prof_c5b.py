"""Profiling driver (run under ncu): the C5b lanes micro-benchmark, parallel-in-r kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("r_segments", int(sys.argv[1]) if len(sys.argv) > 1 else 32)
print(bench.micro_c5b(ctx, 35.0, reps=2, cpu_baseline=False)["kernels"])
