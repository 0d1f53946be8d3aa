import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in (68, 69, 70)]
for ug in (1, 0, 1, 0):
    ctx.set_option("use_graph", ug)
    ctx.solve_batch(opts, keep_steps=False)
    t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); t1 = time.perf_counter()
    print("tail3 use_graph", ug, "wall ms", round(1e3 * (t1 - t0), 2), "dev ms", round(ctx.last_timing()[0], 2), "launches", ctx.last_timing()[1], flush=True)
