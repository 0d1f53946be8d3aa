import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
for sg in (1, 2, 3, 4, 6):
    ctx.set_option("stream_groups", sg)
    ctx.solve_batch(opts, keep_steps=False)
    t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); t1 = time.perf_counter()
    print("stream_groups", sg, "wall ms", round(1e3 * (t1 - t0), 2), "dev ms", round(ctx.last_timing()[0], 2), "finished", sum(r.finished for r in res), flush=True)
big = opts * 8
for sg in (1, 2, 4, 8):
    ctx.set_option("stream_groups", sg)
    ctx.solve_batch(big, keep_steps=False)
    t0 = time.perf_counter(); res = ctx.solve_batch(big, keep_steps=False); t1 = time.perf_counter()
    print("8xC3 stream_groups", sg, "wall ms", round(1e3 * (t1 - t0), 2), flush=True)
