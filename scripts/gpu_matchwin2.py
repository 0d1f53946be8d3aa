import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
for until in (32, 0, 32, 0):
    ctx.set_option("match_win_until_step", until)
    ctx.solve_batch(opts, keep_steps=False)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); ts.append(time.perf_counter() - t0)
    print("match windows until step", until, "wall ms", [round(1e3 * t, 2) for t in ts], "finished", sum(r.finished for r in res), flush=True)
