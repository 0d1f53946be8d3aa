import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
base = None
for wf in (0, 32, 24, 40, 0, 32):
    ctx.set_option("rows_wide_from_step", wf)
    ctx.solve_batch(opts, keep_steps=False)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); ts.append(time.perf_counter() - t0)
    if base is None: base = res
    dE = max(abs(a.Etotal - b.Etotal) for a, b in zip(res, base))
    print("rows_wide_from_step", wf, "wall ms", [round(1e3 * t, 2) for t in ts], "finished", sum(r.finished for r in res), "max|dE final|", f"{dE:.1e}", flush=True)
