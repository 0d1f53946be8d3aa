import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
for until, win in ((0, 8192), (32, 8192), (32, 5632), (32, 12288), (0, 8192)):
    ctx.set_option("match_win_until_step", until); ctx.set_option("match_win_nodes", win)
    ctx.solve_batch(opts, keep_steps=False)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); ts.append(time.perf_counter() - t0)
    print("match windows until step", until, "nodes", win, "wall ms", [round(1e3 * t, 2) for t in ts], "finished", sum(r.finished for r in res), flush=True)
ctx.set_option("match_win_until_step", 0)
rn = [D.Options(86, 17, 50.0, 0.0001, 0.5, 1)]
c4 = [D.Options(Z, 16, 50.0, 0.0002, 0.5, 1) for Z in list(range(21, 31)) + list(range(57, 72))]
for name, o in (("Rn", rn), ("C4", c4)):
    ctx.solve_batch(o, keep_steps=False)
    t0 = time.perf_counter(); res = ctx.solve_batch(o, keep_steps=False); t1 = time.perf_counter()
    print(name, "wall ms", round(1e3 * (t1 - t0), 2), "finished", sum(r.finished for r in res), flush=True)
