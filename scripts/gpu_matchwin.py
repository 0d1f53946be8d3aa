import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
base = None
for until, win in ((0, 8192), (32, 8192), (32, 5632), (32, 4096), (32, 12288), (100, 8192)):
    ctx.set_option("match_win_until_step", until); ctx.set_option("match_win_nodes", win)
    ctx.solve_batch(opts, keep_steps=False)
    t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); t1 = time.perf_counter()
    if base is None: base = res
    dE = max(abs(a.Etotal - b.Etotal) for a, b in zip(res, base)); dn = sum(a.n_steps != b.n_steps for a, b in zip(res, base))
    print("match windows until step", until, "nodes", win, "wall ms", round(1e3 * (t1 - t0), 2), "finished", sum(r.finished for r in res), "max|dE final|", f"{dE:.2e}", "atoms with another stop step", dn, flush=True)
