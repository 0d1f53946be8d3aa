"""Development aid: cluster-mode Poisson (poisson_cluster.cu) vs one CTA per density: SCF trajectories and kernel-class timings."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("profile", 1)
def run(opts, cluster):
    ctx.set_option("cluster_poisson", cluster)
    ctx.solve_batch(opts, keep_steps=False)
    res = ctx.solve_batch(opts)
    ms, nl = ctx.last_timing()
    pr = ctx.last_profile()
    return res, ms, {k: round(v["ms"], 2) for k, v in pr.items()}
cases = [("L14 4 atoms", [D.Options(Z, 14, 25.0, 0.0005, 0.5, m) for Z, m in [(4, 0), (18, 0), (26, 1), (70, 0)]]),
         ("L13", [D.Options(Z, 13, 25.0, 0.001, 0.5, 0) for Z in (10, 36)]),
         ("L12", [D.Options(Z, 12, 20.0, 0.001, 0.5, 0) for Z in (10, 36)]),
         ("L11", [D.Options(Z, 11, 15.0, 0.002, 0.5, 0) for Z in (10, 36)]),
         ("C3 sweep", [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]),
         ("tail Z=68-70", [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in (68, 69, 70)])]
for name, opts in cases:
    r0, ms0, p0 = run(opts, 0)
    r1, ms1, p1 = run(opts, 1)
    dE = de = 0.
    for a, b in zip(r0, r1):
        n = min(a.n_steps, b.n_steps)
        for k in range(n):
            dE = max(dE, max(abs(getattr(a.steps[k], key) - getattr(b.steps[k], key)) for key in ("Etotal", "Ekin", "Ecoul", "Eenuc", "Exc")))
            de = max(de, max(abs(x - y) for ca, cb in zip(a.steps[k].E, b.steps[k].E) for x, y in zip(ca, cb)))
    print(f"{name}: one-CTA {ms0:.2f} ms {p0} | cluster {ms1:.2f} ms {p1} | steps {[r.n_steps for r in r0][:6]} vs {[r.n_steps for r in r1][:6]} | max dE {dE:.2e} deig {de:.2e}", flush=True)
