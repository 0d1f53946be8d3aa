"""GPU experiment: C4 / Rn device time against seg_threshold / r_segments."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("profile", 1)
rn = [D.Options(86, 17, 50.0, 0.0001, 0.5, 1)]
c4 = [D.Options(Z, 16, 50.0, 0.0002, 0.5, 1) for Z in list(range(21, 31)) + list(range(57, 72))]
for name, opts in (("C4", c4), ("Rn", rn)):
    for segs, thr in ((32, 300), (32, 100000), (16, 100000), (8, 100000)):
        ctx.set_option("r_segments", segs); ctx.set_option("seg_threshold", thr)
        res = ctx.solve_batch(opts, keep_steps=False)
        pr = ctx.last_profile()
        nsteps = max(r.n_steps for r in res)
        print(name, "segs", segs, "thr", thr, "dev ms", round(ctx.last_timing()[0], 1), "max steps", nsteps, "search ms/step", round(pr["search"]["ms"] / nsteps, 3),
              {k: round(v["ms"], 1) for k, v in pr.items()}, flush=True)
