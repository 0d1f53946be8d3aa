"""GPU experiment: C3 sweep device time against seg_threshold / r_segments."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("profile", 1)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
ctx.solve_batch(opts, keep_steps=False)
for segs in (-1,):
    for thr in (2400,):
        ctx.set_option("r_segments", segs); ctx.set_option("seg_threshold", thr)
        best = 1e9
        for _ in range(2):
            res = ctx.solve_batch(opts, keep_steps=False)
            best = min(best, ctx.last_timing()[0])
        pr = ctx.last_profile()
        print("segs", segs, "thr", thr, "dev ms", round(best, 1), {k: round(v["ms"], 1) for k, v in pr.items()}, "fin", sum(r.finished for r in res), flush=True)
