"""GPU experiment: C3 (16385 nodes) with the Poisson solve in stream mode (levels above 2^mid nodes by slab windows) or one CTA per density."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("profile", 1)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
c4 = [D.Options(Z, 16, 50.0, 0.0002, 0.5, 1) for Z in list(range(21, 31)) + list(range(57, 72))]
ctx.solve_batch(opts, keep_steps=False)
for name, o, cfgs in (("C3", opts, ((15, 11), (14, 11), (14, 12))), ("C4", c4, ((14, 14), (14, 11), (14, 12)))):
    for minlev, mid in cfgs:
        ctx.set_option("stream_min_levels", minlev); ctx.set_option("stream_mid_levels", mid)
        best = 1e9
        for _ in range(2):
            res = ctx.solve_batch(o, keep_steps=False)
            best = min(best, ctx.last_timing()[0])
        print(name, "stream_min_levels", minlev, "mid", mid, "dev ms", round(best, 1), {k: round(v["ms"], 1) for k, v in ctx.last_profile().items()}, "fin", sum(r.finished for r in res),
              "launches", ctx.last_timing()[1], flush=True)
