"""GPU experiment: the tail regime of C3 - only the three atoms that run all 100 steps (Z = 68, 69, 70)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("profile", 1)
for kv in sys.argv[1:]:
    k, v = kv.split("="); ctx.set_option(k, float(v))
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in (68, 69, 70)]
ctx.solve_batch(opts, keep_steps=False)
res = ctx.solve_batch(opts, keep_steps=False)
pr = ctx.last_profile()
n = max(r.n_steps for r in res)
print("steps", n, "dev ms", round(ctx.last_timing()[0], 2), "us per step:", {k: round(1e3 * v["ms"] / n, 1) for k, v in pr.items()},
      "rounds/solve", pr["density"]["work"] / max(1, pr["match"]["work"]))
