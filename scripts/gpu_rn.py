"""Development aid: Radon (C2) SCF with option overrides: steps to stop and timing."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("profile", 1)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    ctx.set_option(k, float(v))
o = [D.Options(86, 17, 50.0, 0.0001, 0.5, 1)]
ctx.solve_batch(o, keep_steps=False)
r = ctx.solve_batch(o)[0]
print(sys.argv[1:], "steps", r.n_steps, "finished", r.finished, "dev ms", round(ctx.last_timing()[0], 1), {k: round(v["ms"], 1) for k, v in ctx.last_profile().items()})
et = [s.Etotal for s in r.steps]
print("  |dE/E| tail:", ["%.1e" % abs((et[k] - et[k - 1]) / et[k]) for k in range(max(1, len(et) - 12), len(et))])
