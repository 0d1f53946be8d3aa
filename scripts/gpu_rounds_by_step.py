"""Search rounds per orbital solve by SCF step of the C3 sweep (cumulative counters of runs capped at k steps, differenced)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
for kv in sys.argv[1:]:
    k, v = kv.split("="); ctx.set_option(k, float(v))
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
ctx.set_option("profile", 1)
ctx.set_option("stream_groups", 1)
prev = (0.0, 0.0, 0.0)
for cap in list(range(1, 13)) + [16, 24, 32, 48, 100]:
    ctx.set_option("step_cap", cap)
    ctx.solve_batch(opts, keep_steps=False)
    pr = ctx.last_profile()
    cur = (pr["density"]["work"], pr["match"]["work"], pr["search"]["ms"])
    d = [c - p for c, p in zip(cur, prev)]
    print(f"steps < {cap:3d}: rounds {d[0]:8.0f} solves {d[1]:7.0f} rounds/solve {d[0] / max(d[1], 1):5.2f} search ms {d[2]:6.2f}")
    prev = cur
