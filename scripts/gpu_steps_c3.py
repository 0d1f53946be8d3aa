"""Per-step, per-kernel-class time of one C3 sweep (host-driven loop, profile = 1, DFTATOM_DEBUG_STEPS=1 prints to stderr)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
for kv in sys.argv[1:]:
    k, v = kv.split("="); ctx.set_option(k, float(v))
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
ctx.solve_batch(opts, keep_steps=False)
ctx.set_option("profile", 1)
res = ctx.solve_batch(opts, keep_steps=False)
print("device ms", ctx.last_timing()[0], file=sys.stderr)
