"""Profiling driver (run under ncu): one C3 sweep, nothing else."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
zs = range(1, 93) if len(sys.argv) < 2 else [int(z) for z in sys.argv[1].split(",")]
ctx = D.Context(0)
for kv in os.environ.get("DFT_OPTS", "").split():
    k_, v_ = kv.split("=")
    ctx.set_option(k_, float(v_))
res = ctx.solve_batch([D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in zs], keep_steps=False)
print("steps", [r.n_steps for r in res][:8], "dev ms", ctx.last_timing())
