import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("step_cap", 6)
for kv in sys.argv[1:]:
    k, v = kv.split("="); ctx.set_option(k, float(v))
r = ctx.solve_batch([D.Options(6, 12, 20.0, 0.001, 0.5, 0)], keep_steps=False)
print("scf L12", sys.argv[1:], [x.n_steps for x in r], r[0].Etotal, flush=True)
