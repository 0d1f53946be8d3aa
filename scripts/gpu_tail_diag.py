import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dftatom_b200 as D
ctx = D.Context(0)
for kv in sys.argv[1:]:
    k, v = kv.split("="); ctx.set_option(k, float(v))
Zs = [18, 29, 30, 68, 90]
res = ctx.solve_batch([D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in Zs])
for Z, r in zip(Zs, res):
    t = np.array([s.Etotal for s in r.steps]); d = np.abs(np.diff(t)) / abs(t[-1])
    print(Z, r.n_steps, ' '.join(f"{x:.1e}" for x in d[-14:]))
