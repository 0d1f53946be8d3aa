"""Development aid: per-level cycle counters of one Poisson solve (CTA 0)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dftatom_b200 as D
L = int(sys.argv[1]) if len(sys.argv) > 1 else 14
delta = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0005
nd = int(sys.argv[3]) if len(sys.argv) > 3 else 92
ctx = D.Context(0)
N = (1 << L) + 1
rp = 25.0 / (np.exp((N - 1) * delta) - 1)
r = rp * (np.exp(np.arange(N) * delta) - 1)
rho = np.stack([Z * 8 / np.pi * np.exp(-4 * r) for Z in range(1, nd + 1)])
Zs = np.arange(1, nd + 1, dtype=np.int32)
ctx.poisson_solve(L, delta, 25.0, Zs, rho)
ctx.set_option("profile", 1)
U, used = ctx.poisson_solve(L, delta, 25.0, Zs, rho)
print("ok", U[0, -1], used[:3])
