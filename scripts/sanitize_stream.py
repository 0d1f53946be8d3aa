"""Stream-mode Poisson (15 levels) and windowed match (16 levels) for compute-sanitizer racecheck."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("max_vcycles", 1)
L, delta, rmax = 15, 0.0004, 50.0
N = (1 << L) + 1
rp = rmax / (np.exp((N - 1) * delta) - 1); rr = rp * (np.exp(np.arange(N) * delta) - 1)
rho = np.stack([Z * 8 / np.pi * np.exp(-4 * rr) for Z in (1, 2, 3, 4)])
U, used = ctx.poisson_solve(L, delta, rmax, [1, 2, 3, 4], rho)
print("poisson stream", U[:, -1], used)
if len(sys.argv) > 1:
    L, delta = 16, 0.0002
    N = (1 << L) + 1
    rp = rmax / (np.exp((N - 1) * delta) - 1); rr = rp * (np.exp(np.arange(N) * delta) - 1)
    V = np.zeros(N); V[1:] = -30.0 / rr[1:]
    u, mp = ctx.numerov_orbital(V, L, delta, rmax, 1, -30.0 ** 2 / 8.0)
    print("orbital", mp, float(np.max(np.abs(u))))
