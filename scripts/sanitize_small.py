"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck): SCF on a small grid, stream-mode Poisson at 15 levels,
windowed match at 16 levels (one orbital), parallel-in-r lanes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dftatom_b200 as D
ctx = D.Context(0)
r = ctx.solve_batch([D.Options(10, 10, 15.0, 0.004, 0.5, 0), D.Options(3, 10, 15.0, 0.004, 0.5, 1)])
print("scf", [x.n_steps for x in r], r[0].Etotal)
# stream-mode Poisson solve, 15 levels, 4 densities
L, delta, rmax = 15, 0.0004, 50.0
N = (1 << L) + 1
rp = rmax / (np.exp((N - 1) * delta) - 1); rr = rp * (np.exp(np.arange(N) * delta) - 1)
rho = np.stack([Z * 8 / np.pi * np.exp(-4 * rr) for Z in (1, 2, 3, 4)])
U, used = ctx.poisson_solve(L, delta, rmax, [1, 2, 3, 4], rho)
print("poisson stream", U[:, -1], used)
# matched solution through windows, 16 levels
L, delta = 16, 0.0002
N = (1 << L) + 1
rp = rmax / (np.exp((N - 1) * delta) - 1); rr = rp * (np.exp(np.arange(N) * delta) - 1)
V = np.zeros(N); V[1:] = -30.0 / rr[1:]
u, mp = ctx.numerov_orbital(V, L, delta, rmax, 1, -30.0 ** 2 / 8.0)
print("orbital", mp, float(np.max(np.abs(u))))
sign, lg, cnt = ctx.numerov_lanes(V, L, delta, rmax, np.zeros(64, np.int32), np.linspace(-500, -10, 64), np.zeros(64, np.int32), impl=2)
print("lanes", cnt[:8])
# one SCF on 15 levels with the stream-mode solver (4 atoms)
r = ctx.solve_batch([D.Options(Z, 15, 25.0, 0.00025, 0.5, 0) for Z in (2, 3, 4, 5)], keep_steps=False)
print("scf L15", [x.n_steps for x in r])
