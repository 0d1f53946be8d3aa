"""Development aid: Poisson exact mode - bit check against the oracle, timing, and SCF trajectories (dumped like gpu_dump_traj)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dftatom_b200 as D
import oracle_lib as O
ctx = D.Context(0)
for (L, delta, rmax) in [(10, 0.004, 15.0), (14, 0.0005, 25.0), (16, 0.0002, 50.0)]:
    N, rp, r = O.grid(L, delta, rmax)
    Zs = [1, 18, 86]
    rho = np.stack([Z * k ** 3 / np.pi * np.exp(-2 * k * r) for Z, k in zip(Zs, [0.8, 1.7, 3.1])])
    ctx.set_option("poisson_exact", 1)
    t0 = time.time(); U, used = ctx.poisson_solve(L, delta, rmax, Zs, rho); t1 = time.time()
    ctx.set_option("poisson_exact", 0)
    for j, Z in enumerate(Zs):
        U_o, errs = O.poisson(L, delta, rmax, Z, rho[j], max_vcycles=100)
        print(f"L={L} Z={Z}: bit-identical {np.array_equal(U[j], U_o)} max|dU| {np.max(np.abs(U[j]-U_o)):.3e} differing nodes {int((U[j]!=U_o).sum())} vcycles {used[j]}/{len(errs)}  wall {t1-t0:.3f}s", flush=True)
