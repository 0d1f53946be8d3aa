"""Large grids (C2 Radon 131073 nodes, C4 LSDA batch 65537 nodes): direct warm Poisson solves against the V-cycle paths - time and parity block."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
import bench
ctx = D.Context(0)
rn = [D.Options(86, 17, 50.0, 0.0001, 0.5, 1)]
c4 = [D.Options(Z, 16, 50.0, 0.0002, 0.5, 1) for Z in list(range(21, 31)) + list(range(57, 72))]
for dp in (0, 1):
    ctx.set_option("direct_poisson", dp)
    for name, opts in (("Rn", rn), ("C4", c4)):
        ctx.solve_batch(opts, keep_steps=False)
        t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); t1 = time.perf_counter()
        print(name, "direct_poisson", dp, "wall ms", round(1e3 * (t1 - t0), 2), "steps", [r.n_steps for r in res][:6], "finished", sum(r.finished for r in res), flush=True)
    pb = bench.parity_block(ctx, D)
    print("parity direct_poisson", dp, {c: (f"{v['max_abs_eig_dev_Ha']:.2e}", f"{v['max_abs_energy_dev_Ha']:.2e}", v["atoms_finished"]) for c, v in pb.items() if isinstance(v, dict)}, flush=True)
