"""GPU check of the direct warm Poisson solve (poisson_direct.cu) against the V-cycle kernels: SCF trajectories, timings, parity block."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
import bench
ctx = D.Context(0)
for L, delta, rmax, Zs in ((14, 0.0005, 25.0, (2, 18, 36, 70)), (12, 0.002, 20.0, (6, 30)), (11, 0.004, 15.0, (4, 10)), (13, 0.001, 25.0, (26,))):
    opts = [D.Options(Z, L, rmax, delta, 0.5, 0) for Z in Zs]
    out = {}
    for dp in (0, 1):
        ctx.set_option("direct_poisson", dp)
        out[dp] = ctx.solve_batch(opts, keep_steps=True)
    for r0, r1 in zip(out[0], out[1]):
        n = min(r0.n_steps, r1.n_steps)
        dE = max(abs(r0.steps[i].Etotal - r1.steps[i].Etotal) for i in range(n))
        dC = max(abs(r0.steps[i].Ecoul - r1.steps[i].Ecoul) for i in range(n))
        print(f"L={L} Z={r0.options.Z}: steps {r0.n_steps}/{r1.n_steps} max|dEtotal| {dE:.3e} max|dEcoul| {dC:.3e} E {r1.Etotal:.9f}", flush=True)
ctx.set_option("profile", 1); ctx.set_option("stream_groups", 1)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in (68, 69, 70)]
for dp in (0, 1):
    ctx.set_option("direct_poisson", dp)
    ctx.solve_batch(opts, keep_steps=False)
    res = ctx.solve_batch(opts, keep_steps=False)
    pr = ctx.last_profile(); n = max(r.n_steps for r in res)
    print("tail3 direct_poisson", dp, "dev ms", round(ctx.last_timing()[0], 2), {k: round(1e3 * v["ms"] / n, 1) for k, v in pr.items()})
ctx.set_option("profile", 0); ctx.set_option("stream_groups", 3)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
for dp in (0, 1):
    ctx.set_option("direct_poisson", dp)
    ctx.solve_batch(opts, keep_steps=False)
    t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); t1 = time.perf_counter()
    pb = bench.parity_block(ctx, D)
    print("C3 direct_poisson", dp, "wall ms", round(1e3 * (t1 - t0), 2), "finished", sum(r.finished for r in res), "steps", sum(r.n_steps for r in res),
          {c: (f"{v['max_abs_eig_dev_Ha']:.2e}", f"{v['max_abs_energy_dev_Ha']:.2e}", v["atoms_finished"]) for c, v in pb.items() if isinstance(v, dict)}, flush=True)
