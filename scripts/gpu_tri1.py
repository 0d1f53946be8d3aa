"""GPU check of the exact coarse solve (poisson_tri.cuh) in both warm Poisson kernels: SCF trajectories against the swept coarse levels, timings."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
for L, delta, rmax, Zs in ((14, 0.0005, 25.0, (2, 18, 36, 70)), (12, 0.002, 20.0, (6, 30)), (11, 0.004, 15.0, (4, 10)), (13, 0.001, 25.0, (26,))):
    opts = [D.Options(Z, L, rmax, delta, 0.5, 0) for Z in Zs]
    out = {}
    for name, kv in (("swept/cluster", dict(coarse_exact=0, warm_poisson=0)), ("exact/cluster", dict(coarse_exact=1, warm_poisson=0)), ("exact/warm", dict(coarse_exact=1, warm_poisson=1, warm_until_step=0))):
        for k_, v_ in kv.items(): ctx.set_option(k_, v_)
        out[name] = ctx.solve_batch(opts, keep_steps=True)
    ref = out["swept/cluster"]
    for name in ("exact/cluster", "exact/warm"):
        for r0, r1 in zip(ref, out[name]):
            n = min(r0.n_steps, r1.n_steps)
            dE = max(abs(r0.steps[i].Etotal - r1.steps[i].Etotal) for i in range(n))
            dC = max(abs(r0.steps[i].Ecoul - r1.steps[i].Ecoul) for i in range(n))
            print(f"L={L} Z={r0.options.Z} {name}: steps {r0.n_steps}/{r1.n_steps} max|dEtotal| {dE:.3e} max|dEcoul| {dC:.3e} E {r1.Etotal:.9f}", flush=True)
ctx.set_option("profile", 1)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in (68, 69, 70)]
for name, kv in (("swept/cluster", dict(coarse_exact=0, warm_poisson=0)), ("exact/cluster", dict(coarse_exact=1, warm_poisson=0)), ("exact/warm", dict(coarse_exact=1, warm_poisson=1, warm_until_step=0)), ("swept/warm", dict(coarse_exact=0, warm_poisson=1, warm_until_step=0))):
    for k_, v_ in kv.items(): ctx.set_option(k_, v_)
    ctx.solve_batch(opts, keep_steps=False)
    res = ctx.solve_batch(opts, keep_steps=False)
    pr = ctx.last_profile(); n = max(r.n_steps for r in res)
    print("tail3", name, "dev ms", round(ctx.last_timing()[0], 2), {k: round(1e3 * v["ms"] / n, 1) for k, v in pr.items()})
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
ctx.set_option("profile", 0)
for name, kv in (("swept hybrid", dict(coarse_exact=0, warm_poisson=1, warm_until_step=32)), ("exact hybrid", dict(coarse_exact=1, warm_poisson=1, warm_until_step=32)), ("exact warm only", dict(coarse_exact=1, warm_until_step=0))):
    for k_, v_ in kv.items(): ctx.set_option(k_, v_)
    ctx.solve_batch(opts, keep_steps=False)
    res = ctx.solve_batch(opts, keep_steps=False)
    print("C3", name, "dev ms", round(ctx.last_timing()[0], 2), "finished", sum(r.finished for r in res), "steps", sum(r.n_steps for r in res))
