"""Worst deviation from the reference goldens (bench.parity_block) and C3 time for option variants given as k=v,k=v groups on the command line."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
import bench
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
for grp in sys.argv[1:]:
    kv = dict(x.split("=") for x in grp.split(",") if x)
    for k, v in kv.items(): ctx.set_option(k, float(v))
    ctx.solve_batch(opts, keep_steps=False)
    t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); t1 = time.perf_counter()
    pb = bench.parity_block(ctx, D)
    print(grp or "defaults", "| C3 wall ms", round(1e3 * (t1 - t0), 2), "finished", sum(r.finished for r in res), "|",
          {c: (f"{v['max_abs_eig_dev_Ha']:.2e}", f"{v['max_abs_energy_dev_Ha']:.2e}", v["atoms_finished"]) for c, v in pb.items() if isinstance(v, dict)}, flush=True)
