"""Development aid: increment-form warm Poisson solves (delta_poisson) on the golden configs: trajectories dumped for scripts/analyse_traj.py,
plus |dE/E| tails."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in (18, 60, 68, 70, 90)]
for dp in (0, 1):
    ctx.set_option("delta_poisson", dp)
    res = ctx.solve_batch(opts)
    for r in res:
        et = [s.Etotal for s in r.steps]
        tail = [abs((et[k] - et[k - 1]) / et[k]) for k in range(max(1, len(et) - 12), len(et))]
        print(f"delta={dp} Z={r.options.Z} steps {r.n_steps} fin {r.finished} |dE/E| tail:", " ".join(f"{x:.1e}" for x in tail), flush=True)
