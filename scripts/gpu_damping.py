"""Development aid: opt-in adaptive damping on the C3 sweep: SCF steps with and without, energies of the newly converged atoms."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
base = ctx.solve_batch(opts, keep_steps=False); t0 = ctx.last_timing()[0]
ctx.set_option("adaptive_mixing", 1)
damp = ctx.solve_batch(opts, keep_steps=False); t1 = ctx.last_timing()[0]
print("steps  base", sum(r.n_steps for r in base), "finished", sum(r.finished for r in base), f"{t0:.1f} ms | damped", sum(r.n_steps for r in damp), "finished", sum(r.finished for r in damp), f"{t1:.1f} ms")
for b, d in zip(base, damp):
    if b.n_steps != d.n_steps:
        print(f"Z={b.options.Z}: {b.n_steps} -> {d.n_steps} steps, fin {b.finished}->{d.finished}, Etotal {b.Etotal:.9f} -> {d.Etotal:.9f}  (diff {d.Etotal - b.Etotal:+.2e})")
