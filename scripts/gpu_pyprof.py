"""Where the host time of Context.solve_batch goes (cProfile over 20 C3 sweeps)."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
ctx.solve_batch(opts, keep_steps=False)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    ctx.solve_batch(opts, keep_steps=False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
