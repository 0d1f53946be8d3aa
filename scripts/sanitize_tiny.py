"""Tiny SCF for compute-sanitizer racecheck (all five SCF kernels on a 1025-node grid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
r = ctx.solve_batch([D.Options(4, 10, 15.0, 0.004, 0.5, 0)])
print("scf", [x.n_steps for x in r], r[0].Etotal)
