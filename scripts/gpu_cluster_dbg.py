import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
ctx.solve_batch([D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in (68, 69, 70)], keep_steps=False)
