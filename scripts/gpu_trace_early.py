"""Round-by-round traces of long level searches in the first SCF steps (debug build: make EXTRA=-DDFT_ROWS_DEBUG)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("profile", 1); ctx.set_option("stream_groups", 1); ctx.set_option("step_cap", int(sys.argv[1]) if len(sys.argv) > 1 else 3)
ctx.solve_batch([D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)], keep_steps=False)
