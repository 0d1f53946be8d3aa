"""Host-side stage timing of dftatom_solve_batch on C3 (DFTATOM_DEBUG_HOST=1 prints to stderr)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["DFTATOM_DEBUG_HOST"] = "1"
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
for i in range(4):
    t0 = time.perf_counter(); ctx.solve_batch(opts, keep_steps=False); t1 = time.perf_counter()
    print("call", i, "wall ms", round(1e3 * (t1 - t0), 2), "dev ms", round(ctx.last_timing()[0], 2), file=sys.stderr, flush=True)
