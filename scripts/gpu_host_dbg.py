import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
for i in range(3):
    t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); t1 = time.perf_counter()
    print(f"python wall {1e3*(t1-t0):.2f} ms, device {ctx.last_timing()[0]:.2f} ms", flush=True)
