import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
for cfg in (0x111, 0x112, 0x212, 0x211, 0x122, 0x412):
    ctx.set_option("rows_cfg", cfg)
    ctx.solve_batch(opts, keep_steps=False)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); ts.append(time.perf_counter() - t0)
    print("rows_cfg", hex(cfg), "wall ms", [round(1e3 * t, 2) for t in ts], flush=True)
