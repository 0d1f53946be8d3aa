"""Round-2 kernels under compute-sanitizer (memcheck / racecheck): cluster-mode Poisson (L = 11: one distributed level, and L = 12), the
increment-form solves, the bit-reproducible Poisson mode, the CUDA-graph SCF loop, the uniform-grid pair, the outward node count."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("step_cap", int(os.environ.get("SAN_STEPS", "7")))
r = ctx.solve_batch([D.Options(4, 11, 15.0, 0.002, 0.5, 0), D.Options(3, 11, 15.0, 0.002, 0.5, 1)], keep_steps=False)
print("scf L11 cluster + graph", [x.n_steps for x in r], r[0].Etotal, "graph iterations", ctx.last_graph_iterations())
r = ctx.solve_batch([D.Options(6, 12, 20.0, 0.001, 0.5, 0)], keep_steps=False)
print("scf L12 cluster", [x.n_steps for x in r], r[0].Etotal)
ctx.set_option("use_graph", 0)
r = ctx.solve_batch([D.Options(4, 11, 15.0, 0.002, 0.5, 0)], keep_steps=False)
print("scf L11 host loop", [x.n_steps for x in r], r[0].Etotal)
ctx.set_option("use_graph", 1)
r = ctx.solve_batch([D.Options(2, 10, 15.0, 0.0, 0.5, 2), D.Options(3, 10, 15.0, 0.0, 0.5, 3)], keep_steps=False)
print("scf uniform grid", [x.n_steps for x in r], r[0].Etotal)
ctx.set_option("step_cap", 3)
ctx.set_option("poisson_exact", 1)
r = ctx.solve_batch([D.Options(4, 10, 15.0, 0.004, 0.5, 0)], keep_steps=False)
ctx.set_option("poisson_exact", 0)
print("scf exact poisson", [x.n_steps for x in r], r[0].Etotal)
L, delta, rmax = 10, 0.004, 15.0
N = (1 << L) + 1
rp = rmax / (np.exp((N - 1) * delta) - 1); rr = rp * (np.exp(np.arange(N) * delta) - 1)
V = np.zeros(N); V[1:] = -10.0 / rr[1:]
_, _, cnt = ctx.numerov_lanes(V, L, delta, rmax, np.zeros(8, np.int32), np.linspace(-40, -1, 8), np.full(8, 9, np.int32), impl=3)
print("outward count", cnt)
