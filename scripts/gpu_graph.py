"""Development aid: SCF loop as a CUDA-graph while node vs the host-driven loop: identical records, device time."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
cases = [("C3", [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]),
         ("tail", [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in (68, 69, 70)]),
         ("L10 mixed", [D.Options(Z, 10, 15.0, 0.004, 0.5, Z % 2) for Z in range(1, 20)]),
         ("C4 part", [D.Options(Z, 16, 50.0, 0.0002, 0.5, 1) for Z in (21, 22, 23, 24, 57)]),
         ("Rn", [D.Options(86, 17, 50.0, 0.0001, 0.5, 1)])]
for name, opts in cases:
    out = {}
    for ug in (0, 1):
        ctx.set_option("use_graph", ug)
        ctx.solve_batch(opts, keep_steps=False)
        t0 = time.perf_counter(); res = ctx.solve_batch(opts); wall = time.perf_counter() - t0
        out[ug] = (res, ctx.last_timing(), wall, ctx.last_graph_iterations())
    same = all(a.n_steps == b.n_steps and [s.Etotal for s in a.steps] == [s.Etotal for s in b.steps] and [s.E for s in a.steps] == [s.E for s in b.steps]
               for a, b in zip(out[0][0], out[1][0]))
    print(f"{name}: host loop {out[0][1][0]:.2f} ms dev / {out[0][2]*1e3:.2f} ms wall ({out[0][1][1]} launches) | graph {out[1][1][0]:.2f} ms dev / {out[1][2]*1e3:.2f} ms wall "
          f"({out[1][1][1]} launches, {out[1][3]} iterations in the while node) | identical records: {same}", flush=True)
