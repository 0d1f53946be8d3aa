"""compute-sanitizer probe of the CUDA-graph SCF loop at L = 12: argv[1] = order of the two solves ("hg": host loop then graph, "g": graph only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("step_cap", 6)
for mode in sys.argv[1]:
    ctx.set_option("use_graph", 1 if mode == "g" else 0)
    r = ctx.solve_batch([D.Options(6, 12, 20.0, 0.001, 0.5, 0)], keep_steps=False)
    print("scf L12", mode, [x.n_steps for x in r], r[0].Etotal, "graph iterations", ctx.last_graph_iterations(), flush=True)
