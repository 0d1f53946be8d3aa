"""GPU experiment: Rn (C2) and the LSDA batch (C4) with the stream-mode Poisson solver on / off."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
ctx = D.Context(0)
ctx.set_option("profile", 1)
rn = [D.Options(86, 17, 50.0, 0.0001, 0.5, 1)]
c4 = [D.Options(Z, 16, 50.0, 0.0002, 0.5, 1) for Z in list(range(21, 31)) + list(range(57, 72))]
for name, opts in (("Rn", rn), ("C4", c4)):
    for stream in (1, 0):
        ctx.set_option("stream_poisson", stream)
        ctx.solve_batch(opts, keep_steps=False)
        res = ctx.solve_batch(opts, keep_steps=False)
        ms, nl = ctx.last_timing()
        prof = {k: round(v["ms"], 1) for k, v in ctx.last_profile().items()}
        print(name, "stream", stream, "dev ms", round(ms, 1), "launches", nl, "steps", [r.n_steps for r in res][:6], "fin", sum(r.finished for r in res), prof,
              "Etotal", res[0].Etotal, flush=True)
