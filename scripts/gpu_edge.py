"""GPU experiment: edge-case options against the oracle (coarse grids, heaviest atom, hydrogen LSDA)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dftatom_b200 as D
import oracle_lib as O
ctx = D.Context(0)
cases = [(1, 10, 15.0, 0.004, 0.5, 1), (118, 12, 25.0, 0.002, 0.5, 0), (118, 12, 25.0, 0.002, 0.5, 1), (2, 8, 10.0, 0.02, 0.5, 0), (10, 6, 10.0, 0.08, 0.5, 0),
         (36, 12, 10.0, 0.001, 0.0, 0), (36, 12, 10.0, 0.001, 0.9, 0), (3, 4, 5.0, 0.3, 0.5, 1), (57, 20, 50.0, 1.25e-5, 0.5, 0)]
for Z, L, rmax, delta, mix, m in cases:
    try:
        r = ctx.solve_batch([D.Options(Z, L, rmax, delta, mix, m)])[0]
    except Exception as e:
        print((Z, L, rmax, delta, mix, m), "GPU error", e); continue
    if L <= 14:
        ref = O.scf(Z, L, mix, rmax, delta, m, max_vcycles=12)
        n = min(r.n_steps, len(ref["steps"]))
        de = max(abs(r.steps[k].Etotal - ref["steps"][k]["Etotal"]) for k in range(n))
        dl = max(max(abs(a - b) for a, b in zip([x for ch in r.steps[k].E for x in ch], ref["steps"][k]["E"][0] + ref["steps"][k]["E"][1])) for k in range(n))
        print((Z, L, rmax, delta, mix, m), "steps", r.n_steps, len(ref["steps"]), "fin", r.finished, ref["finished"], "max|dEtot|", "%.2e" % de, "max|deig|", "%.2e" % dl, "Etot", r.Etotal, flush=True)
    else:
        print((Z, L, rmax, delta, mix, m), "steps", r.n_steps, "fin", r.finished, "status", r.status, "Etot", r.Etotal, "dev ms", ctx.last_timing()[0], flush=True)
