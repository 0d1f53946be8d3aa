"""Development aid: run a golden config on the GPU and print per-atom deviations from the reference."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dftatom_b200 as D

name = sys.argv[1] if len(sys.argv) > 1 else "sweep"
g = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))["atoms"]
ctx = D.Context(0)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ctx.set_option(k, float(v))
opts = [D.Options(a["options"]["Z"], a["options"]["levels"], a["options"]["rmax"], a["options"]["delta"], a["options"]["mixing"], a["options"]["method"]) for a in g]
t0 = time.time()
res = ctx.solve_batch(opts)
print("wall", time.time() - t0, "dev ms", ctx.last_timing())
pr = ctx.last_profile()
print("profile", {k: round(v["ms"], 1) for k, v in pr.items()}, "orbital solves", pr["match"]["work"], "rounds/solve", pr["density"]["work"] / max(1, pr["match"]["work"]), "fallback rounds", pr["potential"]["work"])
worst_e = worst_t = 0
for r, a in zip(res, g):
    nref = a.get("n_steps", len(a["steps"]))
    traj = a.get("etotal_per_step") or [s["Etotal"] for s in a["steps"]]
    n = min(r.n_steps, len(traj))
    dt = max(abs(r.steps[k].Etotal - traj[k]) for k in range(n))
    last = a["steps"][-1]
    k = min(r.n_steps, nref) - 1
    de = max(abs(x - l["E"]) for x, l in zip([x for ch in r.steps[k].E for x in ch], last["levels"])) if r.n_steps >= nref or not a["finished"] else \
        max(abs(x - l["E"]) for x, l in zip([x for ch in r.steps[-1].E for x in ch], last["levels"]))
    nbad = sum(not s.levels_converged for s in r.steps)
    flag = "" if (r.finished == a["finished"] and dt < 1e-5 and de < 1e-6) else "  <<<<"
    worst_e = max(worst_e, de); worst_t = max(worst_t, dt)
    print(f"Z={a['options']['Z']:3d} steps {r.n_steps:3d}/{nref:3d} fin {int(r.finished)}/{int(a['finished'])} status {r.status} max|dEtot| {dt:.2e} max|deig| {de:.2e} unconverged-level-steps {nbad}{flag}")
print("worst eig", worst_e, "worst etot traj", worst_t)
