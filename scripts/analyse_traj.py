"""Development aid (CPU): compare gpurun_out/traj_<name><tag>.json (scripts/gpu_dump_traj.py) with tests/golden/<name>.json.
Prints, per atom: stop steps, max per-step deviations at equal step index, and FINAL-vs-FINAL deviations regardless of stop step."""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name = sys.argv[1]
tag = sys.argv[2] if len(sys.argv) > 2 else ""
verbose = len(sys.argv) > 3
g = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))["atoms"]
t = json.load(open(os.path.join(ROOT, "gpurun_out", f"traj_{name}{tag}.json")))
KEYS = ("Etotal", "Ekin", "Ecoul", "Eenuc", "Exc")
worst = dict(step_eig=0., step_E=0., fin_eig=0., fin_E=0.)
nbad = 0
for a, r in zip(g, t):
    nref = a.get("n_steps", len(a["steps"]))
    traj = a.get("etotal_per_step") or [s["Etotal"] for s in a["steps"]]
    eigs = a.get("eig_per_step") or [[l["E"] for l in s["levels"]] for s in a["steps"]]
    en = a.get("energies_per_step")
    if en is None and len(a["steps"]) == nref:
        en = [[s[k] for k in KEYS] for s in a["steps"]]
    n = min(r["n_steps"], nref)
    de = max(np.abs(np.array([x for ch in r["steps"][k]["E"] for x in ch]) - np.array(eigs[k])).max() for k in range(n))
    dt = max(abs(r["steps"][k]["Etotal"] - traj[k]) for k in range(n))
    d5 = max(max(abs(r["steps"][k][key] - en[k][j]) for j, key in enumerate(KEYS)) for k in range(n)) if en else float("nan")
    last_ref = a["steps"][-1]
    fe = np.abs(np.array([x for ch in r["steps"][-1]["E"] for x in ch]) - np.array([l["E"] for l in last_ref["levels"]])).max()
    fE = max(abs(r["steps"][-1][k] - last_ref[k]) for k in KEYS)
    conf_ref = [[tuple(x) for x in a["final"]["alpha"]]] + ([[tuple(x) for x in a["final"]["beta"]]] if a["final"].get("beta") else [])
    conf = [[tuple(x) for x in ch] for ch in r["sorted"]]
    fin = r["status"] == 0
    comparable = a["finished"] or r["n_steps"] == nref
    flag = ""
    if comparable and (fe > 1e-6 or fE > 1e-5 or conf != conf_ref): flag += " FINAL"
    if de > 1e-6 or dt > 1e-5 or (en and d5 > 1e-5): flag += " STEP"
    if fin != a["finished"]: flag += " FIN"
    if flag: nbad += 1
    worst["step_eig"] = max(worst["step_eig"], de); worst["step_E"] = max(worst["step_E"], d5 if en else dt)
    if comparable:
        worst["fin_eig"] = max(worst["fin_eig"], fe); worst["fin_E"] = max(worst["fin_E"], fE)
    if verbose or flag:
        print(f"Z={a['options']['Z']:3d} m={a['options']['method']} steps {r['n_steps']:3d}/{nref:3d} fin {int(fin)}/{int(a['finished'])} step: eig {de:.1e} Etot {dt:.1e} 5E {d5:.1e} | final: eig {fe:.1e} E {fE:.1e} conf {'ok' if conf == conf_ref else 'DIFF'}{flag}")
print(name, tag, "atoms", len(g), "flagged", nbad, {k: f"{v:.2e}" for k, v in worst.items()}, "finished", sum(r["status"] == 0 for r in t), "/ ref", sum(a["finished"] for a in g))
