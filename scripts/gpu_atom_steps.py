"""Development aid: per-step deviation of one golden atom (all-steps fixtures: argon, radon, small)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dftatom_b200 as D
name = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
a = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))["atoms"][idx]
ctx = D.Context(0)
for kv in sys.argv[3:]:
    k, v = kv.split("="); ctx.set_option(k, float(v))
o = a["options"]
r = ctx.solve_batch([D.Options(o["Z"], o["levels"], o["rmax"], o["delta"], o["mixing"], o["method"])])[0]
print("steps", r.n_steps, len(a["steps"]))
for k in range(min(r.n_steps, len(a["steps"]))):
    g = a["steps"][k]; s = r.steps[k]
    e = np.array([x for ch in s.E for x in ch]); eg = np.array([l["E"] for l in g["levels"]])
    j = int(np.argmax(np.abs(e - eg)))
    print(f"{k:3d} dEtot {s.Etotal-g['Etotal']:+.2e} dEkin {s.Ekin-g['Ekin']:+.2e} dEcoul {s.Ecoul-g['Ecoul']:+.2e} dEenuc {s.Eenuc-g['Eenuc']:+.2e} dExc {s.Exc-g['Exc']:+.2e} "
          f"max|deig| {abs(e[j]-eg[j]):.2e} (level {j}: {g['levels'][j]['n']}{'spdf'[g['levels'][j]['l']]})")
