"""Development aid: run golden configs on the GPU and dump every step record (full precision) to gpurun_out/traj_<name><tag>.json
for offline comparison with tests/golden/*.json (scripts/analyse_traj.py).  usage: gpu_dump_traj.py name[,name...] [tag] [key=value ...]"""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D

names = sys.argv[1].split(",")
tag = sys.argv[2] if len(sys.argv) > 2 and "=" not in sys.argv[2] else ""
ctx = D.Context(0)
for kv in sys.argv[2:]:
    if "=" in kv:
        k, v = kv.split("=")
        ctx.set_option(k, float(v))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for name in names:
    g = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))["atoms"]
    groups = {}
    for i, a in enumerate(g):
        o = a["options"]
        groups.setdefault((o["levels"], o["delta"], o["rmax"]), []).append(i)
    out = [None] * len(g)
    for idx in groups.values():
        opts = [D.Options(g[i]["options"]["Z"], g[i]["options"]["levels"], g[i]["options"]["rmax"], g[i]["options"]["delta"], g[i]["options"]["mixing"], g[i]["options"]["method"]) for i in idx]
        t0 = time.time()
        res = ctx.solve_batch(opts)
        print(name, "atoms", len(opts), "wall", round(time.time() - t0, 3), "dev ms / launches", ctx.last_timing(), flush=True)
        for i, r in zip(idx, res):
            out[i] = dict(Z=r.options.Z, method=r.options.method, status=r.status, n_steps=r.n_steps,
                          steps=[dict(E=s.E, Etotal=s.Etotal, Ekin=s.Ekin, Ecoul=s.Ecoul, Eenuc=s.Eenuc, Exc=s.Exc, ok=s.levels_converged) for s in r.steps],
                          sorted=[[(L.n, L.l, L.occ) for L in ch] for ch in r.sorted_levels])
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"traj_{name}{tag}.json"), "w"))
