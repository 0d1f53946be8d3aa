#!/usr/bin/env python
"""Generate golden fixtures by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/dftatom_ref_hp, built by
oracle/Makefile from /root/reference) and parsing its stdout.  Test infrastructure only.

usage: python scripts/make_goldens.py NAME [--jobs J]
  NAME in: small, argon, radon, sweep, lsda_batch, uniform
Writes tests/golden/<NAME>.json.  Each atom record holds every SCF step the reference printed
(eigenvalues + five energies, 17 significant digits) unless --final-only semantics apply (sweep,
lsda_batch keep compact per-step arrays - the five energies and all eigenvalues of every step - plus the full record of the
last step, to keep fixtures small).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dftatom_b200.report import parse_report  # noqa: E402

REF_HP = os.path.join(ROOT, "oracle", "_ref", "dftatom_ref_hp")

C1 = dict(levels=14, mixing=0.5, rmax=25.0, delta=0.0005, method=0)
CONFIGS = {
    # small cases the CPU test-suite can re-run in seconds
    "small": [dict(Z=z, levels=10, mixing=0.5, rmax=15.0, delta=0.004, method=m)
              for z, m in [(1, 0), (2, 0), (3, 1), (10, 0), (7, 1), (26, 0), (64, 1), (92, 0)]]
             + [dict(Z=2, levels=12, mixing=0.5, rmax=15.0, delta=0.001, method=0),
                dict(Z=3, levels=12, mixing=0.5, rmax=15.0, delta=0.001, method=1)],
    "argon": [dict(Z=18, **C1)],
    "radon": [dict(Z=86, levels=17, mixing=0.5, rmax=50.0, delta=0.0001, method=1)],
    "sweep": [dict(Z=z, **C1) for z in range(1, 93)],
    "lsda_batch": [dict(Z=z, levels=16, mixing=0.5, rmax=50.0, delta=0.0002, method=1)
                   for z in list(range(21, 31)) + list(range(57, 72))],
    # the uniform-grid pair (DFTAtom.h:15,18; no live caller): method 2 = LDA, 3 = LSDA in dftatom_ref; delta unused
    "uniform": [dict(Z=z, levels=lv, mixing=0.5, rmax=15.0, delta=0.0, method=m)
                for z, lv, m in [(1, 12, 3), (2, 12, 2), (2, 14, 2), (7, 12, 3), (10, 12, 2), (10, 12, 3), (18, 13, 2)]],
}
# How well does the reference reproduce ITSELF?  The same unmodified binary with the mixing parameter changed by one unit in the last
# place (0.5 -> 0.5 (1 + 2^-52)): physically the same calculation, but every rounding downstream differs.  Used by the parity tests to
# tell the reference's own noise floor (its FP64 multigrid sits on a chaotic rounding floor) from a deviation of this implementation.
ULP_MIXING = 0.5 * (1.0 + 2.0 ** -52)
CONFIGS["self_repro"] = ([dict(Z=86, levels=17, mixing=ULP_MIXING, rmax=50.0, delta=0.0001, method=1)]
                         + [dict(Z=z, levels=16, mixing=ULP_MIXING, rmax=50.0, delta=0.0002, method=1) for z in (29, 68, 69, 70)]
                         + [dict(Z=70, levels=14, mixing=ULP_MIXING, rmax=25.0, delta=0.0005, method=0)])
FINAL_ONLY = {"sweep", "lsda_batch", "self_repro"}


def run_one(cfg):
    t0 = time.time()
    # optional cache of raw reference stdout (these runs take up to 20 minutes each): gpurun_out/ref_stdout/<key>.txt
    key = "Z{Z}_L{levels}_m{method}_a{mixing!r}_r{rmax!r}_d{delta!r}".format(**cfg)
    cache = os.path.join(ROOT, "gpurun_out", "ref_stdout", key + ".txt")
    if os.path.exists(cache):
        out = open(cache).read()
    else:
        out = subprocess.run([REF_HP, str(cfg["Z"]), str(cfg["levels"]), repr(cfg["mixing"]), repr(cfg["rmax"]),
                              repr(cfg["delta"]), str(cfg["method"])], capture_output=True, text=True, check=True).stdout
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        with open(cache, "w") as f:
            f.write(out)
    rec = parse_report(out)
    rec["options"] = cfg
    rec["ref_seconds"] = round(time.time() - t0, 2)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name", choices=sorted(CONFIGS))
    ap.add_argument("--jobs", type=int, default=max(1, (os.cpu_count() or 2) - 2))
    a = ap.parse_args()
    cfgs = CONFIGS[a.name]
    # longest first
    order = sorted(range(len(cfgs)), key=lambda i: -cfgs[i]["Z"])
    res = [None] * len(cfgs)
    with ThreadPoolExecutor(a.jobs) as ex:
        futs = {i: ex.submit(run_one, cfgs[i]) for i in order}
        for i, f in futs.items():
            res[i] = f.result()
            print(f"Z={cfgs[i]['Z']} steps={len(res[i]['steps'])} finished={res[i]['finished']} t={res[i]['ref_seconds']}s", flush=True)
    if a.name in FINAL_ONLY:
        for r in res:
            r["etotal_per_step"] = [s["Etotal"] for s in r["steps"]]
            r["energies_per_step"] = [[s[k] for k in ("Etotal", "Ekin", "Ecoul", "Eenuc", "Exc")] for s in r["steps"]]
            r["eig_per_step"] = [[l["E"] for l in s["levels"]] for s in r["steps"]]
            r["steps_kept"] = "last"
            r["n_steps"] = len(r["steps"])
            r["steps"] = r["steps"][-1:]
    meta = dict(generator="scripts/make_goldens.py", source="oracle/_ref/dftatom_ref_hp (unmodified reference, g++ -O2, glibc)",
                name=a.name)
    path = os.path.join(ROOT, "tests", "golden", a.name + ".json")
    with open(path, "w") as f:
        json.dump(dict(meta=meta, atoms=res), f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
