"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]
    t = float(r[14]) / (1e3 if r[13] == "ns" else 1.0)   # -> us
    c = tot.setdefault(name, [0, 0.0, r[7], r[8]])
    c[0] += 1; c[1] += t
total = sum(v[1] for v in tot.values())
# not part of the SCF step: the bench's live FP64-peak microbenchmark, torch's L2-flush fill, the once-per-grid table kernels
aside = lambda k: ("dfma_peak" in k) or k.startswith("void at::") or ("coarse_" in k)
scf = sum(v[1] for k, v in tot.items() if not aside(k))
print(f"| kernel | launches | block | grid | total us | mean us | share | share of the SCF kernels |\n|---|---|---|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {v[0]} | {v[2]} | {v[3]} | {v[1]:.1f} | {v[1]/v[0]:.1f} | {100*v[1]/total:.1f}% | {'-' if aside(k) else f'{100*v[1]/scf:.1f}%'} |")
print(f"\ntotal {total/1e3:.2f} ms over {len(rows)} launches (cold-cache, serialised by ncu: compare shares, not absolutes)")
if len(sys.argv) > 2:      # compact copy of the launch list: id, kernel (no argument list), block, grid, ns
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "block", "grid", "gpu__time_duration_ns"])
        for r in rows:
            w.writerow([r[0], r[4].split("(")[0], r[7], r[8], int(float(r[14]) * (1 if r[13] == "ns" else 1e3))])
