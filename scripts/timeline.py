"""Per-kernel launch durations along the SCF steps from an ncu launch list (every 8th launch)."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
per = collections.defaultdict(list)
for r in rows:
    name = r[4].split('(')[0].replace('void ', '').split('<')[0]
    per[name].append(float(r[14]) / (1e3 if r[13] == 'ns' else 1.0))
for k, v in per.items():
    if len(v) > 50: print(f"{k:28s} total {sum(v)/1e3:7.1f} ms ", [int(x) for x in v[0:100:8]])
