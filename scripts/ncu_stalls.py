"""Aggregate an ncu source page: total stall samples by reason, excluding barrier waits, plus per-reason top instructions.
usage: ncu_stalls.py file.ncu-rep [top]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for idx, r in enumerate(rows):
    if r and r[0] == "Address": hdr = r; start = idx + 1; break
col = {}
for i, h in enumerate(hdr): col.setdefault(h, i)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); data = []
for r in rows[start:]:
    if len(r) < len(hdr): continue
    try:
        st = {k: int(r[col[k]]) for k in reasons}
    except ValueError: continue
    for k, v in st.items(): tot[k] += v
    data.append((r[col["Source"]].strip(), int(r[col["# Samples"]]), int(r[col["Instructions Executed"]]), st))
all_s = sum(tot.values())
print("total samples", all_s)
for k, v in tot.most_common(): print(f"  {k:28s} {v:8d} {100.0*v/all_s:5.1f}%")
nb = [(s, n - st["stall_barrier"], e, st) for s, n, e, st in data]
tnb = sum(x[1] for x in nb)
print("non-barrier samples", tnb)
hot = sorted(range(len(nb)), key=lambda i: -nb[i][1])[:top]
for i in sorted(hot):
    s, n, e, st = nb[i]
    main = max((k for k in st if k != "stall_barrier"), key=lambda k: st[k])
    print(f"  [{i:5d}] {100.0*n/tnb:5.2f}%  exec {e:>9d}  {main[6:]:12s} {s[:70]}")
