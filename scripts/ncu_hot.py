"""Summarise an ncu report: key metrics + hottest SASS instructions by stall samples. usage: ncu_hot.py file.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
for r in rows[2:3]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name", "")[:100])
    for k in keys:
        if k in d: print(f"  {k} = {d[k]}  [{rows[1][hdr.index(k)]}]")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for idx, r in enumerate(rows):
    if r and r[0] == "Address": hdr = r; start = idx + 1; break
col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[start:]:
    if len(r) < len(hdr): 
        if data: break
        continue
    try: data.append((int(r[col["# Samples"]]), int(r[col["Instructions Executed"]]), r[col["Source"]].strip()))
    except ValueError: pass
tot = sum(x[0] for x in data)
print("total samples", tot, "instructions", len(data))
hot = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
for i in sorted(hot):
    print(f"  [{i:5d}] {100.0*data[i][0]/tot:5.2f}%  exec {data[i][1]:>10d}  {data[i][2][:80]}")
