"""A/B of option variants on the C3 sweep: median wall / device ms of 7 sweeps per variant (k=v,k=v groups on the command line; '' = defaults)."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dftatom_b200 as D
DEFAULTS = {"stream_groups": 4, "rows_wide_from_step": 32, "match_win_until_step": 32, "match_win_nodes": 8192, "use_pdl": 1, "graph_phases": 1, "search_predict": 1,
            "direct_after": 4, "warm_after": 1, "step_cap": 0, "unit_guess": 1}
ctx = D.Context(0)
opts = [D.Options(Z, 14, 25.0, 0.0005, 0.5, 0) for Z in range(1, 93)]
for grp in sys.argv[1:]:
    kv = dict(x.split("=") for x in grp.split(",") if x)
    for k, v in kv.items(): ctx.set_option(k, float(v))
    ctx.solve_batch(opts, keep_steps=False)
    w, d = [], []
    for _ in range(7):
        t0 = time.perf_counter(); res = ctx.solve_batch(opts, keep_steps=False); w.append(1e3 * (time.perf_counter() - t0)); d.append(ctx.last_timing()[0])
    print(f"{grp or 'defaults':50s} wall ms {statistics.median(w):6.2f} dev ms {statistics.median(d):6.2f} finished {sum(r.finished for r in res)}", flush=True)
    for k in kv: ctx.set_option(k, DEFAULTS[k])
