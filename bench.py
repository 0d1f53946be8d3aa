#!/usr/bin/env python
"""bench.py — atoms/sec of converged SCF on the periodic-table sweep (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU solver on the host cores

A "step" is one pass of the hot path over one batch: the whole C3 workload (Z = 1..92, LDA, 14 multigrid levels =
16385 nodes, delta 0.0005, mixing 0.5, Rmax 25) solved to the reference's stop criterion.  Weak scaling:
every rank solves one full sweep (atoms are independent, no collective on the data path), value = N*92 / max-rank time.
The strong-scaling figure (the same 92 atoms sharded over the ranks) is reported beside it under "strong_c3".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3 = dict(levels=14, delta=0.0005, mixing=0.5, rmax=25.0, method=0)
WORKLOAD = "C3 periodic-table sweep Z=1-92 LDA, 14 levels (16385 nodes), delta 0.0005, mixing 0.5, Rmax 25"
FLOP_PER_NODE_STEP = 11.0          # SURVEY §8(d) accounting convention for the Numerov shooting kernel
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "dftatom_ref")
ORACLE_EXE = os.path.join(ROOT, "oracle", "dftatom_oracle")


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # under load = samples above the idle clock
        load = [x for x in sm if x > 0.5 * max(smax or [0])] or sm
        return dict(sm_mhz=statistics.median(load) if load else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------------------
# CPU reference arm (also the cpu_baseline of the CUDA arm)
# ------------------------------------------------------------------------------------------------------------
def _golden_costs():
    """Per-Z single-core seconds of the unmodified reference on C3, recorded when tests/golden/sweep.json was made."""
    with open(os.path.join(ROOT, "tests", "golden", "sweep.json")) as f:
        g = json.load(f)
    return {a["options"]["Z"]: float(a["ref_seconds"]) for a in g["atoms"]}


def cpu_reference_sample(budget_s, cores=None):
    """Run the reference's own CPU solver (oracle/_ref/dftatom_ref = unmodified reference compiled headless; falls back to
    the C restatement oracle/dftatom_oracle) on a stratified sample of C3 atoms, one single-threaded process per atom
    (the solver is single-threaded), all host cores in use.  The sweep throughput is then the LPT bound
    92 / max(sum_cost/cores, max_cost) with every atom's cost scaled by measured/recorded time of the sample."""
    cores = cores or os.cpu_count() or 1
    exe, kind = (REF_EXE, "reference") if os.path.exists(REF_EXE) else (ORACLE_EXE, "port")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "restatement"], check=True)
    cost = _golden_costs()
    # stratified in Z; the recorded costs were taken on a slower, shared VM, so they over-estimate
    cand = [z for z in range(4, 93, 8) if cost[z] <= budget_s] or [min(cost, key=cost.get)]
    sample = cand[:cores]
    t0 = time.time()
    procs = [(z, time.time(), subprocess.Popen([exe, str(z), str(C3["levels"]), str(C3["mixing"]), str(C3["rmax"]), str(C3["delta"]), "0"],
                                               stdout=subprocess.DEVNULL)) for z in sample]
    secs = {}
    for z, ts, p in procs:
        p.wait()
        secs[z] = time.time() - ts
    wall = time.time() - t0
    ratio = sum(secs.values()) / sum(cost[z] for z in sample)
    total = sum(cost.values()) * ratio
    longest = max(cost.values()) * ratio
    sweep_s = max(total / cores, longest)
    return dict(value=92.0 / sweep_s, unit="atoms/s", cores=cores, kind=kind,
                sample=f"Z={sample} of C3 run concurrently ({wall:.1f}s wall, {sum(secs.values()):.1f} core-s); per-Z costs of the full sweep "
                       f"scaled by {ratio:.3f} => {total:.0f} core-s, longest atom {longest:.1f}s; value = 92/max(core-s/cores, longest)",
                sweep_core_seconds=total, per_core_atoms_per_s=92.0 / total, wall_s=wall)


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = max(8.0, 240.0 / max(1, a.steps + a.warmup))
    for _ in range(a.warmup):
        cpu_reference_sample(budget)
    vals, last = [], None
    t0 = time.time()
    for _ in range(a.steps):
        last = cpu_reference_sample(budget)
        vals.append(last["value"])
    ms = (time.time() - t0) * 1e3 / max(1, a.steps)
    v = statistics.mean(vals)
    line = dict(impl="reference", metric="atoms/sec converged SCF (Z=1-92 LDA, 16385 nodes)", value=v, unit="atoms/s", n_gpus=a.gpus,
                steps=a.steps, warmup=a.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", config=dict(workload=WORKLOAD, note="CPU: bounded stratified sample per step, see cpu_baseline.sample"),
                cpu_baseline=dict(value=v, unit="atoms/s", cores=last["cores"], kind=last["kind"], sample=last["sample"]),
                e2e=dict(value=v, unit="atoms/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------------------
def run_cuda_arm(a):
    import torch
    import dftatom_b200 as D
    from dftatom_b200.shard import partition_atoms

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (dftatom_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = D.Context(local)
    ctx.set_option("profile", 1)
    opts = [D.Options(Z, C3["levels"], C3["rmax"], C3["delta"], C3["mixing"], C3["method"]) for Z in range(1, 93)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    for _ in range(a.warmup):
        flush.zero_()
        res = ctx.solve_batch(opts, keep_steps=False)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms = 0.0
    launches = 0
    prof = {k: dict(ms=0.0, launches=0, work=0.0) for k in D.api.KERNEL_CLASSES}
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        flush.zero_()
        res = ctx.solve_batch(opts, keep_steps=False)      # host options in, host results out: the e2e path
        ms, nl = ctx.last_timing()
        dev_ms += ms
        launches += nl
        for k, v in ctx.last_profile().items():
            for f in ("ms", "launches", "work"):
                prof[k][f] += v[f]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    wall = max_over_ranks(wall)
    dev_s = max_over_ranks(dev_ms * 1e-3)
    n_atoms_total = len(opts) * world * a.steps
    n_finished = int(sum_over_ranks(sum(r.finished for r in res)))
    opt_bytes = 40 * len(opts)
    res_bytes = (5 * 4 + 2 * 24 * 24 * 2 + 5 * 8) * len(opts)

    # strong scaling on C3 as given: the same 92 atoms sharded over the ranks (LPT by orbital count)
    strong = None
    if world > 1:
        mine = partition_atoms([o.Z for o in opts], world)[rank]
        my_opts = [opts[i] for i in mine]
        ctx.solve_batch(my_opts, keep_steps=False)
        barrier()
        t1 = time.perf_counter()
        ctx.solve_batch(my_opts, keep_steps=False)
        barrier()
        tw = max_over_ranks(time.perf_counter() - t1)
        strong = dict(value=92.0 / tw, unit="atoms/s", scaling="strong", atoms=92, seconds=tw)

    if rank == 0:
        peak = ctx.measure_fp64_peak()
        s = prof["search"]
        achieved = FLOP_PER_NODE_STEP * s["work"] / (s["ms"] * 1e-3) / 1e12 if s["ms"] > 0 else 0.0
        shares = {k: (v["ms"] / (dev_ms or 1.0)) for k, v in prof.items()}
        line = dict(
            metric="atoms/sec converged SCF (Z=1-92 LDA, 16385 nodes)", value=n_atoms_total / dev_s, unit="atoms/s", n_gpus=world,
            steps=a.steps, warmup=a.warmup, ms_per_step=wall * 1e3 / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="f64", data="synthetic",
            config=dict(workload=WORKLOAD, atoms_per_gpu=92, l2="flushed between steps (256 MiB memset)", atoms_converged=n_finished,
                        note="every rank solves one full sweep; 89/92 atoms meet the reference's stop test, Z=68-70 run to the 100-step cap like the reference"),
            e2e=dict(value=n_atoms_total / wall, unit="atoms/s", h2d_bytes_per_step=opt_bytes, d2h_bytes_per_step=res_bytes),
            gpu_launches=int(launches),
            roofline=dict(kernel="search_fused_kernel + search_seg_kernel (Numerov shooting: Sturm-count search, serial-in-r / parallel-in-r)",
                          bound="fp64", achieved=achieved, peak=peak, unit="TFLOP/s",
                          frac=achieved / peak if peak else None, traffic=None,
                          peak_source="measured live: DFMA microbench in libdftatom_b200 (MEASURED_PEAKS.json has no FP64 entry)",
                          flop_per_lane_node_step=FLOP_PER_NODE_STEP, lane_node_steps=s["work"], kernel_ms=s["ms"], share_of_step=shares),
            kernels={k: dict(ms=v["ms"], launches=int(v["launches"]), work=v["work"]) for k, v in prof.items()},
            search=dict(orbital_solves=prof["match"]["work"], rounds_per_solve=prof["density"]["work"] / max(1.0, prof["match"]["work"]),
                        inward_sweeps_per_solve_reference=140, note="one round = 32 concurrent inward sweeps"),
            poisson=dict(solves=int(prof["poisson"]["launches"]) * len(opts), gs_node_updates=prof["poisson"]["work"], ms=prof["poisson"]["ms"],
                         gs_updates_per_s=prof["poisson"]["work"] / (prof["poisson"]["ms"] * 1e-3) if prof["poisson"]["ms"] else None,
                         bound="shared memory / L2 latency (grid resident on chip at 16385 nodes; compulsory HBM traffic 16 N B per solve)"),
            clocks=clocks,
        )
        if strong:
            line["strong_c3"] = strong
        if world == 1 and not a.no_batch:
            # the same sweep replicated 8x in ONE batch (736 atoms): C3 as given is bounded by the SCF chain of its slowest atoms
            # (from step ~35 on fewer than 30 atoms are left), a larger batch shows what the kernels sustain when the GPU is full
            big = opts * 8
            ctx.solve_batch(big, keep_steps=False)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            ctx.solve_batch(big, keep_steps=False)
            tb = time.perf_counter() - t1
            line["batch_8xC3"] = dict(value=len(big) / tb, unit="atoms/s", atoms=len(big), seconds=tb, device_ms=ctx.last_timing()[0],
                                      note="8 copies of the Z=1-92 sweep in one dftatom_solve_batch call, host options in, host results out")
        if world == 1 and not a.no_rn:
            # second half of BASELINE.json's metric: wall-ms of one Radon SCF (C2: Z=86 LSDA, 17 levels = 131073 nodes, delta 1e-4,
            # mixing 0.5, Rmax 50) through the same public call, host options in, host results out; warm-up run first
            rn = [D.Options(86, 17, 50.0, 0.0001, 0.5, 1)]
            ctx.solve_batch(rn, keep_steps=False)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            r_rn = ctx.solve_batch(rn, keep_steps=False)[0]
            rn_ms = (time.perf_counter() - t1) * 1e3
            line["rn_scf"] = dict(metric="Rn SCF ms", value=rn_ms, unit="ms", device_ms=ctx.last_timing()[0], scf_steps=r_rn.n_steps, finished=bool(r_rn.finished),
                                  Etotal=r_rn.Etotal, workload="C2 Radon Z=86 LSDA, 17 levels (131073 nodes), delta 0.0001, mixing 0.5, Rmax 50",
                                  reference_cpu_seconds_1core=518.0, reference_source="SURVEY.md section 6 (unmodified reference, g++ -O2, one core)")
        if world == 1 and not a.no_cpu_baseline:
            cb = cpu_reference_sample(30.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rn", action="store_true", help="skip the Radon (C2) SCF timing")
    ap.add_argument("--no-batch", action="store_true", help="skip the replicated-batch (8 x C3) throughput figure")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_cuda_arm(a)


if __name__ == "__main__":
    main()
