"""First GPU check: components + small SCF against the CPU oracle (development aid; the real tests are tests/)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import dftatom_b200 as D
import oracle_lib as O

ctx = D.Context(0)
L, delta, rmax = 12, 0.001, 15.0
N, rp, r = O.grid(L, delta, rmax)
Z = 18
V = np.zeros(N); V[1:] = -Z / r[1:]

# numerov lanes
ls, Es, lim = [], [], []
for n, l in [(1, 0), (2, 0), (2, 1), (3, 0), (3, 1), (3, 2), (4, 3)]:
    En = -Z * Z / (2.0 * n * n)
    for E in np.linspace(1.5 * En, 0.5 * En, 9):
        ls.append(l); Es.append(E); lim.append(n - l - 1)
for E in [-400., -1., 10., 49.]:
    for l in range(4):
        ls.append(l); Es.append(E); lim.append(2)
t0 = time.time()
sign, lg, cnt = ctx.numerov_lanes(V, L, delta, rmax, ls, Es, lim, impl=1)
y0, cnt_o = O.numerov_lanes(V, delta, rmax, ls, Es, lim)
print("numerov lanes:", len(Es), "sign mismatches", int(np.sum(sign != (y0 > 0))), "count mismatches", int(np.sum(cnt != cnt_o)),
      "max |log2 diff|", float(np.nanmax(np.abs(lg - np.log2(np.abs(y0))))))
bad = np.nonzero((sign != (y0 > 0)) | (cnt != cnt_o))[0]
for k in bad[:10]:
    print("  lane", k, "l", ls[k], "E", Es[k], "gpu", sign[k], cnt[k], "cpu", y0[k], cnt_o[k])

# level search
ns = [1, 2, 2, 3, 3, 3, 4]; lls = [0, 0, 1, 0, 1, 2, 3]
E_g, ok = ctx.level_search(V, L, delta, rmax, Z, ns, lls)
E_o, ok_o = O.level_search(V, delta, rmax, Z, ns, lls)
print("level search: max |dE|", float(np.max(np.abs(E_g - E_o))), "ok", ok.tolist(), ok_o.tolist())
print("  E gpu", E_g.tolist())

# orbital
for (n, l), E in zip(zip(ns, lls), E_o):
    u_g, mp_g = ctx.numerov_orbital(V, L, delta, rmax, l, E)
    u_o, mp_o = O.orbital(V, delta, rmax, l, E)
    print(f"orbital {n}{'spdf'[l]}: match {mp_g} vs {mp_o}, max|du| {np.max(np.abs(u_g - u_o)):.3e}")

# poisson
a = 1.7
rho = Z * a ** 3 / np.pi * np.exp(-2 * a * r)
U_g, used = ctx.poisson_solve(L, delta, rmax, [Z], rho)
U_o, errs = O.poisson(L, delta, rmax, Z, rho)
U_x = Z * (1 - np.exp(-2 * a * r) * (1 + a * r))
print("poisson: vcycles", used.tolist(), "max|U_g-U_o|", float(np.max(np.abs(U_g[0] - U_o))), "max|U_o-exact|", float(np.max(np.abs(U_o - U_x))),
      "max|U_g-exact|", float(np.max(np.abs(U_g[0] - U_x))))

# vwn
rr = 10.0 ** np.linspace(-20, 4, 200)
v_g, e_g = ctx.vwn(rr); v_o, e_o = O.vwn_lda(rr)
print("vwn lda: max rel", float(np.max(np.abs(v_g - v_o) / (np.abs(v_o) + 1e-300))), float(np.max(np.abs(e_g - e_o) / (np.abs(e_o) + 1e-300))))
rb = rr * np.linspace(0, 1, 200)
g4 = ctx.vwn(rr, rb); o4 = O.vwn_lsda(rr, rb)
print("vwn lsda: max rel", [float(np.max(np.abs(x - y) / (np.abs(y) + 1e-300))) for x, y in zip(g4, o4)])
# simpson
v = np.random.default_rng(0).standard_normal((3, N))
print("simpson:", ctx.simpson38(1.0, v) - np.array([O.simpson38(1.0, row) for row in v]))

# SCF
for (Zz, LL, dd, rm, meth) in [(2, 12, 0.001, 15.0, 0), (10, 10, 0.004, 15.0, 0), (3, 12, 0.001, 15.0, 1), (26, 10, 0.004, 15.0, 0), (64, 10, 0.004, 15.0, 1)]:
    t0 = time.time()
    res = ctx.solve_batch([D.Options(Zz, LL, rm, dd, 0.5, meth)])[0]
    t1 = time.time()
    o = O.scf(Zz, LL, 0.5, rm, dd, meth)
    ms, nl = ctx.last_timing()
    ns_ = min(res.n_steps, o["n_steps"])
    dE = max(abs(res.steps[k].Etotal - o["steps"][k]["Etotal"]) for k in range(ns_))
    de = max(abs(a_ - b_) for k in range(ns_) for s in range(len(res.levels)) for a_, b_ in zip(res.steps[k].E[s], o["steps"][k]["E"][s]))
    print(f"SCF Z={Zz} L={LL} m={meth}: steps gpu {res.n_steps} cpu {o['n_steps']} status {res.status} fin {o['finished']} "
          f"Etot {res.Etotal:.9f} vs {o['steps'][-1]['Etotal']:.9f} max|dEtot| over steps {dE:.2e} max|deig| {de:.2e} wall {t1-t0:.2f}s dev {ms:.1f} ms launches {nl}")
