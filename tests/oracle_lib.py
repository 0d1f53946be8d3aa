"""TEST INFRASTRUCTURE: ctypes access to the CPU oracle (oracle/libdftatom_oracle.so, the C restatement) and, when
present, to the unmodified reference's own component functions (oracle/_ref/libdftatom_ref.so).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
MAXLV = 24

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class OrcLevel(C.Structure):
    _fields_ = [("n0", C.c_int), ("l", C.c_int), ("occ", C.c_int), ("E", C.c_double)]


class OrcOptions(C.Structure):
    _fields_ = [("Z", C.c_int), ("levels", C.c_int), ("max_r", C.c_double), ("delta", C.c_double), ("mixing", C.c_double), ("method", C.c_int)]


class OrcStep(C.Structure):
    _fields_ = [("step", C.c_int), ("n_levels", C.c_int * 2), ("lv", (OrcLevel * MAXLV) * 2), ("Etotal", C.c_double), ("Ekin", C.c_double),
                ("Ecoul", C.c_double), ("Eenuc", C.c_double), ("Exc", C.c_double), ("level_search_converged", C.c_int)]


class OrcResult(C.Structure):
    _fields_ = [("n_steps", C.c_int), ("finished", C.c_int), ("last", OrcStep), ("n_sorted", C.c_int * 2), ("sorted", (OrcLevel * MAXLV) * 2)]


STEP_CB = C.CFUNCTYPE(None, C.POINTER(OrcStep), C.c_void_p)

_orc = None
_ref = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "restatement"], check=True)


def oracle():
    global _orc
    if _orc is None:
        path = os.path.join(ORACLE_DIR, "libdftatom_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.orc_aufbau.argtypes = [C.c_int, C.POINTER(OrcLevel)]
        L.orc_split_spin.argtypes = [C.c_int, C.POINTER(OrcLevel), C.c_int, C.POINTER(OrcLevel), _ip, C.POINTER(OrcLevel), _ip, _ip, _ip]
        L.orc_n_nodes.argtypes = [C.c_int]
        L.orc_rp.restype = C.c_double
        L.orc_rp.argtypes = [C.c_int, C.c_double, C.c_double]
        L.orc_numerov_start_index.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double]
        L.orc_numerov_count_nodes.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int]
        L.orc_numerov_count_all.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double]
        L.orc_numerov_count_from_nucleus.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int]
        L.orc_numerov_y0.restype = C.c_double
        L.orc_numerov_y0.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double]
        L.orc_numerov_match.restype = C.c_long
        L.orc_numerov_match.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, _dp]
        L.orc_normalize.argtypes = [_dp, C.c_int, C.c_double, C.c_double]
        L.orc_level_search.restype = C.c_double
        L.orc_level_search.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _ip]
        L.orc_simpson38.restype = C.c_double
        L.orc_simpson38.argtypes = [C.c_double, _dp, C.c_int]
        L.orc_integrate.restype = C.c_double
        L.orc_integrate.argtypes = [C.c_int, C.c_double, _dp, C.c_int]
        L.orc_vwn_lda.argtypes = [_dp, C.c_int, _dp, _dp]
        L.orc_xc_chachiyo.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp]
        L.orc_vwn_lsda.argtypes = [_dp, _dp, C.c_int, _dp, _dp, _dp, _dp]
        L.orc_poisson.argtypes = [C.c_int, C.c_double, C.c_int, C.c_double, _dp, _dp, C.c_int, _dp, _ip]
        L.orc_poisson_uniform.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp, C.c_int]
        L.orc_u_count_nodes.argtypes = [_dp, C.c_double, C.c_int, C.c_double, C.c_long, C.c_int]
        L.orc_u_y0.restype = C.c_double
        L.orc_u_y0.argtypes = [_dp, C.c_double, C.c_int, C.c_double, C.c_long]
        L.orc_u_match.restype = C.c_long
        L.orc_u_match.argtypes = [_dp, C.c_double, C.c_int, C.c_double, C.c_long, _dp]
        L.orc_poisson_vcycles.restype = C.c_double
        L.orc_poisson_vcycles.argtypes = [C.c_int, C.c_double, _dp, _dp, C.c_int]
        L.orc_scf.argtypes = [C.POINTER(OrcOptions), C.POINTER(OrcResult), STEP_CB, C.c_void_p, C.c_int]
        L.orc_scf_uniform.argtypes = [C.POINTER(OrcOptions), C.POINTER(OrcResult), STEP_CB, C.c_void_p, C.c_int]
        _orc = L
    return _orc


def ref_components():
    """The unmodified reference's component functions, or None when oracle/_ref was not built."""
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libdftatom_ref.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_numerov_count_nodes.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int]
        if hasattr(L, "ref_numerov_count_from_nucleus"):
            L.ref_numerov_count_from_nucleus.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int]
        L.ref_numerov_solution_in_zero.restype = C.c_double
        L.ref_numerov_solution_in_zero.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double]
        L.ref_numerov_lanes.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, _ip, _dp, _ip, _dp, _ip]
        L.ref_numerov_match.restype = C.c_long
        L.ref_numerov_match.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, _dp]
        L.ref_poisson_nonuniform.argtypes = [C.c_int, C.c_double, C.c_int, C.c_double, _dp, _dp]
        L.ref_vwn_lda.argtypes = [_dp, C.c_int, _dp, _dp]
        if hasattr(L, "ref_numerov_uniform_lanes"):
            L.ref_numerov_uniform_lanes.argtypes = [_dp, C.c_int, C.c_double, C.c_int, _ip, _dp, _ip, _dp, _ip]
            L.ref_numerov_uniform_match.restype = C.c_long
            L.ref_numerov_uniform_match.argtypes = [_dp, C.c_int, C.c_double, C.c_int, C.c_double, _dp]
            L.ref_poisson_uniform.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp]
        if hasattr(L, "ref_xc_chachiyo"):
            L.ref_xc_chachiyo.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp]
        L.ref_vwn_lsda.argtypes = [_dp, _dp, C.c_int, _dp, _dp, _dp, _dp]
        L.ref_simpson38.restype = C.c_double
        L.ref_simpson38.argtypes = [C.c_double, _dp, C.c_int]
        if hasattr(L, "ref_integrate"):
            L.ref_integrate.restype = C.c_double
            L.ref_integrate.argtypes = [C.c_int, C.c_double, _dp, C.c_int]
        L.ref_aufbau.argtypes = [C.c_int, _ip, C.c_int]
        _ref = L
    return _ref


def d(a):
    return a.ctypes.data_as(_dp)


def ip(a):
    return a.ctypes.data_as(_ip)


# ---------------- convenience wrappers over the C restatement ----------------
def aufbau(Z):
    buf = (OrcLevel * MAXLV)()
    n = oracle().orc_aufbau(Z, buf)
    return [(buf[k].n0 + 1, buf[k].l, buf[k].occ) for k in range(n)]


def split_spin(Z):
    allv = (OrcLevel * MAXLV)()
    n = oracle().orc_aufbau(Z, allv)
    a = (OrcLevel * MAXLV)(); b = (OrcLevel * MAXLV)()
    na, nb, ea, eb = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    oracle().orc_split_spin(Z, allv, n, a, C.byref(na), b, C.byref(nb), C.byref(ea), C.byref(eb))
    return ([(a[k].n0 + 1, a[k].l, a[k].occ) for k in range(na.value)], [(b[k].n0 + 1, b[k].l, b[k].occ) for k in range(nb.value)],
            ea.value, eb.value)


def grid(levels, delta, max_r):
    n = oracle().orc_n_nodes(levels)
    rp = oracle().orc_rp(n, delta, max_r)
    i = np.arange(n, dtype=np.float64)
    return n, rp, rp * (np.exp(i * delta) - 1.0)


def numerov_lanes(V, delta, max_r, l, E, limit):
    V = np.ascontiguousarray(V, np.float64)
    y0 = np.array([oracle().orc_numerov_y0(d(V), len(V), delta, max_r, int(li), float(Ei)) for li, Ei in zip(l, E)])
    cnt = np.array([oracle().orc_numerov_count_nodes(d(V), len(V), delta, max_r, int(li), float(Ei), int(k)) for li, Ei, k in zip(l, E, limit)], np.int32)
    return y0, cnt


def numerov_count_from_nucleus(V, delta, max_r, l, E, limit):
    """SolveSchrodingerCountNodesFromNucleus (Numerov.h:204-270) for lanes (l, E, limit)."""
    V = np.ascontiguousarray(V, np.float64)
    return np.array([oracle().orc_numerov_count_from_nucleus(d(V), len(V), delta, max_r, int(a), float(b), int(c)) for a, b, c in zip(l, E, limit)], np.int32)


def numerov_count_all(V, delta, max_r, l, E):
    V = np.ascontiguousarray(V, np.float64)
    return np.array([oracle().orc_numerov_count_all(d(V), len(V), delta, max_r, int(li), float(Ei)) for li, Ei in zip(l, E)], np.int32)


def level_search(V, delta, max_r, Z, n, l, chained=True):
    """Eigenvalues of levels (n,l).  chained=True threads BottomEnergy through the levels like the reference
    (DFTAtom.cpp:407,541); chained=False starts every level from -Z^2-1."""
    V = np.ascontiguousarray(V, np.float64)
    bottom = C.c_double(-float(Z) * Z - 1.0)
    E, ok = [], []
    for nn, ll in zip(n, l):
        if not chained:
            bottom = C.c_double(-float(Z) * Z - 1.0)
        c = C.c_int()
        E.append(oracle().orc_level_search(d(V), len(V), delta, max_r, int(nn) - 1, int(ll), C.byref(bottom), C.byref(c)))
        ok.append(c.value)
    return np.array(E), np.array(ok)


def orbital(V, delta, max_r, l, E):
    V = np.ascontiguousarray(V, np.float64)
    psi = np.zeros(len(V))
    mp = oracle().orc_numerov_match(d(V), len(V), delta, max_r, int(l), float(E), d(psi))
    rp = oracle().orc_rp(len(V), delta, max_r)
    oracle().orc_normalize(d(psi), len(V), rp, delta)
    return psi, int(mp)


def poisson(levels, delta, max_r, Z, rho, max_vcycles=100):
    rho = np.ascontiguousarray(rho, np.float64)
    U = np.zeros_like(rho)
    errs = np.zeros(max(1, max_vcycles)); k = C.c_int()
    oracle().orc_poisson(levels, delta, int(Z), max_r, d(rho), d(U), max_vcycles, d(errs), C.byref(k))
    return U, errs[:k.value]


def uniform_lanes(V, max_r, l, E, limit):
    """(y0, CountNodes) of the uniform-grid sweeps (Numerov.h:272-401, IsUniform() branch) for lanes (l, E, limit)."""
    V = np.ascontiguousarray(V, np.float64)
    n = len(V)
    y0 = np.array([oracle().orc_u_y0(d(V), float(max_r), int(a), float(b), n - 1) for a, b in zip(l, E)])
    cnt = np.array([oracle().orc_u_count_nodes(d(V), float(max_r), int(a), float(b), n - 1, int(c)) for a, b, c in zip(l, E, limit)], np.int32)
    return y0, cnt


def uniform_orbital(V, max_r, l, E):
    V = np.ascontiguousarray(V, np.float64)
    psi = np.zeros_like(V)
    mp = oracle().orc_u_match(d(V), float(max_r), int(l), float(E), len(V) - 1, d(psi))
    return psi, mp


def poisson_uniform(levels, max_r, Z, rho, max_vcycles=100):
    """SolvePoissonUniform (PoissonSolver.h:20-49) on the uniform grid r_i = i MaxR / (N - 1)."""
    rho = np.ascontiguousarray(rho, np.float64)
    U = np.zeros_like(rho)
    oracle().orc_poisson_uniform(int(levels), int(Z), float(max_r), d(rho), d(U), int(max_vcycles))
    return U


def poisson_vcycles(levels, delta, phi, src, n_cycles):
    phi = np.ascontiguousarray(phi, np.float64).copy(); src = np.ascontiguousarray(src, np.float64)
    err = oracle().orc_poisson_vcycles(levels, delta, d(phi), d(src), n_cycles)
    return phi, err


def vwn_lda(rho):
    rho = np.ascontiguousarray(rho, np.float64)
    v = np.zeros_like(rho); e = np.zeros_like(rho)
    oracle().orc_vwn_lda(d(rho), len(rho), d(v), d(e))
    return v, e


def xc_chachiyo(rho, improved=0):
    rho = np.ascontiguousarray(rho, np.float64)
    v = np.zeros_like(rho); e = np.zeros_like(rho)
    oracle().orc_xc_chachiyo(d(rho), len(rho), int(improved), d(v), d(e))
    return v, e


def vwn_lsda(ra, rb):
    ra = np.ascontiguousarray(ra, np.float64); rb = np.ascontiguousarray(rb, np.float64)
    va = np.zeros_like(ra); vb = np.zeros_like(ra); v = np.zeros_like(ra); e = np.zeros_like(ra)
    oracle().orc_vwn_lsda(d(ra), d(rb), len(ra), d(va), d(vb), d(v), d(e))
    return va, vb, v, e


RULES = ("Trapezoid", "SimpsonOneThird", "Simpson38", "Boole", "Romberg")


def integrate(rule, step, v):
    """Integral.h:11-155 by rule index (see RULES)."""
    v = np.ascontiguousarray(v, np.float64)
    return oracle().orc_integrate(int(rule), float(step), d(v), len(v))


def simpson38(step, v):
    v = np.ascontiguousarray(v, np.float64)
    return oracle().orc_simpson38(step, d(v), len(v))


def scf(Z, levels, mixing, max_r, delta, method, max_vcycles=100):
    """Run the oracle SCF; returns dict(steps=[...], finished, n_steps, sorted=[[(n,l,occ)]...])."""
    steps = []

    def cb(st, _):
        s = st.contents
        steps.append(dict(step=s.step,
                          E=[[s.lv[sp][k].E for k in range(s.n_levels[sp])] for sp in range(2)],
                          levels=[[(s.lv[sp][k].n0 + 1, s.lv[sp][k].l, s.lv[sp][k].occ) for k in range(s.n_levels[sp])] for sp in range(2)],
                          Etotal=s.Etotal, Ekin=s.Ekin, Ecoul=s.Ecoul, Eenuc=s.Eenuc, Exc=s.Exc, ok=s.level_search_converged))

    o = OrcOptions(Z, levels, max_r, delta, mixing, method)
    r = OrcResult()
    if method in (2, 3):        # the uniform-grid pair (CalculateUniformLDA / LSDA)
        oracle().orc_scf_uniform(C.byref(o), C.byref(r), STEP_CB(cb), None, max_vcycles)
    else:
        oracle().orc_scf(C.byref(o), C.byref(r), STEP_CB(cb), None, max_vcycles)
    nsp = 2 if method in (1, 3) else 1
    return dict(steps=steps, finished=bool(r.finished), n_steps=r.n_steps,
                sorted=[[(r.sorted[sp][k].n0 + 1, r.sorted[sp][k].l, r.sorted[sp][k].occ) for k in range(r.n_sorted[sp])] for sp in range(nsp)])
