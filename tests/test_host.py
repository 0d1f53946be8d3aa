"""CPU tests of the host logic: C-ABI exports, Aufbau rules of the product library, report format, sharding."""
import ctypes
import os
import re

import numpy as np
import pytest

import dftatom_b200 as D
import oracle_lib as O
from conftest import ROOT, golden
from dftatom_b200.report import format_report, parse_report
from dftatom_b200.shard import partition_atoms


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dftatom_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dftatom_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = D.load_library()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n


def test_no_cpu_fallback():
    """Without a CUDA device creating a context must fail loudly (DFTATOM_E_NO_DEVICE), never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(D.DFTAtomError):
        D.Context(0)


def test_product_aufbau_matches_oracle():
    for Z in range(1, 119):
        assert [(l.n, l.l, l.occ) for l in D.aufbau(Z)] == O.aufbau(Z)
        assert all(l.nodes == l.n - l.l - 1 for l in D.aufbau(Z))
        a, b, ea, eb = D.split_spin(Z)
        oa, ob, oea, oeb = O.split_spin(Z)
        assert [(l.n, l.l, l.occ) for l in a] == oa and [(l.n, l.l, l.occ) for l in b] == ob and (ea, eb) == (oea, oeb)
    assert D.n_nodes(14) == 16385 and D.n_nodes(17) == 131073


def test_report_roundtrip_against_reference_text():
    """format_report(parse_report(reference stdout)) reproduces the reference's text byte for byte (6 decimals)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dftatom_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built")
    import subprocess
    for args, method in ((["3", "10", "0.5", "15", "0.004", "1"], 1), (["10", "10", "0.5", "15", "0.004", "0"], 0)):
        text = subprocess.run([exe] + args, capture_output=True, text=True, check=True).stdout
        rec = parse_report(text)
        steps = [dict(levels=[(l["n"], l["l"], l["E"], l["nodes"]) for l in s["levels"]], **{k: s[k] for k in ("Etotal", "Ekin", "Ecoul", "Eenuc", "Exc")})
                 for s in rec["steps"]]
        out = format_report(rec["Z"], method, steps, rec["finished"], rec["final"]["alpha"], rec["final"].get("beta"))
        assert out.rstrip("\n") == text.rstrip("\n")


def test_partition_atoms_lpt():
    """Shards of the C3 sweep (dftatom_partition: longest-processing-time-first on orbitals x expected SCF steps): every atom exactly
    once, balanced loads, and the atoms that run 100 steps in the reference (Er, Tm, Yb: nearly full nodeless 4f shell) on different ranks."""
    import dftatom_b200 as D
    zs = list(range(1, 93))
    for g in (1, 2, 4, 8):
        parts = partition_atoms(zs, g, method=0)
        assert sorted(sum(parts, [])) == list(range(92))
        loads = [sum(D.estimate_cost(zs[i], 0) for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(D.estimate_cost(z, 0) for z in zs)          # LPT: within one heaviest item
        if g >= 4:
            owners = [next(r for r, p in enumerate(parts) if z - 1 in p) for z in (68, 69, 70)]
            assert len(set(owners)) == 3
    # the cost model: slow convergers weigh more than their neighbours with the same number of orbitals
    assert D.estimate_cost(70, 0) > 1.5 * D.estimate_cost(72, 0) and D.estimate_cost(29, 0) > 1.5 * D.estimate_cost(27, 0)
    assert D.estimate_cost(29, 1) > D.estimate_cost(29, 0)
    # C ABI argument checks
    assert D.estimate_cost(0, 0) == 0.0 and D.estimate_cost(119, 0) == 0.0


def test_cli_fails_loudly_without_gpu():
    """bin/dftatom (the headless replacement of the wx shell) has no CPU path: without a device it exits 1 and says so."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = os.path.join(ROOT, "bin", "dftatom")
    if not os.path.exists(exe):
        pytest.skip("bin/dftatom not built")
    r = subprocess.run([exe, "--Z", "10", "--levels", "10", "--delta", "0.004", "--rmax", "15"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and r.stdout == ""
    assert subprocess.run([exe, "--bogus"], capture_output=True).returncode == 2



def test_python_structs_mirror_the_header():
    """The numpy views solve_batch converts the host structs through (built from the ctypes mirrors of dftatom_result / dftatom_step) have the
    same size and see the same fields at the same places as ctypes does."""
    import ctypes as C
    from dftatom_b200 import api
    assert api._RESULT_DTYPE.itemsize == C.sizeof(api._CResult) and api._STEP_DTYPE.itemsize == C.sizeof(api._CStep)
    n = 3
    cres = (api._CResult * n)()
    cres[1].status = 1; cres[1].n_steps = 7; cres[1].n_spin = 2
    cres[1].n_levels[0] = 2; cres[1].n_levels[1] = 1
    cres[1].levels[0][1].n = 3; cres[1].levels[0][1].l = 1; cres[1].levels[0][1].occ = 4; cres[1].levels[0][1].nodes = 1; cres[1].levels[0][1].E = -1.5
    cres[1].sorted[1][0].occ = 5
    cres[1].Etotal = -9.25
    ra = np.frombuffer(cres, dtype=api._RESULT_DTYPE, count=n)
    assert ra[["status", "n_steps", "n_spin", "Etotal"]].tolist()[1] == (1, 7, 2, -9.25)
    assert ra["n_levels"].tolist()[1] == [2, 1]
    assert ra["levels"].tolist()[1][0][1] == (3, 1, 4, 1, -1.5) and ra["sorted"].tolist()[1][1][0][2] == 5
    cs = (api._CStep * 4)()
    cs[2].E[1][3] = 2.5; cs[2].Etotal = -3.0; cs[2].stop_criterion_met = 1
    sa = np.frombuffer(cs, dtype=api._STEP_DTYPE, count=4)
    assert sa["E"].tolist()[2][1][3] == 2.5 and sa[["Etotal", "stop_criterion_met"]].tolist()[2] == (-3.0, 1)


def test_result_builds_levels_on_first_use():
    """Result keeps the (n, l, occ, nodes, E) records the library returned and turns them into Level objects when asked; it pickles (the
    torch.distributed gather of dftatom_b200/distributed.py sends Results between ranks)."""
    import pickle
    r = D.Result(D.Options(3, 10, 10.0, 0.01, 0.5, 0), 0, 5, [[(1, 0, 2, 0, -1.9), (2, 0, 1, 1, -0.1)]], [[(2, 0, 1, 1, -0.1), (1, 0, 2, 0, -1.9)]],
                 -7.0, 1.0, 2.0, 3.0, 4.0)
    assert r.finished and r.steps == []
    assert [(L.n, L.l, L.occ, L.nodes, L.E) for L in r.levels[0]] == [(1, 0, 2, 0, -1.9), (2, 0, 1, 1, -0.1)]
    assert r.levels is r.levels and isinstance(r.sorted_levels[0][0], D.Level) and r.sorted_levels[0][0].E == -0.1
    q = pickle.loads(pickle.dumps(r))
    assert q.Etotal == -7.0 and q.levels[0][1].nodes == 1 and q.options.Z == 3
