"""GPU parity: dense LU path (replaces numpy.linalg.solve), batched LU, GMRES."""
import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import nodal_b200 as n
import nodal_b200.equiv
from helpers import block_err, golden, write_csv
from nodal_b200 import _lib
from nodal_b200 import generators as gen
from nodal_b200.device import DeviceCSR
from oracle import mna_oracle as orc

pytestmark = pytest.mark.gpu
DOC = golden("doc_netlists.json")
GRIDS = golden("grids.json")
PINNED = golden("pinned_by_reference_tests.json")


def normwise(x, ref):
    x, ref = np.asarray(x, float), np.asarray(ref, float)
    return np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300)


def lu(device, A, b):
    G = device.to_device(np.ascontiguousarray(A, dtype=np.float64))
    x, info = device.lu_solve(G, device.to_device(np.asarray(b, dtype=np.float64)))
    return x.cpu().numpy(), info


# ------------------------------------------------------------------ dense LU
@pytest.mark.parametrize("m", [1, 2, 3, 5, 17, 64, 127, 128, 129, 200, 257, 513, 1000])
def test_lu_random_systems(device, m):
    rng = np.random.default_rng(m)
    A = rng.standard_normal((m, m))
    if m > 1:
        A[np.arange(0, m, 3), np.arange(0, m, 3)] = 0.0      # zero diagonals: pivoting is mandatory
    b = rng.standard_normal(m)
    x, info = lu(device, A, b)
    assert info["status"] == 0
    want = np.linalg.solve(A, b)
    cond = np.linalg.cond(A)
    assert normwise(x, want) < 1e-13 * cond + 1e-12
    assert np.linalg.norm(A @ x - b) / (np.linalg.norm(A) * np.linalg.norm(x) + np.linalg.norm(b)) < 1e-14


def test_lu_2500_tensor_path(device):
    """Several panels + DMMA trailing updates; backward error at LAPACK level."""
    m = 2500
    rng = np.random.default_rng(0)
    A = rng.standard_normal((m, m)) + np.diag(rng.standard_normal(m))
    b = rng.standard_normal(m)
    x, info = lu(device, A, b)
    assert info["status"] == 0
    assert np.linalg.norm(A @ x - b) / (np.linalg.norm(A) * np.linalg.norm(x)) < 1e-14
    assert normwise(x, np.linalg.solve(A, b)) < 1e-9


def test_lu_singular_reports_pivot(device):
    A = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 6.0], [1.0, 0.0, 1.0]])
    _, info = lu(device, A, np.ones(3))
    assert info["status"] == _lib.SINGULAR and info["info"] == 3
    _, info = lu(device, np.zeros((4, 4)), np.ones(4))
    assert info["status"] == _lib.SINGULAR and info["info"] == 1


def test_lu_nan_input_does_not_fault(device):
    """A NaN component value ('nan' passes float()) must give NaNs / a flagged pivot like LAPACK,
    not an out-of-bounds row swap (n > 128 takes the multi-panel path)."""
    m = 300
    rng = np.random.default_rng(7)
    A = rng.standard_normal((m, m)) + 4 * np.eye(m)
    A[37, :] = np.nan
    A[:, 211] = np.nan
    x, info = lu(device, A, rng.standard_normal(m))
    # LAPACK returns NaNs or flags a pivot (numpy then raises LinAlgError); either is fine here
    assert info["status"] == _lib.SINGULAR or (info["status"] == 0 and np.isnan(x).any())
    # the device is still healthy: the next solve is exact
    B = rng.standard_normal((m, m)) + 4 * np.eye(m)
    b = rng.standard_normal(m)
    y, info = lu(device, B, b)
    assert info["status"] == 0 and normwise(y, np.linalg.solve(B, b)) < 1e-10
    # all-NaN matrix: every panel column has no candidate at all
    x, info = lu(device, np.full((m, m), np.nan), b)
    assert info["status"] == _lib.SINGULAR or np.isnan(x).all()
    y, info = lu(device, B, b)
    assert info["status"] == 0 and normwise(y, np.linalg.solve(B, b)) < 1e-10


@pytest.mark.parametrize("name", sorted(k for k, v in DOC.items() if "result_dense" in v))
def test_doc_netlists_dense_solve(device, name, tmp_path, capsys):
    g = DOC[name]
    net = n.Netlist(write_csv(g["rows"], tmp_path / name))
    sol = n.Circuit(net).solve()
    want = np.array(g["result_dense"])
    # whole-vector normwise: the op-amp fixtures have cond ~1e6 and 1e-12 A currents beside
    # 1 V potentials, where the reference itself moves by 3e-6 between LAPACK builds
    assert normwise(sol.result, want) < 1e-9
    G, A = np.array(g["G"]), np.array(g["A"])
    # block-normwise: potentials and branch currents separately.  A backward
    # stable solve bounds the WHOLE-vector error by ~cond * eps * |x|; a block whose entries are
    # tiny next to the rest (1e-12 A currents beside 1 V potentials) only inherits that absolute
    # error, so its relative bound is cond * eps * |x|_inf / |x_block|_inf (floor 1e-9).
    kcl = len(net.nodenum)
    cond = np.linalg.cond(G)
    got = np.asarray(sol.result, float)
    report = []
    for label, blk in (("potentials", slice(0, kcl)), ("currents", slice(kcl, None))):
        if want[blk].size == 0:
            continue
        scale = np.max(np.abs(want[blk]))
        err = np.max(np.abs(got[blk] - want[blk])) / (scale if scale > 0 else 1.0)
        bound = max(1e-9, 8 * cond * np.finfo(float).eps * np.max(np.abs(want)) / (scale if scale > 0 else 1.0))
        report.append(f"{label} {err:.2e} (bound {bound:.2e})")
        assert err <= bound, (name, label, err, bound, cond)
    with capsys.disabled():
        print(f"\n[{name}] cond {cond:.2e} block-normwise error: " + ", ".join(report))
    assert np.linalg.norm(G @ sol.result - A) <= 1e-10 * max(np.linalg.norm(A), 1e-300) + 1e-9 * 0
    text = str(sol).splitlines()
    ref = g["printed"].splitlines()
    assert text[0] == ref[0]
    assert [l.split(" \t= ")[0] for l in text] == [l.split(" \t= ")[0] for l in ref]
    if name in ("1.6.1.csv", "netlist.csv", "test_1.csv"):
        got = [float(l.split("= ")[1]) for l in text[1:]]
        exp = [float(l.split("= ")[1]) for l in PINNED["printed"][name].splitlines()[1:]]
        assert block_err(got, exp, len(net.nodenum)) < 1e-12      # tests.py:52-61,76-79,106-122


def test_unconnected_and_singular_errors(device, tmp_path):
    net = n.Netlist(write_csv(DOC["unconnected_1.csv"]["rows"], tmp_path / "u1.csv"))
    with pytest.raises(n.UnconnectedCircuitError):                 # nodal.py:328-331
        n.Circuit(net).solve()
    rows = [["e1", "E", "1", "1", "g"], ["e2", "E", "1", "1", "g"], ["r1", "R", "1", "1", "g"]]
    net = n.Netlist(write_csv(rows, tmp_path / "s.csv"))
    with pytest.raises(np.linalg.LinAlgError):                     # connected but singular: re-raise
        n.Circuit(net).solve()


@pytest.mark.parametrize("name", sorted(PINNED["equiv"]))
def test_reference_tests_equivalent_resistance_dense(device, name, tmp_path):
    net = n.Netlist(write_csv(DOC[name]["rows"], tmp_path / name))
    assert n.equiv.equivalent_resistance(net, "1", "g") == pytest.approx(PINNED["equiv"][name], rel=1e-12)


@pytest.mark.parametrize("key", ["grid2d_6", "grid2d_20", "lattice3d_5", "lattice3d_6"])
def test_equivalent_resistance_dense_grids(device, key):
    g = GRIDS[key]
    tn = gen.grid2d(g["N"]) if key.startswith("grid2d") else gen.lattice3d(g["N"])
    assert n.equiv.equivalent_resistance(tn, "1", "g") == pytest.approx(g["R_dense"], rel=1e-9)


def test_config_c3_family_dense(device, tmp_path):
    """Config C3 at a size the oracle redoes in seconds: op-amps, E sources, VCVS."""
    rows = gen.random_opamp_network_rows(M=700, P=60, S=50, V=12, seed=0)
    net = n.Netlist(write_csv(rows, tmp_path / "c3.csv"))
    sol = n.Circuit(net).solve()
    onet, G, A, _, want = orc.solve_rows(rows)
    assert net.nodenum == onet.nodenum
    assert np.linalg.norm(G @ sol.result - A) / np.linalg.norm(A) < 1e-10
    cond = np.linalg.cond(G)
    assert normwise(sol.result[: onet.kcl], want[: onet.kcl]) < max(1e-9, 1e-15 * cond)


# ------------------------------------------------------------------ batched LU (config C4)
def test_batched_opamp_sweep(device, tmp_path):
    net = n.Netlist(write_csv(gen.OPAMP_AMPLIFIER_ROWS, tmp_path / "c4.csv"))
    table = net.table()
    batch = 20000
    vals = gen.opamp_sweep_values(batch, seed=0)
    x, info = device.lu_batched(table, device.to_device(vals))
    x, info = x.cpu().numpy(), info.cpu().numpy()
    assert not info.any()
    # oracle: per-copy csv rows through the reference algorithm, for a sample of the batch
    for s in list(range(0, 40)) + [batch - 1]:
        v1, r1, ri, ro, gain, rf = (float(v) for v in vals[s])
        rows = [["v1", "E", repr(v1), "3", "g"], ["r1", "R", repr(r1), "g", "1"],
                ["q1_ri", "R", repr(ri), "3", "1"], ["q1_ro", "R", repr(ro), "q1_internal_node", "2"],
                ["q1_vcvs", "VCVS", repr(gain), "q1_internal_node", "g", "3", "1"],
                ["q1_rf", "R", repr(rf), "1", "2"]]
        onet, G, A, _, want = orc.solve_rows(rows)
        assert onet.nodenum == net.nodenum and onet.anomnum == net.anomnum
        assert normwise(x[s][:4], want[:4]) < 1e-9
        assert np.linalg.norm(G @ x[s] - A) / np.linalg.norm(A) < 1e-10
    e2 = x[:, net.nodenum["2"]]; e3 = x[:, net.nodenum["3"]]
    assert np.allclose(e2 / e3, 1 + vals[:, 5] / vals[:, 1], rtol=2e-2)     # ideal gain 1 + rf/r1


def test_batched_singular_and_all_types(device, tmp_path):
    net = n.Netlist(write_csv(DOC["test_1.csv"]["rows"], tmp_path / "t1.csv"))
    table = net.table()
    base = table.value.copy()
    vals = np.tile(base, (64, 1)) * np.random.default_rng(0).uniform(0.5, 2, (64, len(base)))
    vals[0] = base
    x, info = device.lu_batched(table, device.to_device(vals))
    x = x.cpu().numpy()
    assert not info.cpu().numpy().any()
    assert normwise(x[0], DOC["test_1.csv"]["result_dense"]) < 1e-12
    from helpers import reduce_triples, stamp_on_host
    from nodal_b200.device import coo_stride
    import copy
    for s in (1, 17, 63):
        t = copy.copy(table); t = t.copy(); t.value[:] = vals[s]
        r, c_, v = stamp_on_host(t, coo_stride(t))
        ip, ix, dt, rhs = reduce_triples(r, c_, v, t.n)
        G = sps.csr_matrix((dt, ix, ip), shape=(t.n, t.n)).toarray()
        assert normwise(x[s], np.linalg.solve(G, rhs)) < 1e-11
    net = n.Netlist(write_csv(DOC["unconnected_1.csv"]["rows"], tmp_path / "u1.csv"))
    table = net.table()
    x, info = device.lu_batched(table, device.to_device(np.tile(table.value, (3, 1))))
    assert (info.cpu().numpy() > 0).all() and np.isnan(x.cpu().numpy()).all()


# ------------------------------------------------------------------ GMRES (non-symmetric sparse)
@pytest.mark.parametrize("name", sorted(k for k, v in DOC.items()
                                        if "result_dense" in v and not v["resistive"]))
def test_doc_netlists_sparse_solve(device, name, tmp_path):
    g = DOC[name]
    net = n.Netlist(write_csv(g["rows"], tmp_path / name))
    sol = n.Circuit(net, sparse=True).solve()
    assert sol.stats["status"] == 0
    assert normwise(sol.result, g["result_dense"]) < 1e-9
    G, A = np.array(g["G"]), np.array(g["A"])
    assert np.linalg.norm(G @ sol.result - A) <= 1e-10 * np.linalg.norm(A)


def test_gmres_grid_with_sources(device, tmp_path):
    """30 x 30 resistor grid + E source + VCVS (the non-symmetric case of SURVEY.md 7.3 item 4)."""
    rows = orc.grid2d_rows(30)
    rows += [["e1", "E", "5", "n3_3", "g"], ["d1", "VCVS", "2", "n20_7", "g", "n5_5", "n6_6"],
             ["a1", "A", "1", "1", "g"]]
    net = n.Netlist(write_csv(rows, tmp_path / "gs.csv"))
    sol = n.Circuit(net, sparse=True, rtol=1e-13).solve()
    assert sol.stats["solver"] == "gmres" and sol.stats["status"] == 0
    onet, G, A, _, want = orc.solve_rows(rows, sparse=True, backend="dict")
    assert np.linalg.norm(G @ sol.result - A) / np.linalg.norm(A) <= 1e-10
    assert block_err(sol.result, want, onet.kcl) < 1e-9


def test_gmres_random_nonsymmetric(device):
    rng = np.random.default_rng(3)
    m = 3000
    A = sps.random(m, m, density=0.002, random_state=rng, format="csr") + sps.diags(rng.uniform(2, 4, m))
    A = sps.csr_matrix(A); A.sort_indices()
    b = rng.standard_normal(m)
    csr = DeviceCSR(m, device.to_device(A.indptr.astype(np.int32)), device.to_device(A.indices.astype(np.int32)),
                    device.to_device(A.data))
    x, info = device.gmres(csr, device.to_device(b), rtol=1e-12, restart=40)
    assert info["status"] == 0
    x = x.cpu().numpy()
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= 1.5e-12
    assert normwise(x, spla.spsolve(A.tocsc(), b)) < 1e-9


def test_sparse_singular_returns_nan_like_reference(device, tmp_path):
    """The reference's -s path returns NaNs with a warning instead of raising."""
    net = n.Netlist(write_csv(DOC["unconnected_1.csv"]["rows"], tmp_path / "u1.csv"))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = n.Circuit(net, sparse=True, maxit=300).solve()
    # singular but consistent: a Krylov solver may return a minimum-residual solution where
    # SuperLU returns NaNs; either way no exception and no hang
    if sol.stats["status"] == 0:
        G, A = np.array(DOC["unconnected_1.csv"]["G"]), np.array(DOC["unconnected_1.csv"]["A"])
        assert np.linalg.norm(G @ sol.result - A) <= 1e-9 * np.linalg.norm(A)
    else:
        assert np.isnan(sol.result).all() or sol.stats["status"] == 2


def test_cli_end_to_end(device, tmp_path, capsys):
    """`nodal-solver FILE [-s]` and `nodal-resistance FILE [-s]` print what the reference prints."""
    from nodal_b200 import equiv, solver
    path = write_csv(DOC["1.6.1.csv"]["rows"], tmp_path / "c1.csv")
    for flags in ([], ["-s"]):
        solver.main([path] + flags)
        lines = capsys.readouterr().out.strip().splitlines()
        ref = PINNED["printed"]["1.6.1.csv"].splitlines()
        assert lines[0] == ref[0] and [l.split(" \t= ")[0] for l in lines] == [l.split(" \t= ")[0] for l in ref]
        got = [float(l.split("= ")[1]) for l in lines[1:]]
        exp = [float(l.split("= ")[1]) for l in ref[1:]]
        assert block_err(got, exp, 3) < 1e-9
    path = write_csv(DOC["resistive_3.csv"]["rows"], tmp_path / "r3.csv")
    for flags in ([], ["-s"]):
        equiv.main([path] + flags)
        out = capsys.readouterr().out.strip()
        assert out.startswith("R = ") and float(out[4:]) == pytest.approx(1.0, rel=1e-9)
    equiv.main([path, "-s", "--precond", "amg"])             # additions to the reference flags
    out = capsys.readouterr().out.strip()
    assert float(out[4:]) == pytest.approx(1.0, rel=1e-9)
    path = write_csv(DOC["unconnected_1.csv"]["rows"], tmp_path / "u1.csv")
    for flags in ([], ["-s", "--check-connected"]):
        with pytest.raises(SystemExit) as e:
            solver.main([path] + flags)
        assert e.value.code == 1


def test_gmres_non_convergence_falls_back_to_dense_lu(device, tmp_path):
    """The reference's -s path is a direct solve: an unconverged GMRES iterate must never be handed
    back as the answer.  With an iteration budget GMRES cannot meet, the system goes through the
    dense LU (it fits) and the result still matches the CPU solve to 1e-9 block-normwise."""
    rows = orc.grid2d_rows(24)
    rows += [["e1", "E", "5", "n3_3", "g"], ["d1", "VCVS", "2", "n20_7", "g", "n5_5", "n6_6"],
             ["a1", "A", "1", "1", "g"]]
    net = n.Netlist(write_csv(rows, tmp_path / "gs.csv"))
    sol = n.Circuit(net, sparse=True, maxit=5).solve()
    assert sol.stats["solver"].startswith("lu") and sol.stats["status"] == 0 and sol.stats["gmres_iterations"] == 5
    onet, G, A, _, want = orc.solve_rows(rows, sparse=True, backend="dict")
    assert block_err(sol.result, want, onet.kcl) < 1e-9
    # above the dense limit the failure is reported (NaNs + warning), not hidden
    import warnings
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        sol = n.Circuit(net, sparse=True, maxit=5, dense_fallback_max_rows=10).solve()
    assert np.isnan(sol.result).all() and any("did not converge" in str(w.message) for w in caught)
