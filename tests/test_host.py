"""Host logic of nodal_b200 (no GPU): surface, numbering, component table, the shared stamp
core compiled for the CPU, generators, and the C-ABI library's exported symbols."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest

import nodal_b200 as n
import nodal_b200.equiv
import nodal_b200.solver
from helpers import ROOT, golden, reduce_triples, stamp_on_host, write_csv
from nodal_b200 import constants as K
from nodal_b200 import generators as gen
from nodal_b200.device import coo_stride
from oracle import mna_oracle as orc

DOC = golden("doc_netlists.json")
GRIDS = golden("grids.json")


@pytest.mark.parametrize("name", sorted(DOC))
def test_numbering_bit_exact(name, tmp_path):
    g = DOC[name]
    net = n.Netlist(write_csv(g["rows"], tmp_path / name))
    assert net.ground == g["ground"]
    assert net.nodenum == g["nodenum"]
    assert net.anomnum == g["anomnum"]
    assert net.nums == g["nums"]
    assert net.component_keys == g["component_keys"]
    assert net.degrees == g["degrees"]
    assert n.equiv.check_resistive(net) == g["resistive"]
    assert n.is_connected(net) == g["connected"]


@pytest.mark.parametrize("name", sorted(DOC))
def test_stamp_core_reproduces_reference_matrices(name, tmp_path):
    """Component table -> (shared stamp core on the CPU) -> sort/reduce model == the
    reference's dense G, A and sorted CSR, bit for bit."""
    g = DOC[name]
    if "G" not in g:
        pytest.skip("reference could not build this netlist")
    net = n.Netlist(write_csv(g["rows"], tmp_path / name))
    table, currents = net.table_and_currents()
    table.validate()
    assert currents == g["currents"]
    stride = coo_stride(table)
    rows, cols, vals = stamp_on_host(table, stride)
    indptr, indices, data, rhs = reduce_triples(rows, cols, vals, table.n)
    want = g["csr_sorted"]
    assert indptr.tolist() == want["indptr"]
    assert indices.tolist() == want["indices"]
    assert data.tolist() == want["data"]
    assert rhs.tolist() == g["A"]
    dense = np.zeros((table.n, table.n))
    for r in range(table.n):
        dense[r, indices[indptr[r]:indptr[r + 1]]] = data[indptr[r]:indptr[r + 1]]
    assert np.array_equal(dense, np.array(g["G"]))


def test_check_input_verdicts():
    import csv, io
    for line, verdict in golden("check_input.json").items():
        row = next(csv.reader(io.StringIO(line), skipinitialspace=True), [])
        if verdict == "ok":
            n.Component.check_input(None, row)
        else:
            with pytest.raises(ValueError):
                n.Component.check_input(None, row)


def test_ground_selection():
    assert n.find_ground_node({"a": 1, "g": 0, "b": 5}) == "g"
    assert n.find_ground_node({"a": 1, "b": 5, "c": 2}) == "b"
    assert n.find_ground_node({"a": 3, "b": 3, "c": 1}) == "a"      # ties -> first inserted


def test_csv_parsing_rules(tmp_path):
    text = "# comment\n\nr1, R, 1e3, a , b\n  r2,R,2,b,g\n#r3,R,1,x,y\n"
    p = tmp_path / "p.csv"
    p.write_text(text)
    net = n.Netlist(str(p))
    onet = orc.OracleNetlist(orc.parse_text(text))
    assert net.nodenum == onet.nodenum == {"a ": 0, "b": 1}   # only leading blanks are stripped
    assert net.component_keys == onet.order
    with pytest.raises(FileNotFoundError):
        n.Netlist(str(tmp_path / "missing.csv"))


def test_error_behaviour(tmp_path):
    with pytest.raises(TypeError):
        n.Circuit("not a netlist")
    net = n.Netlist(write_csv([["r1", "R", "0", "1", "g"]], tmp_path / "z.csv"))
    with pytest.raises(ValueError):
        net.table()
    net = n.Netlist(write_csv([["r1", "R", "1", "1", "g"], ["h", "CCVS", "1", "2", "g", "1", "g", "nope"]],
                              tmp_path / "k.csv"))
    with pytest.raises(KeyError):
        net.table()
    net = n.Netlist(write_csv([["q", "OPAMP", "1", "2", "g", "3", "1"]], tmp_path / "o.csv"))
    with pytest.raises(NotImplementedError):
        net.table()
    net = n.Netlist(write_csv(DOC["1.6.1.csv"]["rows"], tmp_path / "nr.csv"))
    with pytest.raises(ValueError):
        n.equiv.equivalent_resistance(net, "1", "g")
    net = n.Netlist(write_csv(DOC["resistive_1.csv"]["rows"], tmp_path / "r1.csv"))
    with pytest.raises(KeyError):
        n.equiv.equivalent_resistance(net, "1", "zz")


def test_models_signatures_match_reference():
    import inspect
    from nodal_b200 import models
    want = {
        "write_R": ["c", "i", "j", "ground", "G"],
        "write_A": ["c", "i", "j", "ground", "A"],
        "write_E": ["c", "i", "j", "ground", "G", "A", "currents", "anomnum", "nums", "nodenum"],
        "write_VCVS": ["c", "i", "j", "ground", "G", "A", "currents", "anomnum", "nums", "nodenum"],
        "write_VCCS": ["c", "i", "j", "ground", "G", "currents", "anomnum", "nums", "nodenum"],
        "write_CCVS": ["c", "i", "j", "ground", "G", "A", "currents", "anomnum", "nums", "nodenum", "components"],
        "write_CCCS": ["c", "i", "j", "ground", "G", "A", "currents", "anomnum", "nums", "nodenum", "components"],
    }
    for name, params in want.items():
        assert list(inspect.signature(getattr(models, name)).parameters) == params


def test_install_as_nodal():
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == "nodal" or k.startswith("nodal.")}
    try:
        n.install_as_nodal()
        import nodal
        import nodal.equiv
        assert nodal.Circuit is n.Circuit and nodal.equiv.equivalent_resistance is n.equiv.equivalent_resistance
        assert nodal.__version__.startswith("1.3.0")
    finally:
        for k in [k for k in sys.modules if k == "nodal" or k.startswith("nodal.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.parametrize("key", ["grid2d_6", "grid2d_20", "lattice3d_5", "lattice3d_6"])
def test_generated_tables_match_csv_path(key, tmp_path):
    g = GRIDS[key]
    N = g["N"]
    if key.startswith("grid2d"):
        tn, rows = gen.grid2d(N), orc.grid2d_rows(N)
    else:
        tn, rows = gen.lattice3d(N), orc.lattice3d_rows(N)
    net = n.Netlist(write_csv(rows, tmp_path / "g.csv"))
    assert tn.ground == net.ground == "g"
    assert dict(tn.nodenum) == net.nodenum              # identical first-appearance numbering
    assert list(tn.nodenum) == list(net.nodenum)
    t1, t2 = tn.table(), net.table()
    for col in ("type", "value", "a", "b", "c", "d", "drv", "branch"):
        assert np.array_equal(getattr(t1, col), getattr(t2, col)), col
    assert (t1.kcl, t1.be) == (t2.kcl, t2.be) == (g["kcl"], 0)
    # and through the stamp core: the reference's CSR, bit for bit
    probe = copy.deepcopy(tn)
    probe.process_component(["a1", "A", "1", "1", "g"])
    assert len(tn.table()) + 1 == len(probe.table())    # deepcopy did not alias
    rows_, cols_, vals_ = stamp_on_host(probe.table(), 4)
    indptr, indices, data, rhs = reduce_triples(rows_, cols_, vals_, t1.n)
    want = g["csr_sorted"]
    assert indptr.tolist() == want["indptr"] and indices.tolist() == want["indices"]
    assert data.tolist() == want["data"]
    assert rhs[tn.nodenum["1"]] == 1.0 and np.count_nonzero(rhs) == 1
    first = "n0_0" if key.startswith("grid2d") else "n0_0_0"
    assert first in tn.nodenum and tn.nodenum[first] == 0
    assert "n999_0" not in tn.nodenum and "zz" not in tn.nodenum and "g" not in tn.nodenum


def random_network_rows(tn):
    """csv rows of a generated random network (component names r0, r1, ...; node labels of the generator)."""
    t = tn.table()
    label = {int(tn.nodenum[name]): name for name in tn.nodenum}
    label[-1] = tn.ground
    return [[f"r{k}", "R", repr(float(t.value[k])), label[int(t.a[k])], label[int(t.b[k])]] for k in range(len(t))]


@pytest.mark.parametrize("locality", [None, 40])
def test_random_network_numbering_matches_csv_path(locality, tmp_path):
    """The sparse random-network generator numbers its rows exactly as the row-by-row Netlist
    (the reference's first-appearance rule) numbers the same components read from a csv file."""
    tn = gen.random_network(1500, degree=6, seed=3, locality=locality)
    rows = random_network_rows(tn)
    net = n.Netlist(write_csv(rows, tmp_path / "rn.csv"))
    assert net.ground == tn.ground == "g"
    assert dict(tn.nodenum) == net.nodenum and list(tn.nodenum) == list(net.nodenum)
    t1, t2 = tn.table(), net.table()
    for col in ("type", "value", "a", "b", "c", "d", "drv", "branch"):
        assert np.array_equal(getattr(t1, col), getattr(t2, col)), col
    assert t1.kcl == t2.kcl == 1499 and "1" in tn.nodenum and "g" not in tn.nodenum
    # first-appearance numbering is not the node order: the network is irregular on purpose
    ids = np.array([int(name[1:]) for name in list(tn.nodenum)[:200] if name.startswith("n")])
    assert not np.all(np.diff(ids) > 0)


def test_grid_nnz_formula():
    for N in (8, 33):
        t = gen.grid2d(N).table()
        nodes, edges = N * N, 2 * N * (N - 1)
        rows, cols, vals = stamp_on_host(t, 4)
        indptr, _, _, _ = reduce_triples(rows, cols, vals, t.n)
        assert indptr[-1] == nodes + 2 * edges - (1 + 2 * 4)


def test_c3_c4_generators(tmp_path):
    rows = gen.random_opamp_network_rows(M=40, P=6, S=5, V=3, seed=1)
    net = n.Netlist(write_csv(rows, tmp_path / "c3.csv"))
    assert net.nums["kcl"] + net.nums["be"] == 40 + 2 * 5 + 4 * 6 + 2 * 3
    assert 3968 + 2 * 2048 + 4 * 2048 + 2 * 64 == 16384          # default = config C3 size
    t = net.table(); t.validate()
    x = orc.solve_rows(rows)[4]
    assert np.all(np.isfinite(x))
    vals = gen.opamp_sweep_values(16)
    assert vals.shape == (16, 6)
    net = n.Netlist(write_csv(gen.OPAMP_AMPLIFIER_ROWS, tmp_path / "c4.csv"))
    assert net.component_keys == ["v1", "r1", "q1_ri", "q1_ro", "q1_vcvs", "q1_rf"]


def test_shared_library_exports_every_declared_symbol():
    from nodal_b200 import _lib
    header = open(os.path.join(ROOT, "include", "nodal_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(nodal_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()                      # dlopen only -- no CUDA call
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.nodal_abi_version() == 1


def test_no_silent_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nodal_b200._lib import NodalLibraryError
    net = n.Netlist(write_csv(DOC["1.6.1.csv"]["rows"], tmp_path / "a.csv"))
    with pytest.raises(NodalLibraryError):
        n.Circuit(net)


def test_cli_error_paths_without_gpu(tmp_path, capsys):
    """Exit codes / messages of the two console scripts (reference solver.py:19-29, equiv.py:69-83)."""
    from nodal_b200 import equiv, solver
    with pytest.raises(SystemExit) as e:
        solver.main([str(tmp_path / "missing.csv")])
    assert e.value.code == 1
    with pytest.raises(SystemExit) as e:
        equiv.main([str(tmp_path / "missing.csv"), "-s"])
    assert e.value.code == 1
    path = write_csv(DOC["1.6.1.csv"]["rows"], tmp_path / "nr.csv")
    with pytest.raises(SystemExit) as e:
        equiv.main([path])
    assert e.value.code == 1
    out = capsys.readouterr().out
    assert "Invalid netlist" in out and "Resistors are the only component allowed" in out
    path = write_csv([["r1", "R", "1", "a", "b"], ["r2", "R", "1", "b", "c"]], tmp_path / "nonode.csv")
    with pytest.raises(SystemExit) as e:
        equiv.main([path])
    assert e.value.code == 1
    assert "not found in netlist" in capsys.readouterr().out
    assert solver.parser.parse_args(["x.csv", "-s"]).sparse is True
    assert equiv.parser.parse_args(["x.csv"]).sparse is False
    # flags added to the reference command line default to its behaviour
    from nodal_b200.cli import circuit_options
    assert circuit_options(solver.parser.parse_args(["x.csv"])) == {}
    assert circuit_options(solver.parser.parse_args(["x.csv", "-s"])) == {}
    assert circuit_options(solver.parser.parse_args(["x.csv", "--precond", "amg"])) == {}     # dense: ignored
    assert circuit_options(equiv.parser.parse_args(["x.csv", "-s", "--precond", "amg", "--check-connected"])) == \
        {"precond": "amg", "check_connected": True}


def test_table_column_scan_is_cached_and_equals_direct_checks():
    """ComponentTable.facts(): one cached scan answers validate / is_spd_structured / stride."""
    from nodal_b200 import constants as K
    from nodal_b200.device import coo_stride
    from nodal_b200.table import ComponentTable
    rng = np.random.default_rng(0)
    m = 1000
    a = rng.integers(-1, 50, m).astype(np.int32)
    b = rng.integers(-1, 50, m).astype(np.int32)
    t = ComponentTable(np.zeros(m, np.uint8), rng.uniform(1, 2, m), a, b, kcl=50, be=0)
    assert t.is_resistive() and t.is_spd_structured() and t.present_types() == [K.T_R]
    assert coo_stride(t) == 4
    t.validate()
    assert t.facts()["validated"]
    t.value[17] = 0.0                      # edited in place: stale until invalidate()
    t.invalidate()
    assert not t.is_spd_structured()
    with pytest.raises(ValueError, match="null resistance"):
        t.validate()
    t.value[17] = -3.0
    t.invalidate()
    t.validate()                           # the reference stamps negative resistances happily
    assert not t.is_spd_structured()
    t.value[17] = np.nan
    t.invalidate()
    assert not t.is_spd_structured()
    mixed = t.copy()                       # a copy starts without cached facts
    mixed.value[17] = 1.0
    mixed.type[3] = K.T_A
    assert mixed.is_spd_structured() and not mixed.is_resistive()
    assert mixed.present_types() == [K.T_R, K.T_A] and coo_stride(mixed) == 4
    mixed.type[4] = K.T_E
    mixed.invalidate()
    assert not mixed.is_spd_structured() and coo_stride(mixed) == 5
    bad = t.copy()
    bad.a[5] = 50
    with pytest.raises(ValueError, match="out of range"):
        bad.validate()
    empty = ComponentTable(np.zeros(0, np.uint8), np.zeros(0), np.zeros(0, np.int32), np.zeros(0, np.int32))
    empty.validate()
    assert empty.is_resistive() and empty.present_types() == [] and coo_stride(empty) == 2


def test_solution_bulk_access_and_binary_export(tmp_path):
    """Solution.potentials / branch_currents / save: arrays instead of one printed line per node."""
    class FakeNetlist:
        nodenum = {"b": 1, "a": 0, "zz": 2}
        nums = {"kcl": 3, "be": 2}
        ground = "g"
        anomnum = {"e2": 1, "e1": 0}
    result = np.array([1.5, -2.0, 0.25, 3.0, -4.0])
    sol = n.Solution(result, FakeNetlist, ["e1", "e2"])
    assert np.array_equal(sol.potentials(), result[:3])
    assert np.array_equal(sol.branch_currents(), result[3:])
    sol.save(tmp_path / "plain.npz")
    with np.load(tmp_path / "plain.npz") as z:
        assert np.array_equal(z["result"], result) and int(z["kcl"]) == 3 and str(z["ground"]) == "g"
        assert "nodes" not in z
    sol.save(tmp_path / "named.npz", names=True)
    with np.load(tmp_path / "named.npz") as z:
        assert list(z["nodes"]) == ["a", "b", "zz"] and list(z["branches"]) == ["e1", "e2"]
    text = str(sol)                       # the reference layout is untouched
    assert text.splitlines()[0] == "Ground node: g" and "e(zz) \t= 0.25" in text and "i(e2) \t= -4.0" in text


def test_equivalent_resistance_argument_errors_need_no_gpu(tmp_path):
    """ValueError / KeyError of equiv.py:43-48 are raised before anything touches the device, for the
    single-pair and the many-port call, for dict netlists and table netlists."""
    from nodal_b200 import equiv
    from nodal_b200 import generators as gen
    mixed = n.Netlist(write_csv(DOC["1.6.1.csv"]["rows"], tmp_path / "mixed.csv"))
    with pytest.raises(ValueError, match="not resistive"):
        equiv.equivalent_resistances(mixed, [("1", "g")])
    grid = gen.grid2d(8)
    for call in (lambda: equiv.equivalent_resistance(grid, "1", "nowhere", sparse=True),
                 lambda: equiv.equivalent_resistances(grid, [("1", "g"), ("nowhere", "g")])):
        with pytest.raises(KeyError, match="nowhere"):
            call()
    # a table netlist whose ground is not called "g": the reference indexes nodenum with the ground
    # label and fails with KeyError (equiv.py:57-59); the table fast path keeps that
    rows = [["r1", "R", "1", "a", "b"], ["r2", "R", "1", "b", "c"], ["r3", "R", "1", "b", "d"]]
    from nodal_b200.ingest import read_table_netlist
    tn = read_table_netlist(write_csv(rows, tmp_path / "nog.csv"))
    assert tn.ground == "b"
    with pytest.raises(KeyError):
        equiv.equivalent_resistance(tn, "a", "b")
