"""GPU connectivity check (csrc/graph.cu) against scipy.sparse.csgraph and the reference's
is_connected semantics (nodal/nodal.py:88-105)."""
import numpy as np
import pytest
import scipy.sparse as sps
from scipy.sparse.csgraph import connected_components

import nodal_b200 as n
from helpers import golden, write_csv
from nodal_b200 import generators as gen
from nodal_b200.table import ComponentTable

pytestmark = pytest.mark.gpu
DOC = golden("doc_netlists.json")


def reference_components(table):
    """Labels = smallest node index of each component; nodes 0..kcl-1, ground = kcl."""
    kcl = table.kcl
    a = np.where(table.a < 0, kcl, table.a)
    b = np.where(table.b < 0, kcl, table.b)
    g = sps.coo_matrix((np.ones(len(a)), (a, b)), shape=(kcl + 1, kcl + 1))
    count, lab = connected_components(g, directed=False)
    smallest = np.full(count, kcl + 1)
    np.minimum.at(smallest, lab, np.arange(kcl + 1))
    labels = smallest[lab]
    return count, int((labels == labels[kcl]).sum()), labels


def random_table(nodes, ncomp, seed, islands=0):
    rng = np.random.default_rng(seed)
    a = rng.integers(-1, nodes, ncomp).astype(np.int32)
    b = rng.integers(-1, nodes, ncomp).astype(np.int32)
    if islands:                    # cut the node range into blocks that only connect internally
        block = nodes // islands
        b = np.where(a >= 0, (a // block) * block + rng.integers(0, block, ncomp), b).astype(np.int32)
        b = np.minimum(b, nodes - 1)
    return ComponentTable(np.zeros(ncomp, np.uint8), np.ones(ncomp), a, b, kcl=nodes, be=0)


@pytest.mark.parametrize("nodes,ncomp,seed,islands", [(1, 0, 0, 0), (5, 3, 1, 0), (1000, 700, 2, 0), (1000, 5000, 3, 0),
                                                      (50000, 60000, 4, 0), (200000, 900000, 5, 7),
                                                      (300000, 299999, 6, 0)])
def test_components_match_scipy(device, nodes, ncomp, seed, islands):
    table = random_table(nodes, ncomp, seed, islands)
    if seed == 6:                  # one long path in index order: the deepest possible hook chains
        table = ComponentTable(np.zeros(ncomp, np.uint8), np.ones(ncomp), np.arange(ncomp, dtype=np.int32),
                               np.arange(1, ncomp + 1, dtype=np.int32), kcl=nodes, be=0)
    count, reached, labels = device.connected_components(table, want_labels=True)
    want_count, want_reached, want_labels = reference_components(table)
    assert count == want_count
    assert reached == want_reached
    assert np.array_equal(labels.cpu().numpy(), want_labels)
    count2, reached2, none = device.connected_components(table)
    assert (count2, reached2, none) == (count, reached, None)


def test_grid_is_one_component(device):
    table = gen.grid2d(512).table()
    count, reached, _ = device.connected_components(table)
    assert count == 1 and reached == table.kcl + 1


@pytest.mark.parametrize("name", sorted(DOC))
def test_circuit_is_connected_equals_reference_bfs(device, name, tmp_path):
    net = n.Netlist(write_csv(DOC[name]["rows"], tmp_path / name))
    circuit = n.Circuit(net)
    assert circuit.is_connected() == n.is_connected(net)
    if name == "unconnected_1.csv":
        assert not circuit.is_connected()


def test_sparse_path_can_raise_for_unconnected_circuits(device, tmp_path):
    """by default `-s` does not raise (scipy returns NaN + a warning, a
    Krylov solver a finite vector); with check_connected=True the lead graph is checked on the
    device and an unconnected circuit raises like the dense path."""
    rows = [["r1", "R", "1", "1", "g"], ["a1", "A", "1", "1", "g"], ["r2", "R", "1", "2", "3"],
            ["a2", "A", "1", "2", "3"]]
    net = n.Netlist(write_csv(rows, tmp_path / "islands.csv"))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = n.Circuit(net, sparse=True, maxit=200).solve()       # no exception, as with scipy
    assert sol.result.shape == (3,)
    with pytest.raises(n.UnconnectedCircuitError):
        n.Circuit(net, sparse=True, maxit=200, check_connected=True).solve()
    connected = n.Netlist(write_csv(rows[:2], tmp_path / "ok.csv"))
    assert n.Circuit(connected, sparse=True, check_connected=True).solve().result[0] == pytest.approx(1.0)
    # a table netlist (no dict of components) gets the same diagnosis on the dense path
    tn = gen.grid2d(6)
    t = tn.table()
    cut = ComponentTable(t.type[:5], t.value[:5], t.a[:5], t.b[:5], kcl=t.kcl, be=0)
    tn._table = cut
    with pytest.raises(n.UnconnectedCircuitError):
        n.Circuit(tn).solve()
