"""CPU checks of the AMG preconditioner's rules (csrc/amg_core.cuh compiled for the host) against
their numpy statement (tests/amg_mirror.py), and of the numpy statement itself."""
import copy

import numpy as np
import pytest
import scipy.sparse as sps

import amg_mirror as mirror
from helpers import amg_aggregate_on_host, amg_host_lib
from nodal_b200 import generators as gen
from oracle.mna_oracle import assemble_resistive_fast


def grid_matrix(N):
    net = copy.deepcopy(gen.grid2d(N))
    t = net.table()
    R = t.type == 0
    A = assemble_resistive_fast(t.a[R], t.b[R], t.value[R], t.n).tocsr()
    b = np.zeros(t.n)
    b[net.nodenum["1"]] = 1.0
    return A, b


def random_network(m, extra, seed, decades=2.0):
    rng = np.random.default_rng(seed)
    i = np.r_[rng.integers(0, m, extra), np.arange(1, m)]
    j = np.r_[rng.integers(0, m, extra), np.arange(0, m - 1)]
    k = i != j
    i, j = i[k], j[k]
    w = 10.0 ** rng.uniform(-decades, decades, len(i))
    L = sps.coo_matrix((np.r_[w, w, -w, -w], (np.r_[i, j, i, j], np.r_[i, j, j, i])), shape=(m, m)).tocsr()
    L = (L + sps.diags(np.where(np.arange(m) % 50 == 0, 1.0, 0.0))).tocsr()
    L.sum_duplicates()
    L.sort_indices()
    return L


def test_edge_hash_matches_numpy():
    lib = amg_host_lib()
    rng = np.random.default_rng(1)
    i = rng.integers(0, 2**31 - 1, 2000)
    j = rng.integers(0, 2**31 - 1, 2000)
    want = mirror.edge_hash(i, j)
    got = np.array([lib.amg_edge_hash_host(int(a), int(b)) for a, b in zip(i, j)], dtype=np.uint32)
    assert (got == want).all()
    assert (mirror.edge_hash(j, i) == want).all()          # symmetric


@pytest.mark.parametrize("case", ["grid24", "grid57", "random", "random_wide", "path", "diagonal"])
def test_aggregation_rules_match_numpy_statement(case):
    if case.startswith("grid"):
        A, _ = grid_matrix(int(case[4:]))
    elif case == "random":
        A = random_network(3000, 9000, 2)
    elif case == "random_wide":
        A = random_network(2000, 3000, 3, decades=6.0)
    elif case == "path":
        A = sps.diags([-np.ones(99), 2.0 * np.ones(100), -np.ones(99)], [-1, 0, 1]).tocsr()
    else:
        A = sps.diags(np.arange(1.0, 41.0)).tocsr()
    match, agg, nc = amg_aggregate_on_host(A)
    m_ref = mirror.pairwise_match(A)
    a_ref, nc_ref = mirror.aggregates(A)
    assert (match == m_ref).all()
    assert nc == nc_ref
    assert (agg == a_ref).all()
    # a matching is an involution without fixed points; aggregates are numbered densely
    paired = match >= 0
    assert (match[match[paired]] == np.flatnonzero(paired)).all()
    assert set(np.unique(agg)) == set(range(nc))
    if case == "diagonal":
        assert nc == A.shape[0]


def test_second_pass_on_coarse_operator_matches():
    A = random_network(3000, 9000, 5)
    agg, nc = mirror.aggregates(A)
    Ac = mirror.galerkin(A, agg, nc)
    match, agg2, nc2 = amg_aggregate_on_host(Ac)
    a_ref, nc_ref = mirror.aggregates(Ac)
    assert nc2 == nc_ref and (agg2 == a_ref).all()


def test_galerkin_statement_equals_triple_product():
    A = random_network(1500, 4000, 7)
    agg, nc = mirror.aggregates(A)
    Ac = mirror.galerkin(A, agg, nc)
    P = sps.csr_matrix((np.ones(A.shape[0]), (np.arange(A.shape[0]), agg)), shape=(A.shape[0], nc))
    ref = (P.T @ A @ P).tocsr()
    assert abs(Ac - ref).max() <= 1e-12 * abs(ref).max()
    assert abs(Ac - Ac.T).max() <= 1e-12 * abs(ref).max()


@pytest.mark.parametrize("N,limit", [(64, 34), (128, 40)])
def test_numpy_amg_pcg_converges_on_grids(N, limit):
    A, b = grid_matrix(N)
    M = mirror.AMG(A)
    x, it = mirror.pcg(A, b, M)
    assert it <= limit
    assert np.linalg.norm(b - A @ x) <= 2e-10 * np.linalg.norm(b)
    # aggregates of ~4-5 rows per level
    assert all(2.5 < a / c < 8 for a, c in zip(M.rows[:-1], M.rows[1:]))


def test_numpy_amg_beats_jacobi_on_irregular_network():
    A = random_network(6000, 18000, 11)
    b = np.random.default_rng(0).standard_normal(A.shape[0])
    x, it = mirror.pcg(A, b, mirror.AMG(A))
    dj = 1.0 / A.diagonal()
    _, itj = mirror.pcg(A, b, lambda r: dj * r, maxit=20000)
    assert it < itj
    assert np.linalg.norm(b - A @ x) <= 2e-10 * np.linalg.norm(b)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_aggregation_statement(world):
    """Design basis of the multi-GPU AMG: aggregates that never cross a row-partition boundary cost
    a handful of iterations, coarse rows stay contiguous per owner."""
    A, b = grid_matrix(128)
    base = mirror.pcg(A, b, mirror.AMG(A))[1]
    M = mirror.AMG(A, partitions=world, gather_below=2000)
    _, it = mirror.pcg(A, b, M)
    assert it <= base + 8
    # first level: every aggregate lies inside one block of partition_rows
    from nodal_b200.dist import partition_rows
    bounds = partition_rows(A.shape[0], world)
    owner = np.searchsorted(bounds, np.arange(A.shape[0]), side="right") - 1
    agg = M.levels[0][3]
    first_owner = np.full(agg.max() + 1, -1)
    first_owner[agg[::-1]] = owner[::-1]
    assert (first_owner[agg] == owner).all()
    assert (np.diff(first_owner) >= 0).all()            # coarse rows are contiguous per owner
    # one partition == the plain statement
    assert mirror.AMG(A, partitions=1).rows == mirror.AMG(A).rows
