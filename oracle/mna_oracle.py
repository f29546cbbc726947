"""CPU oracle for nodal's MNA hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU, the algorithm of the reference's hot path
(numbering -> stamping -> solve).  It is the *checker* for the CUDA path in
``nodal_b200``; nothing in the product package imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.

Parity status: PINNED.  ``tests/test_oracle.py`` checks every function here
against (i) the golden values the reference's own tests hold for the path
(``tests.py:15,52-122``) and (ii) dumps produced by importing and running the
reference itself in the build container (``tests/golden/make_golden.py`` ->
``tests/golden/*.json``).

The arithmetic of the solve lives in third-party code that is *not* in the
reference tree: ``numpy.linalg.solve`` (LAPACK dgesv, call site
``nodal/nodal.py:327``) and ``scipy.sparse.linalg.spsolve`` (SuperLU gssv, call
site ``nodal/nodal.py:325``); the reference's dependency spec is the unpinned
``requires = ["scipy"]`` (``pyproject.toml:11``).  The oracle calls the same two
library routines (numpy 2.3.5 / scipy 1.18.1 in this image) so that its results
are the reference's results.

Layout of this file (reference lines each part follows):
  * ``parse_text`` / ``parse_file``          nodal/nodal.py:259-277 (csv reader)
  * ``opmodel_rows``                         nodal/nodal.py:45-85
  * ``pick_ground``                          nodal/nodal.py:30-42
  * ``OracleNetlist``                        nodal/nodal.py:211-296
  * ``assemble``                             nodal/nodal.py:338-398 + nodal/models.py:13-214
  * ``solve_system``                         nodal/nodal.py:313-336
  * ``equivalent_resistance``                nodal/equiv.py:31-61
  * ``grid2d_rows`` / ``lattice3d_rows``     SURVEY.md section 8(d) generators
  * ``assemble_resistive_fast``              vectorised numpy assembler for big resistor grids
"""
from __future__ import annotations

import csv
import io

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla

# nodal/constants.py:4-12 -- CSV columns
NAME, TYPE, VALUE, A_LEAD, B_LEAD, C_CTRL, D_CTRL, DRIVER = range(8)
# nodal/constants.py:15-18
CURRENT_CONTROLLED = ("CCCS", "CCVS")
DEPENDENT = ("VCVS", "VCCS") + CURRENT_CONTROLLED
BRANCH_TYPES = ("E",) + DEPENDENT  # "anomalous": each adds a branch unknown
ALL_TYPES = ("A", "R") + BRANCH_TYPES + ("OPAMP", "OPMODEL")
# nodal/constants.py:20-30
ARITY = {"OPAMP": 7, "OPMODEL": 7, "R": 5, "A": 5, "E": 5,
         "VCCS": 7, "VCVS": 7, "CCCS": 8, "CCVS": 8}
# nodal/constants.py:33-35
OP_RI, OP_RO, OP_GAIN = 1e7, 10, 1e5


def parse_text(text):
    """csv.reader(..., skipinitialspace=True) -- nodal/nodal.py:270."""
    return [row for row in csv.reader(io.StringIO(text), skipinitialspace=True)]


def parse_file(path):
    with open(path, "r") as fh:
        return parse_text(fh.read())


def validate_row(row):
    """Arity / type / numeric checks of nodal/nodal.py:150-178 (raises ValueError)."""
    if len(row) == 0 or row[0][0] == "#":
        return
    name = row[NAME]
    if len(row) < 5:
        raise ValueError(f"Missing arguments for component {name}")
    kind = row[TYPE]
    if kind not in ALL_TYPES:
        raise ValueError(f"Unknown type {kind} for component {name}")
    if len(row) != ARITY[kind]:
        raise ValueError(f"Wrong number of arguments for component {name}")
    try:
        float(row[VALUE])
    except ValueError:
        raise ValueError(f"Bad input: expected a number for component value of {name}")


def opmodel_rows(row):
    """OPMODEL -> ri, ro, vcvs[, rf] primitive rows (nodal/nodal.py:45-85)."""
    name, rf = row[NAME], row[VALUE]
    out, gnd, pos, neg = row[A_LEAD], row[B_LEAD], row[C_CTRL], row[D_CTRL]
    inner = f"{name}_internal_node"
    rows = [
        [f"{name}_ri", "R", str(OP_RI), pos, neg],
        [f"{name}_ro", "R", str(OP_RO), inner, out],
        [f"{name}_vcvs", "VCVS", str(OP_GAIN), inner, gnd, pos, neg],
    ]
    if rf != "0":  # string compare, nodal/nodal.py:80
        rows.append([f"{name}_rf", "R", rf, neg, out])
    else:
        assert neg == out
    return rows


def pick_ground(degrees):
    """'g' if present, else the first node of maximal degree (nodal/nodal.py:30-42)."""
    if "g" in degrees:
        return "g"
    best, best_deg = None, None
    for node, deg in degrees.items():  # insertion order; strict '>' keeps the first max
        if best_deg is None or deg > best_deg:
            best, best_deg = node, deg
    return best


class OracleNetlist:
    """Numbering of nodes and branch unknowns (nodal/nodal.py:211-296)."""

    def __init__(self, rows):
        self.comps = {}       # name -> dict(name,type,value,a,b,c,d,driver)
        self.order = []       # component keys in stamping order (duplicates kept)
        self.degrees = {}
        self.anomnum = {}
        self.n_anom = 0
        deferred = []
        for row in rows:
            self._add(row, deferred)
        for row in list(deferred):   # op-amp parts go after all csv rows (nodal.py:276-277)
            self._add(row, deferred)
        self.finish()

    def _add(self, row, deferred):
        if row == [] or row[0][0] == "#":
            return
        if row[TYPE] == "OPMODEL":
            deferred.extend(opmodel_rows(row))
            return
        validate_row(row)
        kind = row[TYPE]
        comp = dict(name=row[NAME], type=kind, value=float(row[VALUE]),
                    a=row[A_LEAD], b=row[B_LEAD], c=None, d=None, driver=None)
        if kind in DEPENDENT:
            comp["c"], comp["d"] = row[C_CTRL], row[D_CTRL]
            if kind in CURRENT_CONTROLLED:
                comp["driver"] = row[DRIVER]
        self.order.append(row[NAME])
        self.comps[row[NAME]] = comp
        leads = [row[A_LEAD], row[B_LEAD]]
        fresh = [x for x in leads if x not in self.degrees]
        if kind in BRANCH_TYPES:
            self.anomnum[row[NAME]] = self.n_anom
            self.n_anom += 1
        for x in fresh:
            self.degrees[x] = 0
        for x in leads:
            self.degrees[x] += 1

    def add_row(self, row):
        """Late ``process_component`` call as equiv.py:51 makes (counts are not refreshed)."""
        self._add(row, [])

    def finish(self):
        self.ground = pick_ground(self.degrees)
        self.nodenum = {}
        for node in self.degrees:
            if node != self.ground:
                self.nodenum[node] = len(self.nodenum)
        self.kcl = len(self.nodenum)
        self.be = self.n_anom


class _DictMatrix:
    """Minimal dict-of-keys matrix with scipy-DOK semantics: a stored value that
    becomes falsy is deleted (scipy/sparse/_dok.py ``_set_intXint``); insertion
    order of keys is kept, so ``to_csr`` gives first-touch column order."""

    def __init__(self, n):
        self.n = n
        self.d = {}

    def __getitem__(self, ij):
        return self.d.get(ij, 0.0)

    def __setitem__(self, ij, v):
        v = float(v)
        if v:
            self.d[ij] = v
        elif ij in self.d:
            del self.d[ij]

    def to_csr(self):
        n = self.n
        keys = list(self.d.keys())
        rows = np.fromiter((k[0] for k in keys), dtype=np.int64, count=len(keys))
        cols = np.fromiter((k[1] for k in keys), dtype=np.int32, count=len(keys))
        vals = np.fromiter(self.d.values(), dtype=np.float64, count=len(keys))
        perm = np.argsort(rows, kind="stable")
        indptr = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
        return sps.csr_matrix((vals[perm], cols[perm], indptr), shape=(n, n))


def assemble(net, sparse=False, backend="dok"):
    """Build G, A and the branch-name list (nodal/nodal.py:338-398).

    backend: "dok" uses scipy.sparse.dok_matrix exactly as the reference does
    (slow, faithful); "dict" uses ``_DictMatrix`` (same semantics, ~50x faster;
    checked equal to "dok" in tests/test_oracle.py).
    """
    K = net.kcl
    n = K + net.be
    if not sparse:
        G = np.zeros((n, n))
    elif backend == "dok":
        G = sps.dok_matrix((n, n), dtype=np.float64)
    else:
        G = _DictMatrix(n)
    A = np.zeros(n)
    branches = []
    gnd = net.ground
    num = net.nodenum

    for key in net.order:
        comp = net.comps[key]
        kind, val = comp["type"], comp["value"]
        a, b = comp["a"], comp["b"]
        if kind == "R":                                   # models.py:13-24
            try:
                g = 1 / val
            except ZeroDivisionError:
                raise ValueError("Model error: resistors can't have null resistance")
            if a != gnd:
                G[num[a], num[a]] += g
            if b != gnd:
                G[num[b], num[b]] += g
            if a != gnd and b != gnd:
                G[num[a], num[b]] -= g
                G[num[b], num[a]] -= g
        elif kind == "A":                                 # models.py:27-32
            if a != gnd:
                A[num[a]] += val
            if b != gnd:
                A[num[b]] -= val
        elif kind in BRANCH_TYPES:
            r = K + net.anomnum[comp["name"]]
            branches.append(comp["name"])
            if kind == "CCCS":                            # models.py:161-199
                if a != gnd:
                    assert G[num[a], r] == 0
                    G[num[a], r] = -1
                if b != gnd:
                    assert G[num[b], r] == 0
                    G[num[b], r] = 1
                assert G[r, r] == 0
                G[r, r] = 1
                drv = _driver(net, comp)
                if comp["c"] != gnd:
                    G[r, num[comp["c"]]] = +val / drv["value"]
                if comp["d"] != gnd:
                    G[r, num[comp["d"]]] = -val / drv["value"]
                continue
            if kind == "E":
                A[r] += val                               # models.py:39
            if kind == "CCVS":
                drv = _driver(net, comp)                  # lookup precedes the stamps (models.py:116)
            strict = kind != "CCVS"                       # CCVS incidence has no asserts (models.py:126-133)
            if a != gnd:
                if strict:
                    assert G[r, num[a]] == 0
                G[r, num[a]] = 1
                G[num[a], r] = -1
            if b != gnd:
                if strict:
                    assert G[r, num[b]] == 0
                G[r, num[b]] = -1
                G[num[b], r] = 1
            if kind in ("VCVS", "VCCS"):                  # VCCS routed to write_VCVS (nodal.py:377-380)
                if comp["c"] != gnd:
                    G[r, num[comp["c"]]] += -val
                if comp["d"] != gnd:
                    G[r, num[comp["d"]]] += val
            elif kind == "CCVS":                          # assignment, models.py:140-145
                if comp["c"] != gnd:
                    G[r, num[comp["c"]]] = val / drv["value"]
                if comp["d"] != gnd:
                    G[r, num[comp["d"]]] = -val / drv["value"]
        elif kind == "OPAMP":
            raise NotImplementedError
    if sparse:
        G = G.tocsr() if backend == "dok" else G.to_csr()
    return G, A, branches


def _driver(net, comp):
    """Driver lookup + lead check of models.py:115-125 / 177-190; only R drivers work
    in the reference."""
    if comp["driver"] not in net.comps:
        raise KeyError(f"Driving component {comp['driver']} not found")
    drv = net.comps[comp["driver"]]
    assert (comp["c"] == drv["a"] and comp["d"] == drv["b"]) or (
        comp["c"] == drv["b"] and comp["d"] == drv["a"])
    if drv["type"] != "R":
        raise AttributeError("reference only supports R drivers (models.py:146,200)")
    return drv


def solve_system(G, A, sparse=False):
    """nodal/nodal.py:324-327 -- the two library calls the reference makes."""
    if sparse:
        return spla.spsolve(G, A)
    return np.linalg.solve(G, A)


def solve_rows(rows, sparse=False, backend="dok"):
    net = OracleNetlist(rows)
    G, A, branches = assemble(net, sparse=sparse, backend=backend)
    x = solve_system(G, A, sparse=sparse)
    return net, G, A, branches, x


def format_solution(net, x):
    """Solution.__str__ (nodal/nodal.py:422-434)."""
    out = f"Ground node: {net.ground}"
    for name in sorted(net.nodenum):
        out += f"\ne({name}) \t= {x[net.nodenum[name]]}"
    for name in sorted(net.anomnum):
        out += f"\ni({name}) \t= {x[net.kcl + net.anomnum[name]]}"
    return out


def equivalent_resistance(rows, a, b, sparse=False, backend="dok"):
    """nodal/equiv.py:31-61: inject 1 A from b to a, return e(a) - e(b)."""
    net = OracleNetlist(rows)
    if any(c["type"] != "R" for c in net.comps.values()):
        raise ValueError("Network is not resistive")
    for node in (a, b):
        if node not in net.nodenum and node != net.ground:
            raise KeyError(f"Node `{node}` not found in netlist")
    net.add_row(["a1", "A", "1", a, b])      # counts (kcl/be) are not refreshed, equiv.py:51
    G, A, _ = assemble(net, sparse=sparse, backend=backend)
    x = solve_system(G, A, sparse=sparse)
    e = [0, 0]
    for i, node in enumerate((a, b)):
        if node != "g":                       # literal "g", equiv.py:57
            e[i] = x[net.nodenum[node]]
    return e[0] - e[1]


# --------------------------------------------------------------------------- generators
def grid2d_rows(N, resistance="1.0"):
    """N x N grid of equal resistors, probe
    node "1" at (N//2, N//2) and "g" at the knight's move (N//2+2, N//2+1)."""
    def nm(x, y):
        if (x, y) == (N // 2, N // 2):
            return "1"
        if (x, y) == (N // 2 + 2, N // 2 + 1):
            return "g"
        return f"n{x}_{y}"
    rows, k = [], 0
    for x in range(N):
        for y in range(N):
            if x + 1 < N:
                rows.append([f"r{k}", "R", resistance, nm(x, y), nm(x + 1, y)]); k += 1
            if y + 1 < N:
                rows.append([f"r{k}", "R", resistance, nm(x, y), nm(x, y + 1)]); k += 1
    return rows


def lattice3d_rows(N, resistance="1.0"):
    """N^3 lattice, "1" at (N/2,N/2,N/2),
    "g" at (N/2+2, N/2+1, N/2)."""
    h = N // 2
    def nm(x, y, z):
        if (x, y, z) == (h, h, h):
            return "1"
        if (x, y, z) == (h + 2, h + 1, h):
            return "g"
        return f"n{x}_{y}_{z}"
    rows, k = [], 0
    for x in range(N):
        for y in range(N):
            for z in range(N):
                if x + 1 < N:
                    rows.append([f"r{k}", "R", resistance, nm(x, y, z), nm(x + 1, y, z)]); k += 1
                if y + 1 < N:
                    rows.append([f"r{k}", "R", resistance, nm(x, y, z), nm(x, y + 1, z)]); k += 1
                if z + 1 < N:
                    rows.append([f"r{k}", "R", resistance, nm(x, y, z), nm(x, y, z + 1)]); k += 1
    return rows


def assemble_resistive_fast(a_idx, b_idx, value, n):
    """Vectorised CSR assembly of a pure-resistor network given lead indices
    (-1 = ground).  Used for parity at sizes where the per-item DOK loop takes
    minutes.  Column order is sorted; duplicate summation order is scipy's
    (so ``data`` agrees with the DOK path to rounding, not bit-for-bit, on
    entries with >2 contributions)."""
    a_idx = np.asarray(a_idx, dtype=np.int64)
    b_idx = np.asarray(b_idx, dtype=np.int64)
    g = 1.0 / np.asarray(value, dtype=np.float64)
    ma, mb = a_idx >= 0, b_idx >= 0
    mab = ma & mb
    rows = np.concatenate([a_idx[ma], b_idx[mb], a_idx[mab], b_idx[mab]])
    cols = np.concatenate([a_idx[ma], b_idx[mb], b_idx[mab], a_idx[mab]])
    vals = np.concatenate([g[ma], g[mb], -g[mab], -g[mab]])
    G = sps.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    G.sum_duplicates()
    G.sort_indices()
    return G
