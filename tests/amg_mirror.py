"""TEST-ONLY numpy statement of the aggregation AMG in nodal_b200/csrc/amg.cu.

Used as the checker of the CUDA path (aggregates and coarse operators bit for bit, cycle
and iteration counts to rounding) and to choose the defaults (passes 2, omega 0.8, scale 1.8:
128^2 .. 1024^2 grids converge in 31 .. 45 iterations to 1e-10, Jacobi needs 767 .. 4865).
Nothing in the product imports it.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps


def edge_hash(i, j):
    """csrc/amg_core.cuh: amg_edge_hash (uint32 wrap-around arithmetic)."""
    lo = np.minimum(i, j).astype(np.uint32)
    hi = np.maximum(i, j).astype(np.uint32)
    with np.errstate(over="ignore"):
        h = (lo * np.uint32(2654435761)) ^ (hi * np.uint32(40503) + np.uint32(0x9E3779B9))
        h ^= h >> np.uint32(15)
        h = h * np.uint32(2246822519)
        h ^= h >> np.uint32(13)
    return h


def _edges(A):
    """Off-diagonal entries with a negative value: (row, col, weight, hash)."""
    C = A.tocoo()
    m = (C.row != C.col) & (C.data < 0)
    r, c = C.row[m].astype(np.int64), C.col[m].astype(np.int64)
    return r, c, -C.data[m], edge_hash(r, c).astype(np.int64)


def _preferred(n, r, c, w, h):
    """Per row the edge with the largest (weight, hash, -col); -1 for rows without edges."""
    order = np.lexsort((c, -h, -w, r))
    rs = r[order]
    first = np.r_[True, rs[1:] != rs[:-1]] if len(rs) else np.zeros(0, bool)
    best = np.full(n, -1, dtype=np.int64)
    best[rs[first]] = c[order][first]
    return best


def pairwise_match(A, rounds=8):
    """Handshake matching: match[i] = partner or -1."""
    n = A.shape[0]
    r, c, w, h = _edges(A)
    match = np.full(n, -1, dtype=np.int64)
    idx = np.arange(n)
    for _ in range(rounds):
        free = match < 0
        e = free[r] & free[c]
        if not e.any():
            break
        best = _preferred(n, r[e], c[e], w[e], h[e])
        ok = best >= 0
        mutual = ok & (best[np.where(ok, best, 0)] == idx)
        match[mutual] = best[mutual]
    return match


def aggregates(A, rounds=8):
    """One pairwise pass: (agg, number of aggregates)."""
    n = A.shape[0]
    idx = np.arange(n)
    match = pairwise_match(A, rounds)
    root = np.where(match >= 0, np.minimum(idx, match), idx)
    pref = _preferred(n, *_edges(A))
    lone = (match < 0) & (pref >= 0)
    tgt = np.where(lone, pref, 0)
    joins = lone & (match[tgt] >= 0)
    root = np.where(joins, np.minimum(tgt, match[tgt]), root)
    leader = root == idx
    ids = np.cumsum(leader) - leader
    return ids[root].astype(np.int64), int(leader.sum())


def galerkin(A, agg, nc):
    """P^T A P with the summation order of csrc/csr.cu: entries in CSR order of A, stable sort by
    (agg[row], agg[col]), duplicates added left to right, exact zeros dropped."""
    A = A.tocsr()
    A.sort_indices()
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    key = agg[rows] * (nc + 1) + agg[A.indices]
    order = np.argsort(key, kind="stable")
    key, val = key[order], A.data[order]
    head = np.r_[True, key[1:] != key[:-1]]
    start = np.flatnonzero(head)
    length = np.diff(np.r_[start, len(key)])
    s = val[start].copy()
    for k in range(1, int(length.max()) if len(length) else 0):
        m = length > k
        s[m] = s[m] + val[start[m] + k]
    keep = s != 0.0
    ukey = key[start][keep]
    return sps.csr_matrix((s[keep], (ukey // (nc + 1), ukey % (nc + 1))), shape=(nc, nc))


def coarsen_level(A, passes=2, rounds=8):
    agg = np.arange(A.shape[0])
    Ac = A
    for _ in range(passes):
        a2, na = aggregates(Ac, rounds)
        Ac = galerkin(Ac, a2, na)
        agg = a2[agg]
    return agg, Ac


def hide_cross_couplings(A, owner):
    """Copy of A in which couplings between rows of different owners are invisible to the matching
    (made positive: `_edges` only looks at negative off-diagonals).  Values are NOT used for the
    Galerkin product -- only for deciding aggregates."""
    C = A.tocoo()
    cross = owner[C.row] != owner[C.col]
    return sps.csr_matrix((np.where(cross, np.abs(C.data), C.data), (C.row, C.col)), shape=A.shape)


def coarsen_level_partitioned(A, owner, passes=2, rounds=8):
    """coarsen_level with rank-local aggregation: no aggregate contains rows of two owners (what a
    row-partitioned multi-GPU setup computes without communication beyond halo labels).  Returns
    (agg, Ac, owner of every coarse row); coarse rows of one owner are contiguous."""
    agg = np.arange(A.shape[0])
    Ac, own = A, owner
    for _ in range(passes):
        a2, na = aggregates(hide_cross_couplings(Ac, own), rounds)
        Ac = galerkin(Ac, a2, na)
        coarse_owner = np.zeros(na, dtype=np.int64)
        coarse_owner[a2] = own
        own = coarse_owner
        agg = a2[agg]
    return agg, Ac, own


class AMG:
    """partitions > 1: statement of the multi-GPU hierarchy -- rows are split into `partitions`
    contiguous blocks (nodal_b200.dist.partition_rows); levels with more than `gather_below` rows
    aggregate inside the blocks only, smaller levels are aggregated globally (they are replicated
    on every rank).  partitions == 1 is the single-GPU algorithm of csrc/amg.cu."""

    def __init__(self, A, passes=2, coarse=512, omega=0.8, scale=1.8, maxlevels=30, rounds=8, direct_max=2048,
                 partitions=1, gather_below=20000):
        self.levels, self.omega, self.scale = [], omega, scale
        A = A.tocsr()
        n0 = A.shape[0]
        base, extra = divmod(n0, partitions)
        bounds = np.cumsum([0] + [base + (1 if k < extra else 0) for k in range(partitions)])
        owner = np.searchsorted(bounds, np.arange(n0), side="right") - 1
        while A.shape[0] > coarse and len(self.levels) < maxlevels:
            if partitions > 1 and A.shape[0] > gather_below:
                agg, Ac, next_owner = coarsen_level_partitioned(A, owner, passes, rounds)
            else:
                agg, Ac = coarsen_level(A, passes, rounds)
                next_owner = np.zeros(Ac.shape[0], dtype=np.int64)
            if Ac.shape[0] > 0.9 * A.shape[0]:
                break
            owner = next_owner
            P = sps.csr_matrix((np.ones(A.shape[0]), (np.arange(A.shape[0]), agg)), shape=(A.shape[0], Ac.shape[0]))
            self.levels.append((A, P, 1.0 / A.diagonal(), agg))
            A = Ac
        self.Ac = A
        self.inv = np.linalg.inv(A.toarray()) if A.shape[0] <= direct_max else None
        self.dc = 1.0 / A.diagonal()
        self.rows = [lv[0].shape[0] for lv in self.levels] + [A.shape[0]]
        self.nnz = [lv[0].nnz for lv in self.levels] + [A.nnz]

    def cycle(self, b, l=0):
        if l == len(self.levels):
            return self.inv @ b if self.inv is not None else self.omega * self.dc * b
        A, P, dinv, _ = self.levels[l]
        x = self.omega * dinv * b
        x = x + self.scale * (P @ self.cycle(P.T @ (b - A @ x), l + 1))
        return x + self.omega * dinv * (b - A @ x)

    __call__ = cycle


def pcg(A, b, M, rtol=1e-10, maxit=2000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    bn = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        q = A @ p
        al = rz / (p @ q)
        x += al * p
        r -= al * q
        if np.linalg.norm(r) <= rtol * bn:
            return x, it
        z = M(r)
        rz2 = r @ z
        p = z + (rz2 / rz) * p
        rz = rz2
    return x, maxit
