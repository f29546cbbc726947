"""TEST-SIDE STUDY (not collected by pytest, not part of the product): how many iterations does
rank-local aggregation cost, i.e. an AMG hierarchy in which no aggregate crosses a boundary of the
contiguous row partition nodal_b200.dist.partition_rows makes?  Design input for the multi-GPU
form of csrc/amg.cu.

    python tests/study_partitioned_aggregation.py

Result (numpy statement, 1-ohm grids, rtol 1e-10):
    256^2: global 38 iterations; 2 / 4 / 8 ranks: 37 / 39 / 40
    512^2: global 35 iterations; 2 / 4 / 8 ranks: 39 / 39 / 41
(with the levels below 20 000 rows aggregated globally: 36 / 35 / 36 and 37 / 38 / 37).
"""
import os
import sys

import numpy as np
import scipy.sparse as sps

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import amg_mirror as mirror  # noqa: E402
from test_amg_host import grid_matrix  # noqa: E402

def restricted(A, owner):
    """copy of A whose cross-owner couplings are hidden from the matching (values kept positive ->
    ignored by _edges, which only looks at negative off-diagonals)"""
    C = A.tocoo()
    cross = owner[C.row] != owner[C.col]
    data = np.where(cross, np.abs(C.data), C.data)
    return sps.csr_matrix((data, (C.row, C.col)), shape=A.shape)

class LocalAMG(mirror.AMG):
    def __init__(self, A, world, passes=2, coarse=512, omega=0.8, scale=1.8, maxlevels=30, gather_below=0):
        self.levels, self.omega, self.scale = [], omega, scale
        A = A.tocsr(); n0 = A.shape[0]
        bounds = np.linspace(0, n0, world + 1).astype(np.int64)
        owner = np.searchsorted(bounds, np.arange(n0), side="right") - 1
        while A.shape[0] > coarse and len(self.levels) < maxlevels:
            agg = np.arange(A.shape[0]); Ac = A; own = owner
            local = A.shape[0] > gather_below
            for _ in range(passes):
                a2, na = mirror.aggregates(restricted(Ac, own) if local else Ac)
                Ac = mirror.galerkin(Ac, a2, na)
                o2 = np.zeros(na, dtype=np.int64); o2[a2] = own; own = o2
                agg = a2[agg]
            if Ac.shape[0] > 0.9 * A.shape[0]: break
            P = sps.csr_matrix((np.ones(A.shape[0]), (np.arange(A.shape[0]), agg)), shape=(A.shape[0], Ac.shape[0]))
            self.levels.append((A, P, 1.0 / A.diagonal(), agg)); A = Ac; owner = own
        self.Ac = A; self.inv = np.linalg.inv(A.toarray()) if A.shape[0] <= 2048 else None; self.dc = 1.0 / A.diagonal()
        self.rows = [lv[0].shape[0] for lv in self.levels] + [A.shape[0]]

for N in (256, 512):
    A, b = grid_matrix(N)
    base = mirror.pcg(A, b, mirror.AMG(A))[1]
    out = [f"N={N} global {base}"]
    for world in (2, 4, 8):
        M = LocalAMG(A, world); it = mirror.pcg(A, b, M)[1]
        M2 = LocalAMG(A, world, gather_below=20000); it2 = mirror.pcg(A, b, M2)[1]
        out.append(f"world {world}: local-all-levels {it} rows {M.rows[-3:]} | local above 20k rows, global below {it2}")
    print("\n".join(out), flush=True)
