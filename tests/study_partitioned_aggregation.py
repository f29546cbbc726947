"""TEST-SIDE STUDY (not collected by pytest, not part of the product): how many iterations does
rank-local aggregation cost, i.e. an AMG hierarchy in which no aggregate crosses a boundary of the
contiguous row partition nodal_b200.dist.partition_rows makes?  Design input for the multi-GPU
form of csrc/amg.cu.

    python tests/study_partitioned_aggregation.py

Result (numpy statement amg_mirror.AMG(partitions=P), 1-ohm grids, rtol 1e-10):
    256^2: one block 38 iterations; 2 / 4 / 8 blocks: 38 / 39 / 41
    512^2: one block 35 iterations; 2 / 4 / 8 blocks: 37 / 39 / 40
(with the levels below 20 000 rows aggregated globally: 35 / 36 / 36 and 37 / 37 / 39).
"""
import os
import sys

import numpy as np
import scipy.sparse as sps

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import amg_mirror as mirror  # noqa: E402
from test_amg_host import grid_matrix  # noqa: E402

for N in (256, 512):
    A, b = grid_matrix(N)
    out = [f"N={N} global {mirror.pcg(A, b, mirror.AMG(A))[1]}"]
    for world in (2, 4, 8):
        local = mirror.pcg(A, b, mirror.AMG(A, partitions=world, gather_below=0))[1]
        mixed = mirror.pcg(A, b, mirror.AMG(A, partitions=world, gather_below=20000))[1]
        out.append(f"world {world}: rank-local aggregation on every level {local} | "
                   f"rank-local above 20 000 rows, global below {mixed}")
    print("\n".join(out), flush=True)
