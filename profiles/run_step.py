"""One pass of the headline step (4096 x 4096 grid by default) for profiling under ncu:
    python profiles/run_step.py [N] [assemble|amg|all]
assemble: stamp + CSR build only; amg: + graph-captured AMG-PCG (csrc/dist_amg.cu, one rank)."""
import copy
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from nodal_b200 import dist as ndist
from nodal_b200 import generators as gen
from nodal_b200.device import Device

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
what = sys.argv[2] if len(sys.argv) > 2 else "all"
dev = Device.get(0)
net = copy.deepcopy(gen.grid2d(N))
net.process_component(["a1", "A", "1", "1", "g"])
table = net.table()
dtab = dev.upload_table(table)
torch.cuda.synchronize()
t0 = time.perf_counter()
csr, rhs = dev.assemble_csr(table, dtab=dtab)
torch.cuda.synchronize()
out = {"grid": N, "assemble_ms": (time.perf_counter() - t0) * 1e3, "nnz": csr.nnz}
if what in ("amg", "all"):
    bounds = np.array([0, csr.n], dtype=np.int32)
    x, info = ndist.single_solver(dev).solve_amg(csr.n, bounds, csr.indptr, csr.indices, csr.data, rhs, rtol=1e-10)
    out.update(R=float(x[net.nodenum["1"]]), **{k: info[k] for k in ("iterations", "relres", "setup_ms", "solve_ms")})
print(json.dumps(out), flush=True)
