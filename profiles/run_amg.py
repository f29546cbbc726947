"""Timing of the opt-in AMG-PCG path on 2-D grids (run on the GPU box):
    python profiles/run_amg.py 1024 2048 4096
Prints one JSON line per grid side: hierarchy, setup / solve time, iterations, R.
NODAL_AMG_REPS=1 runs a single (cold) pass, for launch lists under ncu."""
import copy
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from nodal_b200 import generators as gen
from nodal_b200.device import Device

dev = Device.get(0)
for N in [int(a) for a in sys.argv[1:]] or [1024]:
    net = copy.deepcopy(gen.grid2d(N))
    net.process_component(["a1", "A", "1", "1", "g"])
    table = net.table()
    dtab = dev.upload_table(table)
    row = net.nodenum["1"]
    out = {"grid": N, "unknowns": table.n}
    for rep in range(int(os.environ.get("NODAL_AMG_REPS", "2"))):   # first pass warms allocations up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        csr, rhs = dev.assemble_csr(table, dtab=dtab)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        amg = dev.amg(csr)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        x, info = amg.solve(rhs, rtol=1e-10)
        r = float(x[row])
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        amg.close()
    out.update(assemble_ms=(t1 - t0) * 1e3, setup_wall_ms=(t2 - t1) * 1e3, solve_wall_ms=(t3 - t2) * 1e3,
               total_ms=(t3 - t0) * 1e3, R=r, **{k: info[k] for k in (
                   "status", "iterations", "relres", "restarts", "solve_ms", "setup_ms", "levels",
                   "operator_complexity", "grid_complexity", "coarsest_rows", "level_rows")})
    print(json.dumps(out), flush=True)
