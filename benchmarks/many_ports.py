#!/usr/bin/env python
"""Many-port equivalent resistance: P port pairs on an N x N grid against one
assembled matrix and one AMG hierarchy, batched (csrc/amg_multi.cu, 8 right-hand sides per batch) against
one solve per pair.  One JSON line.

    python benchmarks/many_ports.py [--grid 2048] [--ports 16]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import nodal_b200 as n  # noqa: E402
import nodal_b200.equiv  # noqa: E402
from nodal_b200 import generators as gen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=2048)
ap.add_argument("--ports", type=int, default=16)
args = ap.parse_args()
N = args.grid
net = gen.grid2d(N)
rng = np.random.default_rng(0)
pairs = []
while len(pairs) < args.ports:
    a, b = (tuple(int(v) for v in rng.integers(0, N, 2)) for _ in range(2))
    names = [f"n{p[0]}_{p[1]}" for p in (a, b)]
    if all(name in net.nodenum for name in names) and names[0] != names[1]:
        pairs.append(tuple(names))
out = {"grid": N, "unknowns": net.table().n, "ports": len(pairs)}
for label, kw in (("batched", {}), ("one_by_one", {"multi_rhs": False})):
    for _ in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vals = n.equiv.equivalent_resistances(net, pairs, sparse=True, precond="amg", **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    stats = n.equiv.equivalent_resistances.last_stats
    out[label] = {"s": dt, "ms_per_port": 1e3 * dt / len(pairs), "iterations": [st["iterations"] for st in stats],
                  "worst_relres": max(st["relres"] for st in stats), "solver": stats[0]["solver"], "R0": vals[0]}
out["speedup"] = out["one_by_one"]["s"] / out["batched"]["s"]
out["max_rel_diff"] = float(max(abs(a - b) / abs(b) for a, b in zip(
    n.equiv.equivalent_resistances(net, pairs[:4], sparse=True, precond="amg"),
    n.equiv.equivalent_resistances(net, pairs[:4], sparse=True, precond="amg", multi_rhs=False))))
print(json.dumps(out), flush=True)
