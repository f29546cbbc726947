#!/usr/bin/env python
"""Sparse random-network benchmark (north star: "resistor-lattice and random-network netlists at
1, 2, 4 and 8 GPUs"; rows numbered by the reference's first-appearance rule, so the column
pattern and the halo sets of the row partition are irregular).

    python benchmarks/random_network.py [--nodes 16000000] [--degree 8] [--locality W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29514 benchmarks/random_network.py ...

One JSON line (rank 0): assembly time, SELL SpMV time against its algorithmic bytes
(12 nnz + 20 n: every vector element once) on one GPU, and the equivalent-resistance solve
"1" -> "g" with the Jacobi- and the AMG-preconditioned CG (status, iterations, time).
--locality W draws every extra resistor within +-W node ids (a placed netlist); without it the
graph is expander-like: every row block talks to every other one."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from nodal_b200 import _lib  # noqa: E402
from nodal_b200 import dist as ndist  # noqa: E402
from nodal_b200 import generators as gen  # noqa: E402
from nodal_b200.device import Device  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=16_000_000)
    ap.add_argument("--degree", type=int, default=8)
    ap.add_argument("--locality", type=int, default=0)
    ap.add_argument("--decades", type=float, default=2.0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--rtol", type=float, default=1e-10)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = Device.get(local)
    t0 = time.perf_counter()
    net = gen.random_network(args.nodes, args.degree, seed=0, decades=args.decades, locality=args.locality or None)
    table = net.table()
    gen_s = time.perf_counter() - t0
    row_1 = net.nodenum["1"]
    out = {"config": "random network", "nodes": args.nodes, "degree": args.degree, "locality": args.locality or None,
           "n_gpus": world, "unknowns": table.n, "components": len(table), "host_generate_s": gen_s}
    dtab = dev.upload_table(table)
    torch.cuda.synchronize()

    # the probe source: +1 A into "1" (ground is the other end), written into the right-hand side
    def with_probe(rhs, lo):
        if lo <= row_1 < lo + rhs.numel():
            rhs[row_1 - lo] += 1.0
        return rhs

    if world == 1:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dev.assemble_csr(table, dtab=dtab)
        torch.cuda.synchronize()
        ev0.record()
        csr, rhs = dev.assemble_csr(table, dtab=dtab)
        ev1.record()
        torch.cuda.synchronize()
        out.update(assemble_ms=ev0.elapsed_time(ev1), nnz=csr.nnz)
        rhs = with_probe(rhs, 0)
        # ---- SpMV on the solver-private SELL-32 copy
        h = C.c_void_p()
        p = dev.ptr
        _lib.check(dev.lib.nodal_sell_create(dev.ctx, csr.n, csr.nnz, p(csr.indptr), p(csr.indices), p(csr.data),
                                             C.byref(h), dev.stream()), "nodal_sell_create")
        x = torch.rand(csr.n, dtype=torch.float64, device=dev.dev)
        y = torch.empty_like(x)
        for _ in range(3):
            dev.lib.nodal_sell_spmv(dev.ctx, h, p(x), p(y), dev.stream())
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(args.reps):
            dev.lib.nodal_sell_spmv(dev.ctx, h, p(x), p(y), dev.stream())
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / args.reps
        padded = int(dev.lib.nodal_sell_padded_nnz(h))
        dev.lib.nodal_sell_destroy(h)
        alg = 12.0 * csr.nnz + 20.0 * csr.n
        out["spmv"] = {"ms": ms, "algorithmic_bytes": alg, "gbs_algorithmic": alg / ms / 1e6, "sell_padded_nnz": padded,
                       "note": "x gathers follow first-appearance numbering: compare dram__bytes of "
                               "sell_spmv_kernel (ncu) with algorithmic_bytes for the gather amplification"}
        # ---- solves
        for name, fn in (("jacobi", lambda: dev.pcg(csr, rhs, rtol=args.rtol)),
                         ("amg", lambda: dev.amg_pcg(csr, rhs, rtol=args.rtol))):
            try:
                fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                xs, info = fn()
                torch.cuda.synchronize()
                wall = (time.perf_counter() - t0) * 1e3
                resid = dev.spmv(csr, xs) - rhs
                out[name] = {"ms": wall, "status": info["status"], "iterations": info["iterations"], "relres": info["relres"],
                             "relres_checked": float(resid.norm() / rhs.norm()), "R": float(xs[row_1]),
                             "solve_ms": info.get("solve_ms"), "setup_ms": info.get("setup_ms"),
                             "levels": info.get("level_rows")}
            except _lib.NodalLibraryError as exc:
                out[name] = {"error": str(exc)[:300]}
    else:
        bounds = ndist.partition_rows(table.n, world)
        for name in ("jacobi", "amg"):
            runner = ndist.GridRunner(dev, table, row_1, rank, world, rtol=args.rtol, precond=name,
                                      solver=ndist.shared_solver(dev, rank, world))
            lo = int(bounds[rank])
            try:
                res = None
                for _ in range(2):
                    dist.barrier()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    indptr, indices, data, rhs = runner.assemble(dtab)
                    rhs = with_probe(rhs.clone(), lo)
                    torch.cuda.synchronize()
                    t1 = time.perf_counter()
                    if name == "amg":
                        xs, info = runner.pcg.solve_amg(table.n, bounds, indptr, indices, data, rhs, rtol=args.rtol)
                    else:
                        xs, info = runner.pcg.solve(table.n, bounds, indptr, indices, data, rhs, rtol=args.rtol)
                    torch.cuda.synchronize()
                    t2 = time.perf_counter()
                    res = {"assemble_ms": (t1 - t0) * 1e3, "ms": (t2 - t1) * 1e3, "status": info["status"],
                           "iterations": info["iterations"], "relres": info["relres"], "solve_ms": info["solve_ms"],
                           "setup_ms": info["setup_ms"], "halo_recv": info["halo_recv"], "comm": info["comm"]}
                t = torch.tensor([res["assemble_ms"], res["ms"], float(info["halo_recv"])], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                res["assemble_ms"], res["ms"], res["halo_recv_max"] = (float(v) for v in t.tolist())
                r = torch.zeros(1, dtype=torch.float64, device="cuda")
                owner = int(np.searchsorted(bounds, row_1, side="right") - 1)
                if rank == owner:
                    r[0] = xs[row_1 - lo]
                dist.broadcast(r, src=owner)
                res["R"] = float(r.item())
                out[name] = res
            except _lib.NodalLibraryError as exc:
                out[name] = {"error": str(exc)[:300]}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
