#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs that are not the bench.py headline:

  C1  doc/1.6.1.csv-sized dense solve (latency)
  C2  1000 x 1000 grid equivalent resistance, sparse PCG, 1 GPU
  C3  dense random network with op-amps / VCVS / E sources, 16 384 unknowns, blocked LU
  C4  1 M copies of the op-amp amplifier, batched LU
  C5b 256^3 lattice (optional, --c5b)

Each line printed is a JSON object; numbers are CUDA-event timed after warm-up.  The CPU
column is numpy / scipy on the host (the reference's arithmetic), measured in the same run.

    python benchmarks/configs.py [--c3-size 16384] [--c4-batch 1000000] [--c5b] [--no-cpu]
"""
import argparse
import copy
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import nodal_b200 as n  # noqa: E402
import nodal_b200.equiv  # noqa: E402
from nodal_b200 import generators as gen  # noqa: E402
from nodal_b200.device import Device  # noqa: E402


def timed(fn, warmup=1, reps=3):
    for _ in range(warmup):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def write_csv(rows):
    import csv
    fd, path = tempfile.mkstemp(suffix=".csv")
    with os.fdopen(fd, "w", newline="") as fh:
        csv.writer(fh).writerows(rows)
    return path


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-size", type=int, default=16384)
    ap.add_argument("--c4-batch", type=int, default=1_000_000)
    ap.add_argument("--c5b", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only-c4", action="store_true", help="batched LU only (both value layouts)")
    args = ap.parse_args()
    dev = Device.get(0)
    if args.only_c4:
        net = n.Netlist(write_csv(gen.OPAMP_AMPLIFIER_ROWS))
        table = net.table()
        vals = gen.opamp_sweep_values(args.c4_batch, seed=0)
        d_aos = dev.to_device(vals)
        d_soa = dev.to_device(np.ascontiguousarray(vals.T))
        ms_a, (xa, ia) = timed(lambda: dev.lu_batched(table, d_aos), warmup=2, reps=10)
        ms_s, (xs, i_s) = timed(lambda: dev.lu_batched(table, d_soa, layout="soa"), warmup=2, reps=10)
        same = bool(torch.equal(xa, xs.t().contiguous()))
        emit(config="C4", batch=args.c4_batch, aos_ms=ms_a, soa_ms=ms_s, aos_gbs=96.0 * args.c4_batch / ms_a / 1e6,
             soa_gbs=96.0 * args.c4_batch / ms_s / 1e6, soa_equals_aos_bitwise=same,
             singular=int((ia != 0).sum().item()) + int((i_s != 0).sum().item()))
        return

    # ---- C1: latency of the smallest dense case
    rows = [["r1", "R", "2", "1", "4"], ["r2", "R", "2", "1", "g"], ["r3", "R", "0.5", "1", "2"],
            ["e1", "E", "8", "4", "g"], ["a1", "A", "4", "1", "2"], ["d1", "CCCS", "2", "2", "g", "1", "g", "r2"]]
    net = n.Netlist(write_csv(rows))
    t0 = time.perf_counter()
    for _ in range(20):
        sol = n.Circuit(net).solve()
    emit(config="C1", what="Circuit(netlist).solve() wall, dense 5 unknowns", ms=(time.perf_counter() - t0) / 20 * 1e3,
         result=sol.result.tolist())

    # ---- C2
    tn = gen.grid2d(1000)
    probe = copy.deepcopy(tn)
    probe.process_component(["a1", "A", "1", "1", "g"])
    table = probe.table()
    dtab = dev.upload_table(table)

    def c2():
        csr, rhs = dev.assemble_csr(table, dtab=dtab)
        x, info = dev.pcg(csr, rhs, rtol=1e-10)
        return float(x[probe.nodenum["1"]]), info, csr.nnz
    ms, (r, info, nnz) = timed(c2)
    emit(config="C2", what="1000x1000 grid: stamp + CSR build + PCG", ms=ms, R=r, golden_R=0.7732422803670024,
         rel_err=abs(r - 0.7732422803670024) / 0.7732422803670024, iterations=info["iterations"],
         relres=info["relres"], pcg_ms=info["solve_ms"], unknowns=table.n, nnz=nnz,
         pcg_gbs=(12 * nnz + 108 * table.n) * info["iterations"] / info["solve_ms"] / 1e6,
         note="60 MB CSR fits the 126 MB L2: this GB/s figure is L2-assisted, not an HBM figure",
         cpu_reference_s={"build": 233.1, "spsolve": 25.33, "source": "BASELINE.md 2.1 (survey host)"})

    # ---- C3
    nC3 = args.c3_size
    scale = nC3 / 16384.0
    M, P, S, V = int(3968 * scale), int(2048 * scale), int(2048 * scale), int(64 * scale)
    M += nC3 - (M + 2 * S + 4 * P + 2 * V)
    rows = gen.random_opamp_network_rows(M=M, P=P, S=S, V=V, seed=0)
    t0 = time.perf_counter()
    net = n.Netlist(write_csv(rows))
    parse_s = time.perf_counter() - t0
    table = net.table()
    assert table.n == nC3, table.n
    t0 = time.perf_counter()
    circ = n.Circuit(net)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    G0 = circ.G

    def c3():
        work = G0.clone()
        x, info = dev.lu_solve(work, circ.A)
        return x, info
    ms, (x, info) = timed(c3, warmup=1, reps=2)
    flops = 2.0 / 3.0 * nC3 ** 3 + 2.0 * nC3 ** 2
    line = dict(config="C3", what=f"dense LU solve, {nC3} unknowns (op-amps, E, VCVS)", ms=ms,
                tflops=flops / ms / 1e9, status=info["status"], host_parse_s=parse_s, build_s=build_s)
    if not args.no_cpu:
        Gh, Ah = G0.cpu().numpy(), circ.A.cpu().numpy()
        t0 = time.perf_counter()
        want = np.linalg.solve(Gh, Ah)
        line["cpu_numpy_solve_s"] = time.perf_counter() - t0
        line["cpu_threads"] = os.cpu_count()
        xh = x.cpu().numpy()
        kcl = table.kcl
        line["normwise_err_potentials"] = float(np.max(np.abs(xh[:kcl] - want[:kcl])) / np.max(np.abs(want[:kcl])))
        line["normwise_err_currents"] = float(np.max(np.abs(xh[kcl:] - want[kcl:])) / np.max(np.abs(want[kcl:])))
        line["relres"] = float(np.linalg.norm(Gh @ xh - Ah) / np.linalg.norm(Ah))
        line["relres_numpy"] = float(np.linalg.norm(Gh @ want - Ah) / np.linalg.norm(Ah))
    emit(**line)
    del G0, circ

    # ---- C4
    net = n.Netlist(write_csv(gen.OPAMP_AMPLIFIER_ROWS))
    table = net.table()
    batch = args.c4_batch
    vals = gen.opamp_sweep_values(batch, seed=0)
    dvals = dev.to_device(vals)
    ms, (x, info) = timed(lambda: dev.lu_batched(table, dvals), warmup=2, reps=5)
    line = dict(config="C4", what=f"batched LU, {batch} copies of opmodel_amplifier (6 unknowns)", ms=ms,
                systems_per_s=batch / ms * 1e3, gbs_algorithmic=96.0 * batch / ms / 1e6,
                singular=int((info != 0).sum().item()))
    if not args.no_cpu:
        from oracle import mna_oracle as orc
        sub = 2000
        t0 = time.perf_counter()
        for s in range(sub):
            v1, r1, ri, ro, gain, rf = (float(v) for v in vals[s])
            rows = [["v1", "E", repr(v1), "3", "g"], ["r1", "R", repr(r1), "g", "1"],
                    ["q1_ri", "R", repr(ri), "3", "1"], ["q1_ro", "R", repr(ro), "q1_internal_node", "2"],
                    ["q1_vcvs", "VCVS", repr(gain), "q1_internal_node", "g", "3", "1"],
                    ["q1_rf", "R", repr(rf), "1", "2"]]
            want = orc.solve_rows(rows)[4]
        dt = time.perf_counter() - t0
        line["cpu_oracle_systems_per_s"] = sub / dt
        line["cpu_sample"] = f"{sub} copies through the oracle port (numbering + stamping + numpy solve)"
        xs = x[sub - 1].cpu().numpy()
        line["last_sample_normwise_err"] = float(np.max(np.abs(xs[:4] - want[:4])) / np.max(np.abs(want[:4])))
    emit(**line)

    # ---- C5b
    if args.c5b:
        tn = gen.lattice3d(256)
        probe = copy.deepcopy(tn)
        probe.process_component(["a1", "A", "1", "1", "g"])
        table = probe.table()
        dtab = dev.upload_table(table)

        def c5b():
            csr, rhs = dev.assemble_csr(table, dtab=dtab)
            x, info = dev.pcg(csr, rhs, rtol=1e-10)
            return float(x[probe.nodenum["1"]]), info, csr.nnz
        ms, (r, info, nnz) = timed(c5b, warmup=1, reps=1)
        emit(config="C5b", what="256^3 lattice: stamp + CSR build + PCG", ms=ms, R=r, iterations=info["iterations"],
             relres=info["relres"], pcg_ms=info["solve_ms"], unknowns=table.n, nnz=nnz,
             pcg_gbs=(12 * nnz + 108 * table.n) * info["iterations"] / info["solve_ms"] / 1e6)


if __name__ == "__main__":
    main()
