"""Multi-GPU parity check (not a pytest file): run under torchrun on 2+ GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py [N]

Every rank assembles its rows, solves the N x N grid with the partitioned Jacobi-PCG and with the
partitioned AMG-PCG (replicated, distributed and single-pass hierarchies) and the result is
compared with the reference goldens and, for small N, with the single-GPU
solves of rank 0.  NODAL_DIST_NO_P2P=1 runs the same over NCCL instead of peer memory."""
import copy
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from nodal_b200 import dist as ndist
from nodal_b200 import generators as gen
from nodal_b200.device import Device

GOLD = {20: 0.7806032927398155, 50: 0.7743468244758764, 100: 0.7735139127312641, 200: 0.7733079842386412,
        400: 0.7732566450916762, 1000: 0.7732422803670024}


# level sizes of amg_mirror.AMG(grid_matrix(N)[0], partitions=2, gather_below=g) (computed on the CPU)
MIRROR_ROWS_2_RANKS = {(100, 300): [9999, 2057, 425], (400, 4000): [159999, 33087, 6891, 1443]}


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = Device.get(local)
    sizes = [int(a) for a in sys.argv[1:]] or [20, 100, 400]
    ok = True
    for N in sizes:
        net = copy.deepcopy(gen.grid2d(N))
        net.process_component(["a1", "A", "1", "1", "g"])
        table = net.table()
        single = None
        if rank == 0 and N <= 400:
            csr, rhs = dev.assemble_csr(table)
            x1, i1 = dev.pcg(csr, rhs, rtol=1e-10)
            x2, i2 = dev.amg_pcg(csr, rhs, rtol=1e-10)
            single = dict(R=float(x1[net.nodenum["1"]]), iterations=i1["iterations"],
                          R_amg=float(x2[net.nodenum["1"]]), iterations_amg=i2["iterations"])
        # Jacobi form, then the AMG form with everything replicated (gather_below above n), with
        # distributed levels down to ~n/40 rows, and with a single pairwise pass per level
        dtab = dev.upload_table(table)
        variants = [("jacobi", {}), ("amg", {}), ("amg", {"gather_below": max(300, N * N // 40)}),
                    ("amg", {"gather_below": max(300, N * N // 6), "passes": 1})]
        for precond, amg in variants:
            runner = ndist.GridRunner(dev, table, net.nodenum["1"], rank, world, rtol=1e-10, precond=precond, amg=amg)
            r, info = runner.step(dtab)
            msg = dict(N=N, world=world, precond=precond, amg=amg, R=r, iterations=info["iterations"],
                       relres=info["relres"], status=info["status"], halo_recv=info["halo_recv"],
                       solve_ms=info["solve_ms"], setup_ms=info["setup_ms"], comm=info["comm"])
            if precond == "amg":
                msg.update(levels=info["levels"], distributed_levels=info["distributed_levels"],
                           level_rows=info["level_rows"], replicated_rows=info["replicated_rows"])
            # the partitioned hierarchy is the numpy statement's (tests/amg_mirror.AMG(A, partitions=2,
            # gather_below=...).rows): rank-local aggregation reproduces its level sizes exactly
            want_rows = MIRROR_ROWS_2_RANKS.get((N, amg.get("gather_below"))) if world == 2 and amg.get("passes", 2) == 2 else None
            if want_rows is not None:
                msg["level_rows_match_numpy_statement"] = info["level_rows"][: len(want_rows)] == want_rows
                ok &= msg["level_rows_match_numpy_statement"]
            if N in GOLD:
                msg["rel_err_vs_reference"] = abs(r - GOLD[N]) / GOLD[N]
                ok &= msg["rel_err_vs_reference"] < 1e-9
            ok &= info["status"] == 0 and info["relres"] <= 1e-10
            if single is not None:
                msg["single_gpu"] = single
                ok &= abs(single["R"] - r) < 1e-9
                if precond == "amg" and amg.get("passes", 2) == 2 and single["iterations_amg"] >= 10:
                    # rank-local aggregation costs a few iterations at most (tests/study_partitioned_aggregation.py)
                    ok &= info["iterations"] <= single["iterations_amg"] + 8
            runner.pcg.close()
            if rank == 0:
                print(json.dumps(msg), flush=True)
    # ---- irregular halos: sparse random networks (first-appearance numbering), expander-like and banded
    for locality in (None, 200):
        tn = gen.random_network(20000, degree=8, seed=2, locality=locality)
        table = copy.deepcopy(tn)
        table.process_component(["a1", "A", "1", "1", "g"])
        table = table.table()
        row = tn.nodenum["1"]
        r_single = None
        if rank == 0:
            csr, rhs = dev.assemble_csr(table)
            x1, _ = dev.pcg(csr, rhs, rtol=1e-10)
            r_single = float(x1[row])
        dtab = dev.upload_table(table)
        for precond, amg in (("jacobi", {}), ("amg", {"gather_below": 2000})):
            runner = ndist.GridRunner(dev, table, row, rank, world, rtol=1e-10, precond=precond, amg=amg)
            r, info = runner.step(dtab)
            runner.pcg.close()
            good = info["status"] == 0 and info["relres"] <= 1e-10
            if rank == 0:
                good &= abs(r - r_single) <= 1e-9 * abs(r_single)
                print(json.dumps(dict(random_network=20000, locality=locality, precond=precond, R=r, single_gpu_R=r_single,
                                      iterations=info["iterations"], halo_recv=info["halo_recv"], status=info["status"],
                                      ok=bool(good))), flush=True)
            ok &= good

    # ---- the nodal surface under torchrun: Circuit(..., distributed=True) and equivalent_resistance
    import nodal_b200 as n
    import nodal_b200.equiv
    for N in [s for s in sizes if s <= 400][:2]:
        tn = gen.grid2d(N)
        for precond in ("auto", "jacobi"):
            r = n.equiv.equivalent_resistance(tn, "1", "g", sparse=True, distributed=True, precond=precond)
            stats = n.equiv.equivalent_resistance.last_stats
            good = stats["status"] == 0 and (N not in GOLD or abs(r - GOLD[N]) / GOLD[N] < 1e-9)
            ok &= good
            if rank == 0:
                print(json.dumps(dict(surface="equivalent_resistance(distributed=True)", N=N, precond=precond, R=r,
                                      solver=stats["solver"], iterations=stats["iterations"], ok=bool(good))), flush=True)
        probe = copy.deepcopy(tn)
        probe.process_component(["a1", "A", "1", "1", "g"])
        sol = n.Circuit(probe, sparse=True, distributed=True).solve()        # whole vector on every rank
        ref = n.Circuit(probe, sparse=True, precond="jacobi").solve()        # this rank's own single-GPU solve
        err = float(np.max(np.abs(sol.result - ref.result)) / np.max(np.abs(ref.result)))
        ok &= err < 1e-8 and len(sol.result) == probe.table().n
        if rank == 0:
            print(json.dumps(dict(surface="Circuit(distributed=True).solve()", N=N, solver=sol.stats["solver"],
                                  max_rel_diff_vs_single_gpu=err)), flush=True)
    # ---- the host path (share upload + device selection + all_to_all) builds the same rows as the
    # selection from the full resident table
    for N in sizes[:2]:
        net = copy.deepcopy(gen.grid2d(N))
        net.process_component(["a1", "A", "1", "1", "g"])
        table = net.table()
        runner = ndist.GridRunner(dev, table, net.nodenum["1"], rank, world, solver=ndist.shared_solver(dev, rank, world))
        a = runner.assemble(dev.upload_table(table))
        b = runner.assemble_from_host()
        same = all(torch.equal(u, v) for u, v in zip(a, b))
        ok &= same
        if rank == 0:
            print(json.dumps(dict(check="assemble_from_host == assemble(resident table)", N=N, bit_identical=bool(same))), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if flag.item() else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
