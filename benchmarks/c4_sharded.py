#!/usr/bin/env python
"""Config C4 across GPUs: the batch of small systems is split into contiguous slices, one per
rank, with no collective on the data path.  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29513 benchmarks/c4_sharded.py [--batch 1000000]

Rank 0 prints one JSON line: systems/s of the whole job (max over ranks of the device time)."""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import nodal_b200 as n  # noqa: E402
from nodal_b200 import generators as gen  # noqa: E402
from nodal_b200.device import Device  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = Device.get(local)
    import csv
    fd, path = tempfile.mkstemp(suffix=".csv")
    with os.fdopen(fd, "w", newline="") as fh:
        csv.writer(fh).writerows(gen.OPAMP_AMPLIFIER_ROWS)
    table = n.Netlist(path).table()
    vals = gen.opamp_sweep_values(args.batch, seed=0)
    lo, hi = rank * args.batch // world, (rank + 1) * args.batch // world      # contiguous slice of this rank
    dvals = dev.to_device(vals[lo:hi])
    for _ in range(3):
        x, info = dev.lu_batched(table, dvals)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        x, info = dev.lu_batched(table, dvals)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device="cuda", dtype=torch.float64)
    bad = torch.tensor([int((info != 0).sum().item())], device="cuda")
    gain = x[:, 3] / x[:, 0]                                        # e(2) / e(3): ideal 1 + rf / r1
    ok = torch.tensor([float(torch.allclose(gain.cpu(), torch.from_numpy(1 + vals[lo:hi, 5] / vals[lo:hi, 1]),
                                            rtol=2e-2))], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(bad)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"config": "C4 sharded", "n_gpus": world, "batch": args.batch, "ms": ms.item(),
                          "systems_per_s": args.batch / ms.item() * 1e3, "singular": int(bad.item()),
                          "gain_check": bool(ok.item()), "collectives_on_data_path": 0}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
