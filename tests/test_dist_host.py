"""Multi-rank host logic on CPU: gloo, world_size 2 (partitioning, per-rank component tables,
unique-id hand-off).  The NCCL / kernel side is covered by tests/dist_check.py on GPUs."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from helpers import reduce_triples, stamp_on_host
from nodal_b200 import dist as ndist
from nodal_b200 import generators as gen


def test_partition_rows():
    for n, w in ((10, 1), (10, 3), (16777215, 8), (7, 7)):
        b = ndist.partition_rows(n, w)
        assert b[0] == 0 and b[-1] == n and len(b) == w + 1
        sizes = np.diff(b)
        assert sizes.min() >= 1 and sizes.max() - sizes.min() <= 1


def test_local_tables_reproduce_global_rows():
    import copy
    tn = copy.deepcopy(gen.grid2d(24))
    tn.process_component(["a1", "A", "1", "1", "g"])
    table = tn.table()
    r, c, v = stamp_on_host(table, 4)
    ip, ix, dt, rhs = reduce_triples(r, c, v, table.n)
    for world in (2, 3, 5):
        bounds = ndist.partition_rows(table.n, world)
        covered = 0
        for k in range(world):
            rb, re = int(bounds[k]), int(bounds[k + 1])
            loc = ndist.local_component_table(table, rb, re)
            assert len(loc) < len(table)
            lr, lc, lv = stamp_on_host(loc, 4)
            lip, lix, ldt, lrhs = reduce_triples(lr, lc, lv, table.n)
            s, e = lip[rb], lip[re]
            assert np.array_equal(lip[rb:re + 1] - s, ip[rb:re + 1] - ip[rb])
            assert np.array_equal(lix[s:e], ix[ip[rb]:ip[re]])
            assert np.array_equal(ldt[s:e], dt[ip[rb]:ip[re]])          # bit-exact local rows
            assert np.array_equal(lrhs[rb:re], rhs[rb:re])
            covered += re - rb
        assert covered == table.n


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        payload = bytes(range(128)) if rank == 0 else None
        got = ndist.broadcast_bytes(payload, 128, src=0)
        table = gen.grid2d(12).table()
        bounds = ndist.partition_rows(table.n, world)
        loc = ndist.local_component_table(table, int(bounds[rank]), int(bounds[rank + 1]))
        import torch
        cnt = torch.tensor([len(loc)], dtype=torch.int64)
        dist.all_reduce(cnt)
        out.put((rank, got == bytes(range(128)), int(cnt.item()), len(table)))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    total, ncomp = res[0][2], res[0][3]
    assert ncomp < total <= ncomp + 2 * 12 * 2     # boundary components are stamped on both ranks


def test_share_upload_and_selection_preserve_stamping_order():
    """Host path of the row-partitioned assembly (GridRunner.assemble_from_host): rank s uploads the
    s-th contiguous share of the component table, selects per destination rank d the components
    touching d's rows, and d concatenates what it receives in source-rank order.  Statement in numpy
    of why that equals local_component_table(table, rows of d) -- same components, same (global
    stamping) order -- for every rank count, including shares that select nothing."""
    from nodal_b200 import dist as ndist
    from nodal_b200 import generators as gen
    for net in (gen.grid2d(17), gen.random_network(700, degree=6, seed=4), gen.random_network(700, degree=6, seed=5, locality=30)):
        table = net.table()
        m = len(table)
        for world in (1, 2, 3, 8):
            bounds = ndist.partition_rows(table.n, world)
            received = {d: [] for d in range(world)}
            for s in range(world):                                  # source ranks in rank order
                lo, hi = m * s // world, m * (s + 1) // world
                a, b = table.a[lo:hi], table.b[lo:hi]
                for d in range(world):
                    rb, re = int(bounds[d]), int(bounds[d + 1])
                    touch = ((a >= rb) & (a < re)) | ((b >= rb) & (b < re))
                    received[d].append(np.flatnonzero(touch) + lo)  # select keeps the share's order
            for d in range(world):
                got = np.concatenate(received[d])
                want = ndist.local_component_table(table, int(bounds[d]), int(bounds[d + 1]))
                assert len(got) == len(want)
                assert np.array_equal(table.a[got], want.a) and np.array_equal(table.b[got], want.b)
                assert np.array_equal(table.value[got], want.value)
                assert np.all(np.diff(got) > 0)                     # global stamping order
