"""GPU parity: stamping + CSR build (bit-exact against the reference's matrices)."""
import ctypes as C

import numpy as np
import pytest

import nodal_b200 as n
from helpers import golden, reduce_triples, stamp_on_host, write_csv
from nodal_b200 import _lib
from nodal_b200 import generators as gen
from nodal_b200.device import coo_stride, colbits_for
from oracle import mna_oracle as orc

pytestmark = pytest.mark.gpu
DOC = golden("doc_netlists.json")
GRIDS = golden("grids.json")


def csr_build_raw(device, keys_np, vals_np, n, cb):
    """Calls nodal_csr_build / nodal_csr_fetch on caller-made keyed triples."""
    torch = device.torch
    keys = device.to_device(keys_np.astype(np.int64))
    vals = device.to_device(vals_np)
    rhs = device.empty(max(1, n), torch.float64)[:n]
    nnz = C.c_int64(0)
    p = device.ptr
    _lib.check(device.lib.nodal_csr_build(device.ctx, n, len(keys_np), cb, p(keys), p(vals), p(rhs),
                                          C.byref(nnz), device.stream()), "nodal_csr_build")
    nnz = nnz.value
    indptr = device.empty(n + 1, torch.int32)
    indices = device.empty(max(1, nnz), torch.int32)[:nnz]
    data = device.empty(max(1, nnz), torch.float64)[:nnz]
    _lib.check(device.lib.nodal_csr_fetch(device.ctx, n, nnz, p(indptr), p(indices), p(data),
                                          device.stream()), "nodal_csr_fetch")
    torch.cuda.synchronize()
    return indptr.cpu().numpy(), indices.cpu().numpy(), data.cpu().numpy(), rhs.cpu().numpy()


@pytest.mark.parametrize("name", sorted(DOC))
def test_doc_netlists_bit_exact(device, name, tmp_path):
    g = DOC[name]
    if "G" not in g:
        pytest.skip("reference could not build this netlist")
    net = n.Netlist(write_csv(g["rows"], tmp_path / name))
    cs = n.Circuit(net, sparse=True)
    want = g["csr_sorted"]
    assert cs.G.indptr.cpu().numpy().tolist() == want["indptr"]
    assert cs.G.indices.cpu().numpy().tolist() == want["indices"]
    assert cs.G.data.cpu().numpy().tolist() == want["data"]
    assert cs.G.indices.dtype == device.torch.int32
    assert cs.A_host.tolist() == g["A"]
    assert cs.currents == g["currents"]
    # the column order G.tocsr() has in the reference BEFORE the solve sorts it (first touch, with DOK's
    # delete-on-zero / re-insert rule), bit for bit; solving then leaves the sorted form, as spsolve does
    ft = n.Circuit(net, sparse=True, csr_order="first_touch")
    wf = g["csr_first_touch"]
    assert ft.G.indptr.cpu().numpy().tolist() == wf["indptr"]
    assert ft.G.indices.cpu().numpy().tolist() == wf["indices"]
    assert ft.G.data.cpu().numpy().tolist() == wf["data"]
    if "result_sparse" in g and all(v is not None for v in g["result_sparse"]):
        ft.solve()
        assert ft.G.indices.cpu().numpy().tolist() == want["indices"]
    cd = n.Circuit(net, sparse=False)
    assert np.array_equal(cd.G_host, np.array(g["G"]))
    assert cd.A_host.tolist() == g["A"]
    ca = n.Circuit(net, sparse=False, atomic_stamp=True)
    assert np.allclose(ca.G_host, np.array(g["G"]), rtol=1e-15, atol=0)
    assert np.allclose(ca.A_host, np.array(g["A"]), rtol=1e-15, atol=0)


@pytest.mark.parametrize("key", ["grid2d_6", "grid2d_20", "lattice3d_5", "lattice3d_6"])
def test_grids_bit_exact(device, key):
    g = GRIDS[key]
    tn = gen.grid2d(g["N"]) if key.startswith("grid2d") else gen.lattice3d(g["N"])
    csr, rhs = device.assemble_csr(tn.table())
    want = g["csr_sorted"]
    assert csr.nnz == g["nnz"]
    assert csr.indptr.cpu().numpy().tolist() == want["indptr"]
    assert csr.indices.cpu().numpy().tolist() == want["indices"]
    assert csr.data.cpu().numpy().tolist() == want["data"]
    if "csr_first_touch" in g:
        cf, _ = device.assemble_csr(tn.table(), order="first_touch")
        wf = g["csr_first_touch"]
        assert cf.indptr.cpu().numpy().tolist() == wf["indptr"]
        assert cf.indices.cpu().numpy().tolist() == wf["indices"]
        assert cf.data.cpu().numpy().tolist() == wf["data"]
    assert not rhs.cpu().numpy().any()


@pytest.mark.parametrize("N", [50, 100])
def test_grid_checksums(device, N):
    g = GRIDS[f"grid2d_{N}"]
    csr, _ = device.assemble_csr(gen.grid2d(N).table())
    ip, ix, dt = (t.cpu().numpy() for t in (csr.indptr, csr.indices, csr.data))
    assert csr.nnz == g["nnz"]
    assert int(ip.astype(np.int64).sum()) == g["indptr_sum"]
    assert int(ix.astype(np.int64).sum()) == g["indices_sum"]
    assert int((ix.astype(np.int64) * (np.arange(csr.nnz) % 1009)).sum()) == g["indices_wsum"]
    assert float(dt.sum()) == g["data_sum"]


def test_grid_400_against_numpy_model(device):
    """Mid-size structural parity: CUDA stamp + sort + reduce == numpy model of the same
    pipeline == vectorised oracle assembly (indices bit-exact, data exact for a 1-ohm grid)."""
    t = gen.grid2d(400).table()
    csr, _ = device.assemble_csr(t)
    F = orc.assemble_resistive_fast(t.a, t.b, t.value, t.n)
    assert np.array_equal(csr.indptr.cpu().numpy(), F.indptr)
    assert np.array_equal(csr.indices.cpu().numpy(), F.indices)
    assert np.array_equal(csr.data.cpu().numpy(), F.data)
    assert csr.nnz == 400 * 400 + 2 * (2 * 400 * 399) - 9


@pytest.mark.parametrize("nslots,n,seed", [(1, 5, 0), (4095, 300, 1), (4097, 70000, 2),
                                           (1 << 20, 50000, 3), (3_000_001, 1 << 21, 4)])
def test_random_triples_sort_and_reduce(device, nslots, n, seed):
    """Sort stability + in-order sums on adversarial input: many duplicates, random order,
    invalid slots, rhs entries, values that cancel to exact zero."""
    rng = np.random.default_rng(seed)
    cb = colbits_for(n)
    rows = rng.integers(0, n + 1, nslots)              # row == n -> invalid slot
    hot = rng.integers(0, max(1, n // 50), nslots)     # concentrate duplicates
    rows = np.where(rng.random(nslots) < 0.5, np.minimum(rows, hot), rows)
    cols = rng.integers(0, n + 1, nslots)              # col == n -> rhs entry
    cols = np.where(rng.random(nslots) < 0.5, rows % (n + 1), cols)
    vals = rng.standard_normal(nslots)
    vals[rng.random(nslots) < 0.2] = 0.5
    k = nslots // 3
    if k:
        rows[-k:], cols[-k:], vals[-k:] = rows[:k], cols[:k], -vals[:k]   # pairs cancelling... in sum order
    keys = (rows.astype(np.int64) << cb) | cols.astype(np.int64)
    ip, ix, dt, rhs = csr_build_raw(device, keys, vals, n, cb)
    wip, wix, wdt, wrhs = reduce_triples(rows.astype(np.int32), cols.astype(np.int32), vals, n)
    assert np.array_equal(ip, wip)
    assert np.array_equal(ix, wix)
    assert np.array_equal(dt, wdt)                      # bit-exact: same summation order
    assert np.array_equal(rhs, wrhs)


def test_large_sort_three_level_scan(device):
    """> 4096^2 slots: exercises the recursive scan and multi-tile digit buckets."""
    nslots, n = 17_000_000, 1 << 22
    rng = np.random.default_rng(7)
    cb = colbits_for(n)
    rows = rng.integers(0, n, nslots).astype(np.int64)
    cols = (rows + rng.integers(-2, 3, nslots)) % n
    vals = rng.integers(1, 5, nslots).astype(np.float64)
    keys = (rows << cb) | cols
    ip, ix, dt, rhs = csr_build_raw(device, keys, vals, n, cb)
    order = np.argsort(keys, kind="stable")
    ks, vs = keys[order], vals[order]
    head = np.concatenate([[True], ks[1:] != ks[:-1]])
    starts = np.flatnonzero(head)
    sums = np.add.reduceat(vs, starts)                  # small integers: order-independent
    assert len(ix) == len(starts)
    assert np.array_equal(ix, (ks[starts] & ((1 << cb) - 1)).astype(np.int32))
    assert np.array_equal(dt, sums)
    assert np.array_equal(ip, np.searchsorted(ks[starts] >> cb, np.arange(n + 1)).astype(np.int32))
    assert not rhs.any()


def test_stamp_kernel_equals_host_core(device, tmp_path):
    """The CUDA stamp kernel and the CPU compilation of the same core emit identical triples."""
    rows = gen.random_opamp_network_rows(M=300, P=40, S=30, V=10, seed=3)
    net = n.Netlist(write_csv(rows, tmp_path / "c3.csv"))
    t = net.table()
    stride = coo_stride(t)
    dtab = device.upload_table(t)
    keys, vals, cb = device.stamp_coo(dtab, len(t), t.kcl, t.n, stride)
    device.torch.cuda.synchronize()
    keys, vals = keys.cpu().numpy(), vals.cpu().numpy()
    r, c_, v = stamp_on_host(t, stride)
    assert np.array_equal(keys >> cb, r)
    assert np.array_equal(keys & ((1 << cb) - 1), c_)
    assert np.array_equal(vals, v)


def test_empty_and_tiny(device, tmp_path):
    net = n.Netlist(write_csv([["r1", "R", "2", "1", "g"], ["a1", "A", "3", "1", "g"]], tmp_path / "t.csv"))
    cs = n.Circuit(net, sparse=True)
    assert cs.G.indptr.cpu().numpy().tolist() == [0, 1]
    assert cs.G.data.cpu().numpy().tolist() == [0.5]
    assert cs.A_host.tolist() == [3.0]


def test_row_pointer_with_long_empty_runs(device):
    """Entries confined to a few row ranges of a much larger matrix (what a rank of the
    multi-GPU path builds): long runs of empty rows go through the parallel gap fill."""
    rng = np.random.default_rng(11)
    n = 3_000_000
    cb = colbits_for(n)
    rows = np.concatenate([rng.integers(400_000, 400_500, 5000), rng.integers(1_000_000, 1_300_000, 200_000),
                           np.array([2_999_999, 17])])
    cols = rng.integers(0, n, len(rows))
    vals = rng.standard_normal(len(rows))
    keys = (rows.astype(np.int64) << cb) | cols.astype(np.int64)
    ip, ix, dt, rhs = csr_build_raw(device, keys, vals, n, cb)
    wip, wix, wdt, wrhs = reduce_triples(rows.astype(np.int32), cols.astype(np.int32), vals, n)
    assert np.array_equal(ip, wip) and np.array_equal(ix, wix) and np.array_equal(dt, wdt)
