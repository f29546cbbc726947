"""GPU parity at the FULL size of every BASELINE.json config (C1 .. C5b) and of the multi-GPU
path, so that the driver-run `-m gpu` suite -- not a hand-run benchmark script -- is the
evidence for them.

  C1  doc/1.6.1.csv, dense                     -> values pinned by the reference's tests.py:54-61
  C2  1000 x 1000 grid, sparse                 -> reference end-to-end golden
  C3  16 384 unknowns, op-amps / E / VCVS, LU  -> LAPACK dgetrf/dgetrs (what numpy.linalg.solve,
                                                  nodal/nodal.py:327, calls) on the same matrix,
                                                  block-normwise, bounded by max(1e-9, cond * eps)
  C4  1 M copies of opmodel_amplifier          -> oracle on a random sample of 1 000 copies +
                                                  the closed-loop-gain property on all of them
  C5b 256^3 lattice (16.7 M unknowns)          -> size-independent properties (no CPU solver can
                                                  redo it): nnz formula, independently re-verified
                                                  residual, maximum principle
  multi-GPU: tests/dist_check.py under torchrun when the box has >= 2 GPUs.
"""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest

import nodal_b200 as n
import nodal_b200.equiv
from helpers import ROOT, block_err, golden, write_csv
from nodal_b200 import generators as gen
from oracle import mna_oracle as orc

pytestmark = pytest.mark.gpu
DOC = golden("doc_netlists.json")
PINNED = golden("pinned_by_reference_tests.json")


def test_config_c1_doc_161_dense(device, tmp_path):
    g = DOC["1.6.1.csv"]
    net = n.Netlist(write_csv(g["rows"], tmp_path / "1.6.1.csv"))
    sol = n.Circuit(net).solve()
    got = {l.split(" \t= ")[0]: float(l.split("= ")[1]) for l in str(sol).splitlines()[1:]}
    want = {"e(1)": 2.0, "e(2)": -1.0, "e(4)": 8.0, "i(d1)": -2.0, "i(e1)": 3.0}     # tests.py:54-61
    assert got.keys() == want.keys()
    for k in want:
        assert got[k] == pytest.approx(want[k], rel=1e-12)
    assert block_err(sol.result, g["result_dense"], len(net.nodenum)) < 1e-12


def test_config_c2_grid_1000_both_preconditioners(device):
    want = 0.7732422803670024                       # the reference's own end-to-end result
    for precond in ("jacobi", "amg"):
        r = n.equiv.equivalent_resistance(gen.grid2d(1000), "1", "g", sparse=True, precond=precond)
        stats = n.equiv.equivalent_resistance.last_stats
        assert stats["status"] == 0 and stats["relres"] <= 1e-10, (precond, stats)
        assert r == pytest.approx(want, rel=1e-9), precond


def test_config_c3_dense_16384(device, tmp_path, capsys):
    """Config C3 at its full size.  Reference arithmetic: LAPACK LU with partial pivoting on the
    same 16 384 x 16 384 matrix (scipy lu_factor/lu_solve = dgetrf + dgetrs = numpy.linalg.solve's
    dgesv).  Errors are reported per block (potentials / branch currents) and bounded by
    max(1e-9, cond_1 * eps) with the LAPACK 1-norm condition estimate (dgecon)."""
    import scipy.linalg as sla
    rows = gen.random_opamp_network_rows(seed=0)
    net = n.Netlist(write_csv(rows, tmp_path / "c3.csv"))
    circ = n.Circuit(net)
    assert circ.table.n == 16384
    kcl = circ.table.kcl
    sol = circ.solve()
    x = np.asarray(sol.result)
    assert sol.stats["status"] == 0
    G, A = circ.G_host, circ.A_host
    anorm = np.linalg.norm(G, 1)
    lu, piv = sla.lu_factor(G, check_finite=False)
    want = sla.lu_solve((lu, piv), A, check_finite=False)
    rcond, info = sla.lapack.dgecon(lu, anorm, norm="1")
    assert info == 0
    cond = 1.0 / rcond
    relres = np.linalg.norm(G @ x - A) / np.linalg.norm(A)
    relres_ref = np.linalg.norm(G @ want - A) / np.linalg.norm(A)
    scale_p, scale_c = np.max(np.abs(want[:kcl])), np.max(np.abs(want[kcl:]))
    err_p = np.max(np.abs(x[:kcl] - want[:kcl])) / scale_p
    err_c = np.max(np.abs(x[kcl:] - want[kcl:])) / scale_c
    bound = max(1e-9, cond * np.finfo(float).eps)
    with capsys.disabled():
        print(f"\n[C3 n=16384] cond_1 ~ {cond:.3e}  block err potentials {err_p:.3e} currents {err_c:.3e} "
              f"(bound {bound:.3e})  relres ours {relres:.3e} LAPACK {relres_ref:.3e}")
    assert relres <= 1e-10
    assert err_p <= bound and err_c <= bound


def test_config_c4_batched_one_million(device, tmp_path):
    net = n.Netlist(write_csv(gen.OPAMP_AMPLIFIER_ROWS, tmp_path / "c4.csv"))
    table = net.table()
    batch = 1_000_000
    vals = gen.opamp_sweep_values(batch, seed=0)
    x, info = device.lu_batched(table, device.to_device(vals))
    x, info = x.cpu().numpy(), info.cpu().numpy()
    assert x.shape == (batch, 6) and not info.any()
    sample = np.random.default_rng(4).choice(batch, size=1000, replace=False)
    worst = 0.0
    for s in sample:
        v1, r1, ri, ro, gain, rf = (float(v) for v in vals[s])
        rows = [["v1", "E", repr(v1), "3", "g"], ["r1", "R", repr(r1), "g", "1"],
                ["q1_ri", "R", repr(ri), "3", "1"], ["q1_ro", "R", repr(ro), "q1_internal_node", "2"],
                ["q1_vcvs", "VCVS", repr(gain), "q1_internal_node", "g", "3", "1"],
                ["q1_rf", "R", repr(rf), "1", "2"]]
        onet, G, A, _, want = orc.solve_rows(rows)
        assert onet.nodenum == net.nodenum and onet.anomnum == net.anomnum
        worst = max(worst, block_err(x[s], want, onet.kcl))
        assert np.linalg.norm(G @ x[s] - A) / np.linalg.norm(A) < 1e-10
    assert worst < 1e-9, worst
    # every copy: closed-loop gain of the non-inverting stage with finite open-loop gain,
    # e(2)/e(3) = k / (1 + k/gain), k = 1 + rf/r1 (ri = 1e7 and ro ~ 10 move it by < 1e-5)
    e2, e3 = x[:, net.nodenum["2"]], x[:, net.nodenum["3"]]
    ideal = 1 + vals[:, 5] / vals[:, 1]
    assert np.allclose(e2 / e3, ideal / (1 + ideal / vals[:, 4]), rtol=1e-4)
    assert np.allclose(e3, vals[:, 0], rtol=1e-12, atol=0)       # the E source pins node 3


def test_config_c5b_lattice_256_properties(device):
    """256^3 lattice, 16 777 215 unknowns, 117 047 283 non-zeros."""
    N = 256
    probe = copy.deepcopy(gen.lattice3d(N))
    probe.process_component(["a1", "A", "1", "1", "g"])
    table = probe.table()
    csr, rhs = device.assemble_csr(table)
    nodes, edges = N ** 3, 3 * N * N * (N - 1)
    assert csr.n == nodes - 1 and csr.nnz == nodes + 2 * edges - (1 + 2 * 6) == 117_047_283
    results = {}
    for precond in ("jacobi", "amg"):
        x, info = (device.pcg(csr, rhs, rtol=1e-10) if precond == "jacobi"
                   else device.amg_pcg(csr, rhs, rtol=1e-10))
        assert info["status"] == 0 and info["relres"] <= 1e-10, (precond, info)
        resid = device.spmv(csr, x) - rhs                    # independent kernel (generic CSR SpMV)
        assert float(resid.norm() / rhs.norm()) <= 1.05e-10
        r = float(x[probe.nodenum["1"]])
        assert float(x.max()) <= r * (1 + 1e-9) and float(x.min()) >= -1e-9      # maximum principle
        results[precond] = r
    assert results["amg"] == pytest.approx(results["jacobi"], rel=1e-9)
    # Rayleigh monotonicity: the lattice is a sub-network of the infinite cubic lattice, whose
    # resistance between nearest neighbours is 1/3 (a lower bound for any longer hop), and adding
    # resistors never raises a resistance: R(1 -> g) of the finite lattice >= infinite-lattice value
    assert 1.0 / 3.0 < results["jacobi"] < 1.0


def test_multi_gpu_partitioned_solve_matches_goldens():
    """Row-partitioned solvers on 2 GPUs (Jacobi and AMG forms) against the reference goldens and
    the single-GPU result: tests/dist_check.py under torchrun."""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "dist_check.py"), "20", "100", "400"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "DIST_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
