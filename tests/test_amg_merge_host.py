"""CPU check of the merge-based Galerkin product (nodal_b200/csrc/amg_merge_core.cuh, what csrc/amg.cu
and csrc/dist_amg.cu build their coarse operators with): bit-identical to the numpy
statement amg_mirror.galerkin, i.e. to what the shipped sort + in-order segmented sum produces."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sps

import amg_mirror as mirror
from test_amg_host import grid_matrix, random_network

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        out = os.path.join(HERE, "host_check", "_build")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libamg_merge_host.so")
        srcs = [os.path.join(HERE, "host_check", "amg_merge_host.cpp"),
                os.path.join(HERE, "..", "nodal_b200", "csrc", "amg_merge_core.cuh")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", srcs[0], "-o", so])
        _lib = C.CDLL(so)
        _lib.amg_galerkin_merge_host.restype = C.c_int64
        _lib.amg_galerkin_merge_host.argtypes = [C.c_int32] + [C.c_void_p] * 9
    return _lib


def galerkin_by_merge(A, agg, nc):
    A = sps.csr_matrix(A)
    A.sort_indices()
    order = np.argsort(agg, kind="stable")                   # members of every aggregate, rows increasing
    pt_idx = order.astype(np.int32)
    pt_ptr = np.zeros(nc + 1, np.int32)
    np.cumsum(np.bincount(agg, minlength=nc), out=pt_ptr[1:])
    indptr, indices = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    data, agg32 = np.ascontiguousarray(A.data, np.float64), agg.astype(np.int32)
    out_ptr = np.zeros(nc + 1, np.int32)
    out_idx = np.zeros(max(1, A.nnz), np.int32)
    out_val = np.zeros(max(1, A.nnz), np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    nnz = lib().amg_galerkin_merge_host(nc, p(pt_ptr), p(pt_idx), p(indptr), p(indices), p(data), p(agg32),
                                        p(out_ptr), p(out_idx), p(out_val))
    return sps.csr_matrix((out_val[:nnz], out_idx[:nnz], out_ptr), shape=(nc, nc))


@pytest.mark.parametrize("case", ["grid30", "grid64", "random", "random_wide", "cancel"])
def test_merge_galerkin_is_bit_identical_to_the_sorted_one(case):
    if case.startswith("grid"):
        A, _ = grid_matrix(int(case[4:]))
    elif case == "random":
        A = random_network(4000, 12000, 1)
    elif case == "random_wide":
        A = random_network(3000, 5000, 2, decades=6.0)
    else:       # entries that cancel exactly inside an aggregate pair must disappear (DOK semantics)
        A = sps.csr_matrix(np.array([[2.0, -1.0, 1.0, 0.0], [-1.0, 2.0, -1.0, 0.0],
                                     [1.0, -1.0, 2.0, -1.0], [0.0, 0.0, -1.0, 2.0]]))
    for _ in range(2):                                         # both passes of a level
        agg, nc = mirror.aggregates(A)
        if case == "cancel":
            agg, nc = np.array([0, 0, 1, 1]), 2
        want = mirror.galerkin(A, agg, nc)
        want.sort_indices()
        got = galerkin_by_merge(A, agg, nc)
        assert np.array_equal(got.indptr, want.indptr)
        assert np.array_equal(got.indices, want.indices)
        assert np.array_equal(got.data.view(np.uint64), want.data.view(np.uint64))
        if case == "cancel":
            assert got.nnz < 4 or (got.toarray() != 0).all()
            break
        A = want
