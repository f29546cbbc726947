"""Shared helpers of the test-suite (golden loading, netlist files, metrics)."""
import csv
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def write_csv(rows, path):
    with open(path, "w", newline="") as fh:
        csv.writer(fh).writerows(rows)
    return str(path)


def block_err(x, ref, kcl):
    """Block-normwise relative error: max|x-ref| / max|ref| taken
    separately over the potential block and the branch-current block."""
    x, ref = np.asarray(x, float), np.asarray(ref, float)
    errs = []
    for blk in (slice(0, kcl), slice(kcl, None)):
        if ref[blk].size:
            scale = np.max(np.abs(ref[blk]))
            errs.append(np.max(np.abs(x[blk] - ref[blk])) / (scale if scale > 0 else 1.0))
    return max(errs) if errs else 0.0


_stamp_host = None


def stamp_host_lib():
    """Builds (once) the CPU compilation of csrc/stamp_core.cuh used by the not-gpu tests."""
    global _stamp_host
    if _stamp_host is None:
        out = os.path.join(HERE, "host_check", "_build")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libstamp_host.so")
        src = os.path.join(HERE, "host_check", "stamp_host.cpp")
        core = os.path.join(ROOT, "nodal_b200", "csrc", "stamp_core.cuh")
        if (not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src),
                                                                  os.path.getmtime(core))):
            subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
        _stamp_host = C.CDLL(so)
    return _stamp_host


_amg_host = None


def amg_host_lib():
    """CPU compilation of csrc/amg_core.cuh (aggregation rules) for the not-gpu tests."""
    global _amg_host
    if _amg_host is None:
        out = os.path.join(HERE, "host_check", "_build")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libamg_host.so")
        src = os.path.join(HERE, "host_check", "amg_host.cpp")
        core = os.path.join(ROOT, "nodal_b200", "csrc", "amg_core.cuh")
        if (not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src),
                                                                  os.path.getmtime(core))):
            subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
        lib = C.CDLL(so)
        lib.amg_edge_hash_host.restype = C.c_uint32
        lib.amg_edge_hash_host.argtypes = [C.c_int32, C.c_int32]
        lib.amg_aggregate_host.restype = C.c_int32
        lib.amg_aggregate_host.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                           C.c_void_p, C.c_void_p]
        _amg_host = lib
    return _amg_host


def amg_aggregate_on_host(A, rounds=8):
    """(match, agg, nc) of one pairwise pass over scipy CSR `A`, computed by the shared device core."""
    A = A.tocsr()
    A.sort_indices()
    n = A.shape[0]
    indptr = np.ascontiguousarray(A.indptr, np.int32)
    indices = np.ascontiguousarray(A.indices, np.int32)
    data = np.ascontiguousarray(A.data, np.float64)
    match = np.empty(n, np.int32)
    agg = np.empty(n, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    nc = amg_host_lib().amg_aggregate_host(n, p(indptr), p(indices), p(data), rounds, p(match), p(agg))
    return match, agg, nc


def stamp_on_host(table, stride):
    """(rows, cols, vals) the stamp kernel would emit, computed by the shared core on the CPU."""
    lib = stamp_host_lib()
    m = len(table)
    rows = np.empty(m * stride, np.int32)
    cols = np.empty(m * stride, np.int32)
    vals = np.empty(m * stride, np.float64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib.stamp_table_host.argtypes = [C.c_int64] + [C.c_void_p] * 8 + [C.c_int32] * 3 + [C.c_void_p] * 3
    rc = lib.stamp_table_host(m, p(table.type), p(table.value), p(table.a), p(table.b), p(table.c),
                              p(table.d), p(table.drv), p(table.branch), table.kcl, table.n, stride,
                              p(rows), p(cols), p(vals))
    assert rc == 0
    return rows, cols, vals


def reduce_triples(rows, cols, vals, n):
    """numpy model of csrc/csr.cu: stable sort by (row, col), in-order sums, drop exact zeros,
    split col == n into the rhs.  Returns (indptr, indices, data, rhs)."""
    live = rows < n
    rows, cols, vals = rows[live].astype(np.int64), cols[live].astype(np.int64), vals[live]
    order = np.argsort(rows * (n + 1) + cols, kind="stable")
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(n + 1, np.int32)
    indices, data = [], []
    rhs = np.zeros(n)
    i = 0
    while i < len(rows):
        j, s = i + 1, vals[i]
        while j < len(rows) and rows[j] == rows[i] and cols[j] == cols[i]:
            s = s + vals[j]
            j += 1
        if cols[i] == n:
            rhs[rows[i]] = s
        elif s != 0.0:
            indices.append(cols[i]); data.append(s); indptr[rows[i] + 1] += 1
        i = j
    np.cumsum(indptr, out=indptr)
    return indptr, np.array(indices, np.int32), np.array(data, np.float64), rhs
