"""Device-side plumbing: torch tensors as buffers, ctypes calls into libnodal_b200.so.

torch is used only to own device memory, streams and (in dist.py) the process
group; every kernel that runs here is one of ours.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from . import constants as K
from .table import ComponentTable

_STRIDE_OF_TYPE = {K.T_R: 4, K.T_A: 2, K.T_E: 5, K.T_VCVS: 6, K.T_VCCS: 6, K.T_CCVS: 6, K.T_CCCS: 5}


def _torch():
    import torch
    return torch


def coo_stride(table: ComponentTable) -> int:
    """Slots per component the stamp kernel must reserve (see include/nodal_b200.h)."""
    need = max((_STRIDE_OF_TYPE[t] for t in table.present_types()), default=2)
    return 2 if need <= 2 else 4 if need <= 4 else need


def colbits_for(n: int) -> int:
    return max(1, int(n).bit_length())


class DeviceCSR:
    """CSR matrix + right-hand side resident in HBM (int32 indices, f64 data)."""

    def __init__(self, n, indptr, indices, data):
        self.n = int(n)
        self.indptr, self.indices, self.data = indptr, indices, data

    @property
    def nnz(self):
        return int(self.data.numel())

    @property
    def shape(self):
        return (self.n, self.n)

    def tocsr(self):
        """Host scipy.sparse.csr_matrix copy (container only -- no arithmetic)."""
        import scipy.sparse as sps
        return sps.csr_matrix((self.data.cpu().numpy(), self.indices.cpu().numpy(),
                               self.indptr.cpu().numpy()), shape=self.shape)

    def toarray(self):
        return self.tocsr().toarray()

    def __array__(self, dtype=None, copy=None):
        a = self.toarray()
        return a if dtype is None else a.astype(dtype)


class DeviceAMG:
    """Handle of a nodal_amg hierarchy; keeps the matrix arrays alive while it exists."""

    PARAMS = ("passes", "coarse", "omega", "scale", "maxlevels", "rounds", "direct_max", "max_fill")

    def __init__(self, dev, csr, **params):
        unknown = set(params) - set(self.PARAMS)
        if unknown:
            raise TypeError(f"unknown AMG parameter(s): {sorted(unknown)}")
        self.dev, self.csr = dev, csr
        arr = (C.c_double * 8)(*[float(params.get(k, 0.0)) for k in self.PARAMS])
        h = C.c_void_p()
        p = dev.ptr
        st = dev.lib.nodal_amg_create(dev.ctx, csr.n, csr.nnz, p(csr.indptr), p(csr.indices), p(csr.data),
                                      arr, C.byref(h), dev.stream())
        _lib.check(st, "nodal_amg_create")
        self.handle = h
        cap = 64
        nl, rows, nnz = C.c_int32(0), (C.c_int64 * cap)(), (C.c_int64 * cap)()
        ms, direct = C.c_double(0.0), C.c_int32(0)
        _lib.check(dev.lib.nodal_amg_info(h, cap, C.byref(nl), rows, nnz, C.byref(ms), C.byref(direct)),
                   "nodal_amg_info")
        self.rows = [int(rows[k]) for k in range(nl.value)]
        self.nnz = [int(nnz[k]) for k in range(nl.value)]
        self.setup_ms, self.direct = ms.value, bool(direct.value)

    def close(self):
        if self.handle:
            self.dev.lib.nodal_amg_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown
            pass

    def aggregates(self, level):
        """Row -> aggregate map of `level` (device int32 tensor)."""
        dev = self.dev
        agg = dev.empty(max(1, self.rows[level]), dev.torch.int32)[: self.rows[level]]
        _lib.check(dev.lib.nodal_amg_fetch_level(dev.ctx, self.handle, level, dev.ptr(agg), None, None, None,
                                                 dev.stream()), "nodal_amg_fetch_level")
        return agg

    def operator(self, level):
        """The level's operator as a DeviceCSR copy."""
        dev, torch = self.dev, self.dev.torch
        n, nnz = self.rows[level], self.nnz[level]
        indptr = dev.empty(n + 1, torch.int32)
        indices = dev.empty(max(1, nnz), torch.int32)[:nnz]
        data = dev.empty(max(1, nnz), torch.float64)[:nnz]
        p = dev.ptr
        _lib.check(dev.lib.nodal_amg_fetch_level(dev.ctx, self.handle, level, None, p(indptr), p(indices),
                                                 p(data), dev.stream()), "nodal_amg_fetch_level")
        return DeviceCSR(n, indptr, indices, data)

    def apply(self, r):
        """z = M r: one V-cycle."""
        dev = self.dev
        z = dev.empty(max(1, self.csr.n), dev.torch.float64)[: self.csr.n]
        _lib.check(dev.lib.nodal_amg_apply(dev.ctx, self.handle, dev.ptr(r), dev.ptr(z), dev.stream()),
                   "nodal_amg_apply")
        return z

    def solve_multi(self, rhs, rtol=1e-10, maxit=None):
        """Several right-hand sides at once: rhs is a (K, n) device tensor, K <= 8; returns x (K, n) and
        a list of K info dicts (csrc/amg_multi.cu: every level operator is read once per sweep for all K)."""
        dev, n = self.dev, self.csr.n
        K = int(rhs.shape[0])
        if not 1 <= K <= 8 or int(rhs.shape[1]) != n:
            raise ValueError("rhs must be a (K, n) tensor with 1 <= K <= 8")
        rhs = rhs.contiguous()
        x = dev.zeros(max(2, K * n), dev.torch.float64)[: K * n].view(K, n)
        iters, relres, status = (C.c_int32 * K)(), (C.c_double * K)(), (C.c_int32 * K)()
        st = dev.lib.nodal_amg_pcg_multi(dev.ctx, self.handle, K, dev.ptr(rhs), dev.ptr(x), rtol, int(maxit or 1000),
                                         iters, relres, status, dev.stream())
        _lib.check(st, "nodal_amg_pcg_multi", allowed=(_lib.OK, _lib.NOT_CONVERGED, _lib.BREAKDOWN))
        infos = [dict(solver="amg_pcg_multi", status=int(status[k]), iterations=int(iters[k]), relres=float(relres[k]),
                      batch=K, level_rows=list(self.rows)) for k in range(K)]
        return x, infos

    def profile_sweeps(self, reps=64):
        """Average ms per launch of the level-0 SELL sweeps: dict(spmv_dot, residual, jacobi)."""
        ms = (C.c_double * 4)()
        _lib.check(self.dev.lib.nodal_amg_profile_sweeps(self.dev.ctx, self.handle, int(reps), ms, self.dev.stream()),
                   "nodal_amg_profile_sweeps")
        return dict(spmv_dot=ms[0], residual=ms[1], jacobi=ms[2])

    def solve(self, rhs, rtol=1e-10, maxit=None, x0=None):
        dev, n = self.dev, self.csr.n
        x = dev.zeros(max(2, n), dev.torch.float64)[:n] if x0 is None else x0.clone()
        if maxit is None:
            maxit = 1000
        iters, relres = C.c_int32(0), C.c_double(0.0)
        stats = (C.c_double * 16)()
        st = dev.lib.nodal_amg_pcg(dev.ctx, self.handle, dev.ptr(rhs), dev.ptr(x), rtol, maxit,
                                   C.byref(iters), C.byref(relres), stats, dev.stream())
        _lib.check(st, "nodal_amg_pcg", allowed=(_lib.OK, _lib.NOT_CONVERGED, _lib.BREAKDOWN))
        info = dict(solver="amg_pcg", status=st, iterations=iters.value, relres=relres.value,
                    restarts=int(stats[2]), solve_ms=stats[3], setup_ms=stats[4], levels=int(stats[0]),
                    operator_complexity=stats[1], grid_complexity=stats[6], coarsest_rows=int(stats[5]),
                    coarsest_direct=bool(stats[7]), level_rows=list(self.rows))
        return x, info


class Device:
    """One CUDA device + one library context.  Not thread safe (one per thread)."""

    _instances = {}
    _lock = threading.Lock()

    @classmethod
    def get(cls, index=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.NodalLibraryError(
                "nodal_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if index is None:
            index = torch.cuda.current_device()
        key = (threading.get_ident(), int(index))
        with cls._lock:
            if key not in cls._instances:
                cls._instances[key] = cls(int(index))
            return cls._instances[key]

    def __init__(self, index):
        self.torch = _torch()
        self.index = index
        self.dev = self.torch.device("cuda", index)
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.nodal_ctx_create(index, C.byref(h)), "nodal_ctx_create")
        self.ctx = h
        self.launches = 0   # kernels-of-ours launch counter is kept by the callers' estimates

    # ------------------------------------------------------------ helpers
    def stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def empty(self, n, dtype):
        return self.torch.empty(int(n), dtype=dtype, device=self.dev)

    def zeros(self, n, dtype):
        return self.torch.zeros(int(n), dtype=dtype, device=self.dev)

    def to_device(self, arr, pin=False):
        t = self.torch.from_numpy(np.ascontiguousarray(arr))
        if pin:
            t = t.pin_memory()
        return t.to(self.dev, non_blocking=pin)

    @staticmethod
    def ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def upload_table(self, table: ComponentTable, pin=False):
        """Columns of the table in HBM.  R / A-only tables (the SPD path) carry constants in the
        c / d / drv / branch columns: those are not uploaded (the kernels take null pointers for
        them), 17 instead of 33 bytes per component."""
        names = ("type", "value", "a", "b", "c", "d", "drv", "branch")
        if table.is_spd_structured():
            names = names[:4]
        if getattr(table, "_pinned", None):
            out = {name: table._pinned[name].to(self.dev, non_blocking=True) for name in names}
        else:
            out = {name: self.to_device(getattr(table, name), pin=pin) for name in names}
        for name in ("c", "d", "drv", "branch"):
            out.setdefault(name, None)
        return out

    @staticmethod
    def uploaded_bytes(table: ComponentTable):
        """Bytes upload_table moves for this table."""
        per_row = 17 if table.is_spd_structured() else 33
        return per_row * len(table)

    # ------------------------------------------------------------ stamping
    def stamp_coo(self, dtab, ncomp, kcl, n, stride):
        torch = self.torch
        cb = colbits_for(n)
        keys = self.empty(max(1, stride * ncomp), torch.int64)
        vals = self.empty(max(1, stride * ncomp), torch.float64)
        p = self.ptr
        _lib.check(self.lib.nodal_stamp_coo(
            self.ctx, ncomp, p(dtab["type"]), p(dtab["value"]), p(dtab["a"]), p(dtab["b"]),
            p(dtab["c"]), p(dtab["d"]), p(dtab["drv"]), p(dtab["branch"]),
            kcl, n, stride, cb, p(keys), p(vals), self.stream()), "nodal_stamp_coo")
        return keys, vals, cb

    def assemble_csr(self, table: ComponentTable, dtab=None, order="sorted"):
        """Stamp + CSR build.  Returns (DeviceCSR, rhs tensor).  order="first_touch" keeps the columns
        of a row in the order the reference's DOK matrix met them (what `G.tocsr()` gives before
        spsolve sorts it, nodal/nodal.py:396-397); the solvers expect the default, sorted order."""
        if dtab is None:
            dtab = self.upload_table(table)
        return self.assemble_csr_raw(dtab, len(table), table.kcl, table.n, coo_stride(table), order=order)

    def assemble_csr_raw(self, dtab, ncomp, kcl, n, stride, order="sorted"):
        """The same from device-resident columns (`dtab`: name -> tensor, `ncomp` rows used)."""
        torch = self.torch
        if order not in ("sorted", "first_touch"):
            raise ValueError("order must be 'sorted' or 'first_touch'")
        keys, vals, cb = self.stamp_coo(dtab, ncomp, kcl, n, stride)
        rhs = self.empty(max(1, n), torch.float64)[:n]
        nnz = C.c_int64(0)
        p = self.ptr
        _lib.check(self.lib.nodal_csr_build_ordered(self.ctx, n, stride * ncomp, cb, p(keys), p(vals),
                                                    1 if order == "first_touch" else 0, p(rhs), C.byref(nnz),
                                                    self.stream()), "nodal_csr_build")
        nnz = nnz.value
        indptr = self.empty(n + 1, torch.int32)
        indices = self.empty(max(1, nnz), torch.int32)[:nnz]
        data = self.empty(max(1, nnz), torch.float64)[:nnz]
        _lib.check(self.lib.nodal_csr_fetch(self.ctx, n, nnz, p(indptr), p(indices), p(data),
                                            self.stream()), "nodal_csr_fetch")
        # keys / vals may back the pending result until fetch has run on the stream
        self.torch.cuda.current_stream(self.dev).synchronize()
        del keys, vals
        return DeviceCSR(n, indptr, indices, data), rhs

    def select_local(self, dtab, ncomp, rb, re):
        """Columns of the components with a lead on a row in [rb, re), order kept, selected on the
        device from the resident table (row-partitioned assembly).  Returns (columns, count)."""
        torch = self.torch
        pos = self.empty(max(1, ncomp), torch.int32)
        count = C.c_int64(0)
        p = self.ptr
        _lib.check(self.lib.nodal_table_select_scan(self.ctx, ncomp, p(dtab["a"]), p(dtab["b"]), int(rb), int(re),
                                                    p(pos), C.byref(count), self.stream()), "nodal_table_select_scan")
        m = count.value
        names = ("type", "value", "a", "b", "c", "d", "drv", "branch")
        out = {k: (self.empty(max(1, m), dtab[k].dtype) if dtab[k] is not None else None) for k in names}
        _lib.check(self.lib.nodal_table_select_gather(
            self.ctx, ncomp, p(pos), int(rb), int(re), *[p(dtab[k]) for k in names], *[p(out[k]) for k in names],
            self.stream()), "nodal_table_select_gather")
        return out, m

    def csr_to_dense(self, csr: DeviceCSR):
        n = csr.n
        G = self.empty(max(1, n * n), self.torch.float64)[: n * n].view(n, n)
        p = self.ptr
        _lib.check(self.lib.nodal_csr_to_dense(self.ctx, n, p(csr.indptr), p(csr.indices),
                                               p(csr.data), p(G), self.stream()), "nodal_csr_to_dense")
        return G

    def assemble_dense(self, table: ComponentTable, atomic=False):
        """Dense n x n G and rhs.  atomic=False: deterministic (sorted, in-order sums, bit-exact
        with the reference); atomic=True: warp-aggregated atomic scatter-add."""
        if not atomic:
            csr, rhs = self.assemble_csr(table)
            return self.csr_to_dense(csr), rhs
        torch = self.torch
        n, ncomp = table.n, len(table)
        stride = coo_stride(table)
        dtab = self.upload_table(table)
        keys, vals, cb = self.stamp_coo(dtab, ncomp, table.kcl, n, stride)
        G = self.empty(max(1, n * n), torch.float64)[: n * n].view(n, n)
        rhs = self.empty(max(1, n), torch.float64)[:n]
        p = self.ptr
        _lib.check(self.lib.nodal_coo_to_dense_atomic(self.ctx, n, stride * ncomp, cb, p(keys),
                                                      p(vals), p(G), p(rhs), self.stream()),
                   "nodal_coo_to_dense_atomic")
        return G, rhs

    # ------------------------------------------------------------ sparse solves
    def spmv(self, csr: DeviceCSR, x):
        y = self.empty(max(1, csr.n), self.torch.float64)[: csr.n]
        p = self.ptr
        _lib.check(self.lib.nodal_spmv(self.ctx, csr.n, csr.nnz, p(csr.indptr), p(csr.indices),
                                       p(csr.data), p(x), p(y), self.stream()), "nodal_spmv")
        return y

    def pcg(self, csr: DeviceCSR, rhs, rtol=1e-10, maxit=None, flags=0, x0=None):
        n = csr.n
        x = self.zeros(max(2, n), self.torch.float64)[:n] if x0 is None else x0.clone()
        if maxit is None:
            maxit = max(5000, 40 * int(np.sqrt(max(n, 1))))
        iters, relres = C.c_int32(0), C.c_double(0.0)
        stats = (C.c_double * 16)()
        p = self.ptr
        st = self.lib.nodal_pcg(self.ctx, n, csr.nnz, p(csr.indptr), p(csr.indices), p(csr.data),
                                p(rhs), p(x), rtol, maxit, flags, C.byref(iters), C.byref(relres),
                                stats, self.stream())
        _lib.check(st, "nodal_pcg", allowed=(_lib.OK, _lib.NOT_CONVERGED, _lib.BREAKDOWN))
        info = dict(solver="pcg", status=st, iterations=iters.value, relres=relres.value,
                    restarts=int(stats[2]), solve_ms=stats[3], setup_ms=stats[4],
                    format="sell32" if stats[5] else "csr", stored_nnz=int(stats[6]),
                    spmv_grid=int(stats[7]), scaled=bool(stats[12]))
        if stats[11]:
            info["kernel_ms"] = dict(spmv_dot=stats[8], update=stats[9], direction=stats[10],
                                     samples=int(stats[11]))
        return x, info

    def connected_components(self, table: ComponentTable, want_labels=False):
        """(number of components, nodes connected to ground, labels or None) of the lead graph:
        nodes 0..kcl-1 plus ground (= node kcl), one edge per component (csrc/graph.cu)."""
        a, b = self.to_device(table.a), self.to_device(table.b)
        labels = self.empty(table.kcl + 1, self.torch.int32) if want_labels else None
        count, reached = C.c_int32(0), C.c_int32(0)
        st = self.lib.nodal_connected_components(self.ctx, len(table), self.ptr(a), self.ptr(b), table.kcl,
                                                 self.ptr(labels) if want_labels else None,
                                                 C.byref(count), C.byref(reached), self.stream())
        _lib.check(st, "nodal_connected_components")
        return count.value, reached.value, labels

    def amg(self, csr: DeviceCSR, **params):
        """Aggregation-AMG hierarchy for an SPD DeviceCSR (csrc/amg.cu); see DeviceAMG."""
        return DeviceAMG(self, csr, **params)

    def amg_pcg(self, csr: DeviceCSR, rhs, rtol=1e-10, maxit=None, x0=None, **params):
        """AMG-preconditioned CG, hierarchy built and dropped inside the call."""
        amg = DeviceAMG(self, csr, **params)
        try:
            return amg.solve(rhs, rtol=rtol, maxit=maxit, x0=x0)
        finally:
            amg.close()

    def gmres(self, csr: DeviceCSR, rhs, rtol=1e-12, restart=60, maxit=20000):
        n = csr.n
        x = self.zeros(max(2, n), self.torch.float64)[:n]
        iters, relres = C.c_int32(0), C.c_double(0.0)
        p = self.ptr
        st = self.lib.nodal_gmres(self.ctx, n, csr.nnz, p(csr.indptr), p(csr.indices), p(csr.data),
                                  p(rhs), p(x), rtol, restart, maxit, C.byref(iters),
                                  C.byref(relres), self.stream())
        _lib.check(st, "nodal_gmres", allowed=(_lib.OK, _lib.NOT_CONVERGED, _lib.BREAKDOWN))
        return x, dict(solver="gmres", status=st, iterations=iters.value, relres=relres.value)

    # ------------------------------------------------------------ dense solves
    def lu_solve(self, G, rhs):
        """Solves G x = rhs; G (n x n, row-major) is overwritten by its LU factors."""
        n = int(G.shape[0])
        x = self.empty(max(1, n), self.torch.float64)[:n]
        info = C.c_int32(0)
        p = self.ptr
        st = self.lib.nodal_lu_solve(self.ctx, n, p(G), p(rhs), p(x), C.byref(info), self.stream())
        _lib.check(st, "nodal_lu_solve", allowed=(_lib.OK, _lib.SINGULAR))
        return x, dict(solver="lu", status=st, info=info.value)

    def lu_batched(self, table: ComponentTable, values, layout="aos"):
        """Batched small-system solve.  layout "aos": values (batch, ncomp) -> x (batch, n);
        layout "soa": values (ncomp, batch) -> x (n, batch), the coalesced form (n <= 8).
        Returns x, info (batch,)."""
        torch = self.torch
        if layout not in ("aos", "soa"):
            raise ValueError("layout must be 'aos' or 'soa'")
        soa = layout == "soa"
        batch, ncomp = (int(values.shape[1]), int(values.shape[0])) if soa else (int(values.shape[0]), int(values.shape[1]))
        if ncomp != len(table):
            raise ValueError("values must have one entry per component of the topology")
        if not values.is_contiguous():
            values = values.contiguous()
        n = table.n
        dtab = self.upload_table(table)
        x = self.empty(max(1, batch * n), torch.float64)[: batch * n]
        x = x.view(n, batch) if soa else x.view(batch, n)
        info = self.empty(max(1, batch), torch.int32)[:batch]
        p = self.ptr
        fn = self.lib.nodal_lu_batched_soa if soa else self.lib.nodal_lu_batched
        _lib.check(fn(self.ctx, batch, ncomp, p(dtab["type"]), p(dtab["a"]), p(dtab["b"]), p(dtab["c"]),
                      p(dtab["d"]), p(dtab["drv"]), p(dtab["branch"]), table.kcl, n, p(values), p(x),
                      p(info), self.stream()), "nodal_lu_batched")
        return x, info
