"""Struct-of-arrays component table: the host-side input of the stamp kernels.

One row per stamped component, in stamping order (reference: the loop over
``component_keys`` in nodal/nodal.py:357).  Columns:

    type   u8   T_R .. T_CCCS (constants.py)
    value  f64  resistance / current / voltage / gain
    a, b   i32  row index of the anode / bnode, GROUND (-1) if the lead is ground
    c, d   i32  row index of the control nodes, GROUND, or UNUSED (-2)
    drv    i32  table row of the driving resistor (CCVS/CCCS), else -1
    branch i32  branch number k (unknown index is kcl + k), else -1

33 bytes per component; this is the figure SURVEY.md section 8(d) uses for the
algorithmic bytes of the stamp pass.
"""
from __future__ import annotations

import numpy as np

from . import constants as K

_COLS = (("type", np.uint8), ("value", np.float64), ("a", np.int32), ("b", np.int32),
         ("c", np.int32), ("d", np.int32), ("drv", np.int32), ("branch", np.int32))

# COO triples a component can emit; fixed stride of the stamp kernel
MAX_TRIPLES = 6


class ComponentTable:
    __slots__ = tuple(n for n, _ in _COLS) + ("kcl", "be", "_pinned")

    def __init__(self, type, value, a, b, c=None, d=None, drv=None, branch=None, kcl=0, be=0):
        n = len(type)
        self.type = np.ascontiguousarray(type, dtype=np.uint8)
        self.value = np.ascontiguousarray(value, dtype=np.float64)
        self.a = np.ascontiguousarray(a, dtype=np.int32)
        self.b = np.ascontiguousarray(b, dtype=np.int32)
        fill = lambda arr, v: (np.full(n, v, dtype=np.int32) if arr is None  # noqa: E731
                               else np.ascontiguousarray(arr, dtype=np.int32))
        self.c = fill(c, K.UNUSED)
        self.d = fill(d, K.UNUSED)
        self.drv = fill(drv, -1)
        self.branch = fill(branch, -1)
        self.kcl = int(kcl)
        self.be = int(be)
        self._pinned = None
        for name, _ in _COLS:
            if len(getattr(self, name)) != n:
                raise ValueError(f"column {name} has wrong length")

    def __len__(self):
        return len(self.type)

    @property
    def n(self):
        """Number of unknowns (nodal/nodal.py:348)."""
        return self.kcl + self.be

    @property
    def nbytes(self):
        return sum(getattr(self, name).nbytes for name, _ in _COLS)

    def copy(self):
        return ComponentTable(*(getattr(self, name).copy() for name, _ in _COLS),
                              kcl=self.kcl, be=self.be)

    def append(self, type, value, a, b, c=None, d=None, drv=-1, branch=-1):
        """Return a new table with one more row (used by equivalent_resistance's probe source)."""
        row = dict(type=type, value=value, a=a, b=b,
                   c=K.UNUSED if c is None else c, d=K.UNUSED if d is None else d,
                   drv=drv, branch=branch)
        cols = [np.concatenate([getattr(self, name), np.array([row[name]], dtype=dt)])
                for name, dt in _COLS]
        return ComponentTable(*cols, kcl=self.kcl, be=self.be)

    def pin_memory(self):
        """Move the columns into page-locked host memory (torch) so uploads are async DMA."""
        import torch
        pinned = {}
        for name, _ in _COLS:
            t = torch.from_numpy(getattr(self, name)).pin_memory()
            pinned[name] = t
            setattr(self, name, t.numpy())
        self._pinned = pinned
        return self

    def is_resistive(self):
        return bool(np.all(self.type == K.T_R))

    def is_spd_structured(self):
        """Only R (positive) and A rows: G is a weighted graph Laplacian with the
        ground row removed -> symmetric positive definite if connected."""
        t = self.type
        ok = (t == K.T_R) | (t == K.T_A)
        if not bool(np.all(ok)):
            return False
        r = t == K.T_R
        return bool(np.all(self.value[r] > 0))

    def validate(self):
        """Host-side checks the reference performs while stamping."""
        t = self.type
        if np.any((t == K.T_R) & (self.value == 0)):
            raise ValueError("Model error: resistors can't have null resistance")
        cc = (t == K.T_CCVS) | (t == K.T_CCCS)
        if np.any(cc):
            drv = self.drv[cc]
            if np.any(drv < 0):
                raise KeyError("Driving component not found")
            if np.any(self.type[drv] != K.T_R):
                # the reference fails with AttributeError on non-R drivers (models.py:146,200)
                raise AttributeError("only resistors are supported as driving components")
            if np.any(self.value[drv] == 0):
                raise ZeroDivisionError("float division by zero")
        lim = self.kcl
        for name in ("a", "b"):
            v = getattr(self, name)
            if np.any(v >= lim) or np.any(v < K.GROUND):
                raise ValueError(f"lead index out of range in column {name}")
        if np.any(self.branch >= self.be):
            raise ValueError("branch index out of range")

    def coo_upper_bound(self):
        return MAX_TRIPLES * len(self)
