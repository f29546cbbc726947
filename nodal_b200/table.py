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
    __slots__ = tuple(n for n, _ in _COLS) + ("kcl", "be", "_pinned", "_facts")

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
        self._facts = None
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

    # ---- cached column scans -------------------------------------------------------
    # Every question the host asks about the columns (which types occur, are all resistances
    # positive, are the indices in range) is answered by one scan that is kept with the table:
    # at 33.5 M components the separate numpy passes (np.unique, boolean masks, fancy indexing)
    # cost 0.7 s per Circuit, several times the GPU time of the whole solve.  The columns are
    # treated as frozen once scanned; call ``invalidate()`` after editing them in place.
    def invalidate(self):
        self._facts = None

    def facts(self):
        f = self._facts
        if f is not None:
            return f
        t, m = self.type, len(self)
        f = dict(present=0, unknown_type=False, r_zero=False, r_nonpositive=False)
        if m:
            f["unknown_type"] = bool(t.max() >= len(K.TYPE_NAME))
            if not f["unknown_type"]:
                f["present"] = int(np.bitwise_or.reduce(np.left_shift(np.uint8(1), t, dtype=np.uint8)))
            else:
                f["present"] = sum(1 << int(k) for k in np.unique(t) if k < 8)
            has = lambda code: bool(f["present"] >> code & 1)          # noqa: E731
            if has(K.T_R):
                positive = self.value > 0            # False for 0, negatives and NaN
                if f["present"] != 1 << K.T_R:
                    positive |= t != K.T_R
                if not bool(positive.all()):
                    f["r_nonpositive"] = True
                    zero = self.value == 0
                    zero &= t == K.T_R
                    f["r_zero"] = bool(zero.any())
        self._facts = f
        return f

    def has_type(self, code):
        return bool(self.facts()["present"] >> code & 1)

    def present_types(self):
        return [k for k in range(len(K.TYPE_NAME)) if self.has_type(k)]

    def is_resistive(self):
        return len(self) == 0 or self.facts()["present"] == 1 << K.T_R

    def is_spd_structured(self):
        """Only R (positive) and A rows: G is a weighted graph Laplacian with the
        ground row removed -> symmetric positive definite if connected."""
        f = self.facts()
        if f["unknown_type"] or f["present"] & ~((1 << K.T_R) | (1 << K.T_A)):
            return False
        return not f["r_nonpositive"]

    def validate(self):
        """Host-side checks the reference performs while stamping."""
        f = self.facts()
        if f.get("validated"):
            return
        if f["r_zero"]:
            raise ValueError("Model error: resistors can't have null resistance")
        if self.has_type(K.T_CCVS) or self.has_type(K.T_CCCS):
            t = self.type
            cc = (t == K.T_CCVS) | (t == K.T_CCCS)
            drv = self.drv[cc]
            if np.any(drv < 0):
                raise KeyError("Driving component not found")
            if np.any(self.type[drv] != K.T_R):
                # the reference fails with AttributeError on non-R drivers (models.py:146,200)
                raise AttributeError("only resistors are supported as driving components")
            if np.any(self.value[drv] == 0):
                raise ZeroDivisionError("float division by zero")
        if len(self):
            lim = self.kcl
            for name in ("a", "b"):
                v = getattr(self, name)
                if v.max() >= lim or v.min() < K.GROUND:
                    raise ValueError(f"lead index out of range in column {name}")
            if self.branch.max() >= self.be:
                raise ValueError("branch index out of range")
        f["validated"] = True

    def coo_upper_bound(self):
        return MAX_TRIPLES * len(self)
