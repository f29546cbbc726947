"""Stamp functions of the MNA procedure -- host half.

Same names and positional signatures as the reference's nodal/models.py
(write_R :13, write_A :27, write_E :35, write_VCVS :53, write_VCCS :81,
write_CCVS :109, write_CCCS :161).  In the reference these functions do the
arithmetic themselves, one numpy/scipy item access at a time.  Here they only
*record* the component into a struct-of-arrays table (``StampRecorder`` plays
the role of ``G``/``A``); the arithmetic -- conductances, '=' versus '+='
resolution, duplicate summation, zero removal -- happens in the CUDA stamp and
CSR-build kernels (csrc/stamp_core.cuh, csrc/stamp.cu, csrc/csr.cu), which is
where each reference line is cited.

They keep the reference's host-visible error behaviour: ValueError for a null
resistance, KeyError for a missing driver or unknown control node,
AssertionError when a current-controlled source's control nodes are not its
driver's leads, AttributeError for non-resistor drivers (the reference trips
over its own ``c.NODE_TYPES_ANOM`` there, models.py:146,200).
"""
from __future__ import annotations

import numpy as np

from . import constants as K
from .table import ComponentTable


class _RhsProxy:
    """Stands in for the vector ``A`` handed to write_A."""

    def __init__(self, recorder):
        self.recorder = recorder


class StampRecorder:
    """Collects component rows in stamping order; ``finish()`` -> ComponentTable."""

    def __init__(self, netlist):
        self.netlist = netlist
        self.currents = []
        self.rhs_proxy = _RhsProxy(self)
        self._rows = []          # (type, value, a, b, c, d, driver_name_or_None, branch)
        self._row_of_name = {}

    @staticmethod
    def _idx(i):
        return K.GROUND if i is None else int(i)

    def add(self, comp, type_code, a, b, c=K.UNUSED, d=K.UNUSED, driver=None, branch=-1):
        self._row_of_name.setdefault(comp.name, len(self._rows))
        self._rows.append((type_code, comp.value, self._idx(a), self._idx(b), c, d, driver, branch))

    def finish(self):
        rows = self._rows
        drv = [(-1 if r[6] is None else self._row_of_name.get(r[6], -1)) for r in rows]
        nums = self.netlist.nums
        col = lambda k, dt: np.array([r[k] for r in rows], dtype=dt)  # noqa: E731
        return ComponentTable(col(0, np.uint8), col(1, np.float64), col(2, np.int32),
                              col(3, np.int32), col(4, np.int32), col(5, np.int32),
                              np.array(drv, dtype=np.int32), col(7, np.int32),
                              kcl=nums["kcl"], be=nums["be"])


def _recorder(G):
    rec = G.recorder if isinstance(G, _RhsProxy) else G
    if not isinstance(rec, StampRecorder):
        raise TypeError("nodal_b200.models.write_* record into a StampRecorder; dense/dok "
                        "matrices are assembled on the GPU (see Circuit.build_model)")
    return rec


def _node(nodenum, label, ground):
    return K.GROUND if label == ground else nodenum[label]


def write_R(c, i, j, ground, G):
    if c.value == 0:
        raise ValueError("Model error: resistors can't have null resistance")
    _recorder(G).add(c, K.T_R, i, j)


def write_A(c, i, j, ground, A):
    _recorder(A).add(c, K.T_A, i, j)


def write_E(c, i, j, ground, G, A, currents, anomnum, nums, nodenum):
    currents.append(c.name)
    _recorder(G).add(c, K.T_E, i, j, branch=anomnum[c.name])


def write_VCVS(c, i, j, ground, G, A, currents, anomnum, nums, nodenum):
    currents.append(c.name)
    c.cnode, c.dnode = c.pos_control, c.neg_control
    code = K.T_VCCS if c.type == "VCCS" else K.T_VCVS   # same stamp; the code is kept for reporting
    _recorder(G).add(c, code, i, j, _node(nodenum, c.cnode, ground), _node(nodenum, c.dnode, ground),
                     branch=anomnum[c.name])


def write_VCCS(c, i, j, ground, G, currents, anomnum, nums, nodenum):
    """Kept for surface parity.  The reference never calls it (VCCS is dispatched to
    write_VCVS, nodal.py:377-378); recording through it gives the same VCVS-style stamp
    so results stay identical to the reference's."""
    currents.append(c.name)
    c.cnode, c.dnode = c.pos_control, c.neg_control
    _recorder(G).add(c, K.T_VCCS, i, j, _node(nodenum, c.cnode, ground),
                     _node(nodenum, c.dnode, ground), branch=anomnum[c.name])


def _driver_of(c, components):
    try:
        driver = components[c.driver]
    except KeyError:
        raise KeyError(f"Driving component {c.driver} not found")
    c.cnode, c.dnode = c.pos_control, c.neg_control
    assert c.cnode is not None and c.dnode is not None and driver is not None
    assert (c.cnode == driver.anode and c.dnode == driver.bnode) or (
        c.cnode == driver.bnode and c.dnode == driver.anode)
    if driver.type != "R":
        raise AttributeError(
            f"driver {driver.name} of {c.name} is a {driver.type}: only resistors can drive "
            "current-controlled sources (as in the reference)")
    if driver.value == 0:
        raise ZeroDivisionError("float division by zero")
    return driver


def write_CCVS(c, i, j, ground, G, A, currents, anomnum, nums, nodenum, components):
    currents.append(c.name)
    driver = _driver_of(c, components)
    _recorder(G).add(c, K.T_CCVS, i, j, _node(nodenum, c.cnode, ground),
                     _node(nodenum, c.dnode, ground), driver=driver.name, branch=anomnum[c.name])


def write_CCCS(c, i, j, ground, G, A, currents, anomnum, nums, nodenum, components):
    currents.append(c.name)
    driver = _driver_of(c, components)
    _recorder(G).add(c, K.T_CCCS, i, j, _node(nodenum, c.cnode, ground),
                     _node(nodenum, c.dnode, ground), driver=driver.name, branch=anomnum[c.name])
