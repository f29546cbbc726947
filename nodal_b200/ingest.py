"""Vectorised netlist ingest: csv -> component table with the reference's numbering.

The reference parses one row at a time into a dict of ``Component`` objects
(nodal/nodal.py:222-296, ~7 us per row) and ``equivalent_resistance`` then deep-copies
that dict (nodal/equiv.py:50, ~18 us per component): at 33.5 M rows that is tens of
minutes before any arithmetic.  ``read_table_netlist`` produces the same numbering --
component order, first-appearance node discovery (anode then bnode), ground choice,
branch numbers, OPMODEL expansion appended after all csv rows -- with pandas' C parser
and numpy, and returns a ``TableNetlist``.

Rows are validated with the same rules as ``Component.check_input`` (arity per type,
known type, numeric value); errors are raised as ValueError like the reference does.
"""
from __future__ import annotations

import csv
import io

import numpy as np

from . import constants as K
from .generators import TableNetlist
from .nodal import build_opmodel, find_ground_node
from .table import ComponentTable


def _read_rows(path):
    """csv.reader(skipinitialspace=True) semantics, comment / blank lines dropped, as a
    list of 8 equal-length python lists (missing trailing fields are None) + field counts."""
    with open(path, "r", newline="") as fh:
        text = fh.read()
    def python_path():
        rows = [r for r in csv.reader(io.StringIO(text), skipinitialspace=True) if r and r[0][:1] != "#"]
        counts = np.array([len(r) for r in rows], dtype=np.int64)
        cols = [np.array([r[k] if k < len(r) else None for r in rows], dtype=object) for k in range(9)]
        return cols, counts

    try:
        import pandas as pd
        from pandas.errors import ParserError
    except ImportError:  # pragma: no cover - pandas is part of the image
        return python_path()
    if '"' in text:          # quoted fields: let the csv module decide what a field is
        return python_path()
    lines = [ln for ln in text.splitlines() if ln.strip() and not ln.lstrip().startswith("#")]
    counts = np.char.count(np.array(lines, dtype=str), ",").astype(np.int64) + 1 if lines else np.zeros(0, np.int64)
    if len(counts) and counts.max() > 9:
        return python_path()
    try:
        frame = pd.read_csv(io.StringIO("\n".join(lines)), header=None, names=list(range(9)), dtype=str,
                            skipinitialspace=True, comment=None, keep_default_na=False, na_filter=False,
                            skip_blank_lines=False, engine="c")
    except ParserError:
        return python_path()
    if len(frame) != len(lines):
        return python_path()
    cols = [frame[k].to_numpy(dtype=object) for k in range(9)]
    return cols, counts


def read_table_netlist(path):
    """Parse `path` into a TableNetlist whose numbering is bit-identical to
    ``nodal.Netlist(path)`` (checked in tests/test_ingest.py)."""
    cols, counts = _read_rows(path)
    name, kind, value, a, b, c_, d_, drv = (cols[k] for k in range(8))
    m = len(name)
    if m and (counts > 8).any():
        bad = int(np.flatnonzero(counts > 8)[0])
        raise ValueError(f"Wrong number of arguments for component {name[bad]}")
    # ---- validation (Component.check_input, nodal/nodal.py:150-178)
    if m and (counts < 5).any():
        bad = int(np.flatnonzero(counts < 5)[0])
        raise ValueError(f"Missing arguments for component {name[bad]}")
    import pandas as pd
    kcodes, kuniq = pd.factorize(kind, sort=False)
    known_u = np.array([t in K.NODE_TYPES for t in kuniq], dtype=bool)
    if m and not known_u[kcodes].all():
        bad = int(np.flatnonzero(~known_u[kcodes])[0])
        raise ValueError(f"Unknown type {kind[bad]} for component {name[bad]}")
    want = np.array([K.NODE_ARGS_NUMBER[t] for t in kuniq], dtype=np.int64)[kcodes] if m else np.zeros(0, np.int64)
    if (counts != want).any():
        bad = int(np.flatnonzero(counts != want)[0])
        raise ValueError(f"Wrong number of arguments for component {name[bad]}: expected {want[bad]}, "
                         f"got {counts[bad]}")
    if (kind == "OPAMP").any():
        raise NotImplementedError
    # ---- OPMODEL rows are expanded and appended after all csv rows (nodal.py:231-234,273-277)
    op = kind == "OPMODEL"
    if op.any():
        extra = []
        for k in np.flatnonzero(op):
            float(value[k]) if value[k] != "0" else None       # value must parse, as Component() would check
            extra.extend(build_opmodel([name[k], kind[k], value[k], a[k], b[k], c_[k], d_[k]]))
        keep = ~op
        pad = lambda row: row + [None] * (8 - len(row))                      # noqa: E731
        ext = np.array([pad(r) for r in extra], dtype=object).reshape(-1, 8)
        name, kind, value, a, b, c_, d_, drv = (np.concatenate([col[keep], ext[:, k]])
                                                for k, col in enumerate((name, kind, value, a, b, c_, d_, drv)))
        m = len(name)
    try:
        val = value.astype(np.float64) if m else np.zeros(0)
    except ValueError:
        for k in range(m):
            try:
                float(value[k])
            except ValueError:
                raise ValueError("Bad input: expected a number for component value "
                                 f"of {name[k]}, got {value[k]} instead")
        raise
    # ---- first-appearance node numbering (nodal.py:249-257): anode, then bnode, per component
    inter = np.empty(2 * m, dtype=object)
    inter[0::2] = a
    inter[1::2] = b
    codes, uniques = pd.factorize(inter, sort=False)            # codes in first-appearance order
    labels = list(uniques)
    degree = np.bincount(codes, minlength=len(labels))
    degrees = dict(zip(labels, degree.tolist()))
    ground = find_ground_node(degrees) if labels else None
    gcode = labels.index(ground) if labels else -1
    index_of_code = np.arange(len(labels), dtype=np.int32)
    index_of_code[gcode + 1:] -= 1
    if labels:
        index_of_code[gcode] = K.GROUND
    nodenum = {lab: int(index_of_code[i]) for i, lab in enumerate(labels) if i != gcode}
    a_idx = index_of_code[codes[0::2]] if m else np.zeros(0, np.int32)
    b_idx = index_of_code[codes[1::2]] if m else np.zeros(0, np.int32)
    # ---- type codes, branch numbers (nodal.py:251-253), controls, drivers
    kcodes, kuniq = pd.factorize(kind, sort=False)             # again: OPMODEL rows were replaced
    tcode = np.array([K.TYPE_CODE[t] for t in kuniq], dtype=np.uint8)[kcodes] if m else np.zeros(0, np.uint8)
    anom = np.array([t in K.NODE_TYPES_ANOM for t in kuniq], dtype=bool)[kcodes] if m else np.zeros(0, bool)
    branch = np.full(m, -1, dtype=np.int32)
    branch[anom] = np.arange(int(anom.sum()), dtype=np.int32)
    c_idx = np.full(m, K.UNUSED, dtype=np.int32)
    d_idx = np.full(m, K.UNUSED, dtype=np.int32)
    drv_idx = np.full(m, -1, dtype=np.int32)
    dep = np.array([t in K.NODE_TYPES_DEP for t in kuniq], dtype=bool)[kcodes] if m else np.zeros(0, bool)

    def node_index(label):
        if label == ground:
            return K.GROUND
        return nodenum[label]                                   # KeyError as models.py:74,77

    row_of_name = {}
    if any(t in K.NODE_TYPES_CC for t in kuniq):
        for k in range(m):                                      # first row of every name (duplicates keep order)
            row_of_name.setdefault(name[k], k)
    for k in np.flatnonzero(dep):
        c_idx[k] = node_index(c_[k])
        d_idx[k] = node_index(d_[k])
        if kind[k] in K.NODE_TYPES_CC:
            if drv[k] not in row_of_name:
                raise KeyError(f"Driving component {drv[k]} not found")
            j = row_of_name[drv[k]]
            assert (c_[k] == a[j] and d_[k] == b[j]) or (c_[k] == b[j] and d_[k] == a[j])
            drv_idx[k] = j
    kcl = len(nodenum)
    be = int(anom.sum())
    table = ComponentTable(tcode, val, a_idx, b_idx, c_idx, d_idx, drv_idx, branch, kcl=kcl, be=be)
    anomnum = {name[k]: int(branch[k]) for k in np.flatnonzero(anom)}
    names = list(name)
    net = TableNetlist(table, nodenum, ground, names=names.__getitem__, anomnum=anomnum)
    net._currents = [names[k] for k in np.flatnonzero(anom)]
    net._degrees = degrees
    net.nums["components"] = m
    return net
