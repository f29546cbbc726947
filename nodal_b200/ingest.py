"""Vectorised netlist ingest: csv -> component table with the reference's numbering.

The reference parses one row at a time into a dict of ``Component`` objects
(nodal/nodal.py:222-296, ~7 us per row) and ``equivalent_resistance`` then deep-copies
that dict (nodal/equiv.py:50, ~18 us per component): at 33.5 M rows that is tens of
minutes before any arithmetic.  ``read_table_netlist`` produces the same numbering --
component order, first-appearance node discovery (anode then bnode), ground choice,
branch numbers, OPMODEL expansion appended after all csv rows -- and returns a
``TableNetlist``.

How: the file is scanned once as bytes with numpy (line boundaries, comment / blank lines,
fields per line); lines are grouped by their number of fields and every group is parsed by
Arrow's multi-threaded csv reader as string columns; node labels are numbered with Arrow's
dictionary encoding (indices in first-appearance order), values are converted by Arrow's
correctly rounded float parser.  Anything the fast path does not handle exactly like
``csv.reader(skipinitialspace=True)`` + ``float()`` -- quoted fields, carriage returns,
numbers Arrow rejects -- goes through the csv module / ``float`` instead.

Rows are validated with the same rules as ``Component.check_input`` (arity per type,
known type, numeric value); errors are raised as ValueError like the reference does.
"""
from __future__ import annotations

import csv
import io
from collections.abc import Mapping

import numpy as np

from . import constants as K
from .generators import TableNetlist
from .nodal import build_opmodel
from .table import ComponentTable

NCOLS = 8          # name, type, value, anode, bnode, pos_control, neg_control, driver


class DuplicateNameError(ValueError):
    """Two rows share a component name: only the row-by-row Netlist reproduces the reference."""


def _arrow():
    import pyarrow as pa
    import pyarrow.compute as pc
    import pyarrow.csv as pacsv
    return pa, pc, pacsv


def _numpy(arr, dtype):
    """Zero-copy view of a primitive Arrow array without nulls (Array.to_numpy imports pandas,
    2.5 s the first time)."""
    assert arr.null_count == 0
    return np.frombuffer(arr.buffers()[1], dtype=dtype)[arr.offset: arr.offset + len(arr)]


def _pylist(arr):
    """Python strings of an Arrow string array; one decode + slices instead of a scalar per item."""
    pa, _, _ = _arrow()
    if arr.null_count == 0 and arr.type == pa.string() and len(arr) > 256:
        bufs = arr.buffers()
        data = bufs[2].to_pybytes() if bufs[2] is not None else b""
        if data.isascii():                              # byte offsets are character offsets
            text = data.decode("ascii")
            off = np.frombuffer(bufs[1], dtype=np.int32)[arr.offset: arr.offset + len(arr) + 1].tolist()
            return [text[off[k]: off[k + 1]] for k in range(len(arr))]
    return arr.to_pylist()


class _LabelMap(Mapping):
    """label -> integer, in first-appearance order.  The python dict (0.5 us per entry) is only
    built when somebody looks a label up or iterates; the numbering itself never needs it."""

    def __init__(self, dictionary, values, skip=-1):
        self._dictionary, self._values, self._skip = dictionary, values, int(skip)
        self._dict = None

    def _built(self):
        if self._dict is None:
            d = dict(zip(_pylist(self._dictionary), self._values.tolist()))
            if self._skip >= 0:
                d.pop(self._dictionary[self._skip].as_py(), None)
            self._dict = d
        return self._dict

    def __getitem__(self, key):
        return self._built()[key]

    def __contains__(self, key):
        return key in self._built()

    def __iter__(self):
        return iter(self._built())

    def __len__(self):
        return len(self._dictionary) - (1 if self._skip >= 0 else 0)

    def __eq__(self, other):
        return self._built() == (other._built() if isinstance(other, _LabelMap) else other)

    def __repr__(self):
        return f"<{len(self)} labels>" if self._dict is None else repr(self._dict)


def _rows_with_csv_module(text):
    """Exact csv.reader(skipinitialspace=True) semantics (nodal/nodal.py:268); comment and
    blank rows dropped.  Returns 8 Arrow string columns (null = field absent) + field counts."""
    pa, _, _ = _arrow()
    rows = [r for r in csv.reader(io.StringIO(text), skipinitialspace=True) if r and r[0][:1] != "#"]
    counts = np.array([len(r) for r in rows], dtype=np.int64)
    cols = [pa.array([r[k] if k < len(r) else None for r in rows], type=pa.string()) for k in range(NCOLS)]
    return cols, counts


def _scan_lines(buf):
    """Line structure of a byte buffer that ends with a newline: (starts, ends, keep, counts) --
    `ends` are the positions of the newlines, `keep` drops blank and comment lines, `counts` is
    the number of comma separated fields."""
    ends = np.flatnonzero(buf == 10)
    starts = np.empty_like(ends)
    starts[0] = 0
    starts[1:] = ends[:-1] + 1
    first = buf[starts]
    for i in np.flatnonzero(first == 32):            # leading blanks are rare: look past them one line at a time
        line = buf[starts[i]: ends[i] + 1]
        first[i] = line[np.argmax(line != 32)]
    keep = (first != 10) & (first != 35)             # neither empty nor a '#' comment
    commas = np.flatnonzero(buf == 44)
    counts = np.searchsorted(commas, ends) - np.searchsorted(commas, starts) + 1
    return starts, ends, keep, counts


def _parse_group(data, nfields):
    """Arrow csv parse of lines that all have `nfields` fields -> list of string arrays with the
    blanks after every delimiter removed (csv.reader's skipinitialspace)."""
    pa, pc, pacsv = _arrow()
    names = [f"f{k}" for k in range(nfields)]
    table = pacsv.read_csv(
        pa.BufferReader(data),
        read_options=pacsv.ReadOptions(column_names=names, block_size=32 << 20),
        parse_options=pacsv.ParseOptions(delimiter=",", quote_char=False, double_quote=False,
                                         escape_char=False, newlines_in_values=False,
                                         ignore_empty_lines=False),
        convert_options=pacsv.ConvertOptions(column_types={c: pa.string() for c in names}, null_values=[],
                                             strings_can_be_null=False, quoted_strings_can_be_null=False))
    return [pc.utf8_ltrim(table.column(c).combine_chunks(), characters=" ") for c in names]


def _rows_with_arrow(raw):
    """Fast path; returns None when the input needs the csv module."""
    pa, pc, _ = _arrow()
    if b"\r" in raw:
        raw = raw.replace(b"\r\n", b"\n")            # csv.writer's default line terminator
    if b'"' in raw or b"\r" in raw:
        return None
    if not raw.endswith(b"\n"):
        raw = raw + b"\n"
    buf = np.frombuffer(raw, dtype=np.uint8)
    starts, ends, keep, counts = _scan_lines(buf)
    counts = counts[keep]
    m = len(counts)
    if m == 0:
        return [pa.array([], type=pa.string()) for _ in range(NCOLS)], counts.astype(np.int64)
    if counts.max() > NCOLS + 1:
        return None                                   # reported by the csv-module path
    kinds = np.unique(counts)
    if len(kinds) == 1 and keep.all():
        groups = [(int(kinds[0]), None, raw)]
    else:
        # bytes of the lines of every group, in file order (the newline belongs to its line)
        line_of_byte = np.repeat(np.arange(len(ends), dtype=np.int32), ends - starts + 1)
        kept_rank = np.cumsum(keep) - 1               # row number of a kept line
        groups = []
        full = np.zeros(len(ends), dtype=np.int64)
        full[keep] = counts
        for k in kinds:
            sel = full == k
            groups.append((int(k), kept_rank[sel], buf[sel[line_of_byte]].tobytes()))
    cols = [[] for _ in range(NCOLS)]
    order = []
    for k, rows_of_group, data in groups:
        parsed = _parse_group(data, k)
        size = len(parsed[0])
        for c in range(NCOLS):
            cols[c].append(parsed[c] if c < k else pa.nulls(size, type=pa.string()))
        if rows_of_group is not None:
            order.append(rows_of_group)
    cols = [pa.concat_arrays(parts) if len(parts) > 1 else parts[0] for parts in cols]
    if order:
        position = np.empty(m, dtype=np.int64)        # where row r sits in the concatenation
        position[np.concatenate(order)] = np.arange(m)
        cols = [c.take(pa.array(position)) for c in cols]
    if len(cols[0]) != m:
        return None
    return cols, counts.astype(np.int64)


def _read_rows(path):
    with open(path, "rb") as fh:
        raw = fh.read()
    out = None
    try:
        out = _rows_with_arrow(raw)
    except Exception:            # any disagreement about the format: the csv module decides
        out = None
    if out is None:
        out = _rows_with_csv_module(raw.decode("utf-8"))
    return out


def _to_float(value, name):
    """float() of every value string (Component.__init__, nodal/nodal.py:135)."""
    pa, pc, _ = _arrow()
    try:
        return np.array(_numpy(pc.cast(value, pa.float64()), np.float64))
    except (pa.ArrowInvalid, pa.ArrowNotImplementedError):
        pass
    out = np.empty(len(value), dtype=np.float64)
    for k, text in enumerate(_pylist(value)):         # python's float(): blanks, underscores, ...
        try:
            out[k] = float(text)
        except (TypeError, ValueError):
            raise ValueError("Bad input: expected a number for component value "
                             f"of {name[k].as_py()}, got {text} instead")
    return out


def read_table_netlist(path):
    """Parse `path` into a TableNetlist whose numbering is bit-identical to
    ``nodal.Netlist(path)`` (checked in tests/test_ingest.py)."""
    pa, pc, _ = _arrow()
    cols, counts = _read_rows(path)
    name, kind, value, a, b, c_, d_, drv = cols
    m = len(name)

    def label(col, k):
        return col[int(k)].as_py()

    if m and (counts > NCOLS).any():
        bad = np.flatnonzero(counts > NCOLS)[0]
        raise ValueError(f"Wrong number of arguments for component {label(name, bad)}")
    # ---- validation (Component.check_input, nodal/nodal.py:150-178)
    if m and (counts < 5).any():
        bad = np.flatnonzero(counts < 5)[0]
        raise ValueError(f"Missing arguments for component {label(name, bad)}")

    def type_codes(kind):
        enc = pc.dictionary_encode(kind)
        return _numpy(enc.indices, np.int32).astype(np.int64), enc.dictionary.to_pylist()

    kcodes, kuniq = type_codes(kind) if m else (np.zeros(0, np.int64), [])
    known_u = np.array([t in K.NODE_TYPES for t in kuniq], dtype=bool)
    if m and not known_u[kcodes].all():
        bad = np.flatnonzero(~known_u[kcodes])[0]
        raise ValueError(f"Unknown type {label(kind, bad)} for component {label(name, bad)}")
    want = np.array([K.NODE_ARGS_NUMBER[t] for t in kuniq], dtype=np.int64)[kcodes] if m else np.zeros(0, np.int64)
    if (counts != want).any():
        bad = np.flatnonzero(counts != want)[0]
        raise ValueError(f"Wrong number of arguments for component {label(name, bad)}: expected {want[bad]}, "
                         f"got {counts[bad]}")
    if "OPAMP" in kuniq:
        raise NotImplementedError
    # ---- OPMODEL rows are expanded and appended after all csv rows (nodal.py:231-234,273-277)
    if "OPMODEL" in kuniq:
        op = kcodes == kuniq.index("OPMODEL")
        extra = []
        for k in np.flatnonzero(op):
            row = [col[int(k)].as_py() for col in (name, kind, value, a, b, c_, d_)]
            if row[2] != "0":
                float(row[2])                          # the value must parse, as Component() would check
            extra.extend(build_opmodel(row))
        keep = pa.array(~op)
        tail = [pa.array([r[c] if c < len(r) else None for r in extra], type=pa.string()) for c in range(NCOLS)]
        name, kind, value, a, b, c_, d_, drv = (pa.concat_arrays([col.filter(keep), tail[c]])
                                                for c, col in enumerate((name, kind, value, a, b, c_, d_, drv)))
        m = len(name)
        kcodes, kuniq = type_codes(kind)               # again: OPMODEL rows were replaced
    # ---- duplicate component names: the reference keeps ONE record per name (the last row wins,
    # nodal.py:243) and stamps it once per occurrence, with the branch number of the last
    # anomalous occurrence.  That is not a per-row table any more: refuse, so that callers
    # (cli.load_netlist_or_exit) fall back to the row-by-row Netlist, which reproduces it.
    if m:
        ncodes = _numpy(pc.dictionary_encode(name).indices, np.int32)
        if int(ncodes.max()) + 1 != m:
            first = np.flatnonzero(np.bincount(ncodes, minlength=m)[ncodes] > 1)[0]
            raise DuplicateNameError(f"component name {label(name, first)!r} is used more than once; "
                                     "use nodal_b200.Netlist (row by row) for this file")
    val = _to_float(value, name) if m else np.zeros(0)
    # ---- first-appearance node numbering (nodal.py:249-257): anode, then bnode, per component
    if m:
        interleave = np.empty(2 * m, dtype=np.int64)
        interleave[0::2] = np.arange(m)
        interleave[1::2] = np.arange(m, 2 * m)
        enc = pc.dictionary_encode(pa.concat_arrays([a, b]).take(pa.array(interleave)))
        codes = _numpy(enc.indices, np.int32).astype(np.int64)               # first-appearance order
        uniques = enc.dictionary
    else:
        codes, uniques = np.zeros(0, np.int64), pa.array([], type=pa.string())
    nlabels = len(uniques)
    degree = np.bincount(codes, minlength=nlabels)
    degrees = _LabelMap(uniques, degree)
    # ground (find_ground_node, nodal.py:30-42): "g" when present, else the first label with the
    # largest degree
    gcode = -1
    if nlabels:
        gcode = pc.index(uniques, pa.scalar("g", type=pa.string())).as_py()
        if gcode < 0:
            gcode = int(np.argmax(degree))
    ground = uniques[gcode].as_py() if nlabels else None
    index_of_code = np.arange(nlabels, dtype=np.int32)
    index_of_code[gcode + 1:] -= 1
    if nlabels:
        index_of_code[gcode] = K.GROUND
    nodenum = _LabelMap(uniques, index_of_code, skip=gcode)
    a_idx = index_of_code[codes[0::2]] if m else np.zeros(0, np.int32)
    b_idx = index_of_code[codes[1::2]] if m else np.zeros(0, np.int32)
    # ---- type codes, branch numbers (nodal.py:251-253), controls, drivers
    per_type = lambda pred, dtype: (np.array([pred(t) for t in kuniq], dtype=dtype)[kcodes]      # noqa: E731
                                    if m else np.zeros(0, dtype))
    tcode = per_type(lambda t: K.TYPE_CODE[t], np.uint8)
    anom = per_type(lambda t: t in K.NODE_TYPES_ANOM, bool)
    dep = per_type(lambda t: t in K.NODE_TYPES_DEP, bool)
    cc = per_type(lambda t: t in K.NODE_TYPES_CC, bool)
    branch = np.full(m, -1, dtype=np.int32)
    branch[anom] = np.arange(int(anom.sum()), dtype=np.int32)
    c_idx = np.full(m, K.UNUSED, dtype=np.int32)
    d_idx = np.full(m, K.UNUSED, dtype=np.int32)
    drv_idx = np.full(m, -1, dtype=np.int32)

    def node_index(node):
        if node == ground:
            return K.GROUND
        return nodenum[node]                            # KeyError as models.py:74,77

    row_of_name = {}
    if cc.any():
        for k, text in enumerate(_pylist(name)):        # first row of every name (duplicates keep order)
            row_of_name.setdefault(text, k)
    for k in np.flatnonzero(dep):
        ck, dk = label(c_, k), label(d_, k)
        c_idx[k] = node_index(ck)
        d_idx[k] = node_index(dk)
        if cc[k]:
            driver = label(drv, k)
            if driver not in row_of_name:
                raise KeyError(f"Driving component {driver} not found")
            j = row_of_name[driver]
            aj, bj = label(a, j), label(b, j)
            assert (ck == aj and dk == bj) or (ck == bj and dk == aj)
            drv_idx[k] = j
    kcl = len(nodenum)
    be = int(anom.sum())
    table = ComponentTable(tcode, val, a_idx, b_idx, c_idx, d_idx, drv_idx, branch, kcl=kcl, be=be)
    anom_rows = np.flatnonzero(anom)
    anomnum = {label(name, k): int(branch[k]) for k in anom_rows}
    net = TableNetlist(table, nodenum, ground, names=lambda k: name[int(k)].as_py(), anomnum=anomnum)
    net._component_names = name
    net._currents = [label(name, k) for k in anom_rows]
    net._degrees = degrees
    net.nums["components"] = m
    return net


# --------------------------------------------------------------------------- binary form
# A 33.5 M-row csv is 1 GB of text and tens of seconds of parsing; the same netlist as arrays
# loads at disk speed.  Layout (numpy .npz, uncompressed): the eight table columns, kcl, be, the
# ground label, and three newline-joined utf-8 blobs -- node labels in row order, component names
# in table order, names of the anomalous branches in branch order.
_BLOBS = ("node_labels", "component_names", "branch_names")


def _join(strings):
    return np.frombuffer("\n".join(strings).encode("utf-8"), dtype=np.uint8)


def _split(blob, count):
    """Arrow string array (zero copy) of a newline-joined utf-8 blob holding `count` strings."""
    pa, _, _ = _arrow()
    if count == 0:
        return pa.array([], type=pa.string())
    ends = np.flatnonzero(blob == 10)
    assert len(ends) == count - 1, "corrupt netlist file: label count does not match"
    offsets = np.empty(count + 1, dtype=np.int32)
    offsets[0] = 0
    offsets[1:count] = ends + 1
    offsets[count] = len(blob) + 1
    # string k is blob[offsets[k] : offsets[k + 1] - 1]: re-pack without the separators
    lengths = np.diff(offsets) - 1
    packed = np.delete(blob, ends) if len(ends) else blob
    new_off = np.zeros(count + 1, dtype=np.int32)
    np.cumsum(lengths, out=new_off[1:])
    return pa.StringArray.from_buffers(count, pa.py_buffer(new_off.tobytes()), pa.py_buffer(packed.tobytes()))


def save_table_netlist(net, path):
    """Write a TableNetlist (from read_table_netlist, the generators, or any Netlist through its
    table) as one .npz file; load_table_netlist restores the same numbering."""
    table, currents = net.table_and_currents()
    labels = list(net.nodenum)                             # iteration order == row order
    for row, lab in enumerate(labels[:1000]):
        assert net.nodenum[lab] == row
    if any("\n" in s for s in labels):
        raise ValueError("node labels with newlines cannot be stored")
    branch = sorted(net.anomnum, key=net.anomnum.get)
    payload = {name: getattr(table, name) for name in ("type", "value", "a", "b", "c", "d", "drv", "branch")}
    payload.update(kcl=np.int64(table.kcl), be=np.int64(table.be), ground=_join([str(net.ground)]),
                   node_labels=_join(labels), component_names=_join(list(net.component_keys)),
                   branch_names=_join(branch), currents=_join(list(currents)),
                   counts=np.array([len(labels), len(table), len(branch), len(currents)], dtype=np.int64))
    with open(path, "wb") as fh:                            # np.savez would append ".npz" to a bare name
        np.savez(fh, **payload)


def load_table_netlist(path):
    with np.load(path) as z:
        cols = {name: z[name] for name in ("type", "value", "a", "b", "c", "d", "drv", "branch")}
        kcl, be = int(z["kcl"]), int(z["be"])
        nlab, ncomp, nbranch, ncur = (int(v) for v in z["counts"])
        ground = bytes(z["ground"]).decode("utf-8")
        labels = _split(z["node_labels"], nlab)
        names = _split(z["component_names"], ncomp)
        branch = _pylist(_split(z["branch_names"], nbranch))
        currents = _pylist(_split(z["currents"], ncur))
    table = ComponentTable(cols["type"], cols["value"], cols["a"], cols["b"], cols["c"], cols["d"],
                           cols["drv"], cols["branch"], kcl=kcl, be=be)
    nodenum = _LabelMap(labels, np.arange(nlab, dtype=np.int32))
    net = TableNetlist(table, nodenum, ground, names=lambda k: names[int(k)].as_py(),
                       anomnum={name: k for k, name in enumerate(branch)})
    net._component_names = names
    net._currents = currents
    return net


def main(argv=None):
    """python -m nodal_b200.ingest NETLIST.csv NETLIST.npz: convert once, load at disk speed after."""
    import argparse
    ap = argparse.ArgumentParser(description="Convert a csv netlist to the binary table form")
    ap.add_argument("csv_path")
    ap.add_argument("npz_path")
    args = ap.parse_args(argv)
    net = read_table_netlist(args.csv_path)
    save_table_netlist(net, args.npz_path)
    print(f"{net.nums['components']} components, {net.nums['kcl']} nodes, ground {net.ground!r} -> {args.npz_path}")


if __name__ == "__main__":
    main()
