#!/usr/bin/env python
"""Generate golden input/output vectors by RUNNING THE REFERENCE in the build container.

Usage (build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_golden.py

Writes tests/golden/doc_netlists.json, grids.json, check_input.json.  Inputs are
stored as parsed csv rows (the input vectors), outputs are everything the
reference computes on the hot path: numbering, dense G / A, CSR in both
column orders, dense and sparse results, printed solution.
"""
import csv
import io
import json
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("NODAL_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
import nodal as ref  # noqa: E402
import nodal.equiv  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.mna_oracle import grid2d_rows, lattice3d_rows  # noqa: E402  (generators only)


def rows_of(path):
    with open(path) as fh:
        return [r for r in csv.reader(fh, skipinitialspace=True)]


def write_rows(rows, path):
    with open(path, "w", newline="") as fh:
        w = csv.writer(fh)
        for r in rows:
            w.writerow(r)


def dump_netlist(path):
    out = {"rows": rows_of(path)}
    net = ref.Netlist(path)
    out.update(ground=net.ground, nodenum=net.nodenum, anomnum=net.anomnum, nums=net.nums,
               component_keys=net.component_keys, degrees=net.degrees,
               resistive=bool(ref.equiv.check_resistive(net)),
               connected=bool(ref.is_connected(net)))
    try:
        cd = ref.Circuit(net, sparse=False)
    except Exception as e:  # noqa: BLE001
        out["build_error"] = type(e).__name__
        return out
    out.update(G=cd.G.tolist(), A=cd.A.tolist(), currents=cd.currents)
    cs = ref.Circuit(ref.Netlist(path), sparse=True)
    G = cs.G
    out["csr_first_touch"] = dict(indptr=G.indptr.tolist(), indices=G.indices.tolist(),
                                  data=G.data.tolist(), index_dtype=str(G.indices.dtype))
    Gs = G.copy()
    Gs.sort_indices()
    out["csr_sorted"] = dict(indptr=Gs.indptr.tolist(), indices=Gs.indices.tolist(),
                             data=Gs.data.tolist())
    try:
        sol = cd.solve()
        out["result_dense"] = sol.result.tolist()
        out["printed"] = str(sol)
    except Exception as e:  # noqa: BLE001
        out["dense_error"] = type(e).__name__
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            sol = cs.solve()
            out["result_sparse"] = [None if not np.isfinite(v) else float(v) for v in sol.result]
        except Exception as e:  # noqa: BLE001
            out["sparse_error"] = type(e).__name__
    if out["resistive"] and "1" in net.degrees and "g" in net.degrees:
        out["equiv_1_g_dense"] = float(ref.equiv.equivalent_resistance(ref.Netlist(path), "1", "g"))
        out["equiv_1_g_sparse"] = float(
            ref.equiv.equivalent_resistance(ref.Netlist(path), "1", "g", sparse=True))
    return out


def main():
    doc = {}
    for fn in sorted(os.listdir(os.path.join(REF, "doc"))):
        if fn.endswith(".csv"):
            doc[fn] = dump_netlist(os.path.join(REF, "doc", fn))
    # extra netlists exercising stamp corner cases
    extra = {
        "x_vcvs_shared_control.csv": [  # c == a: entry becomes 1 - gain (models.py:73-78)
            ["r1", "R", "2", "1", "g"], ["r2", "R", "3", "2", "g"], ["r3", "R", "5", "1", "2"],
            ["e1", "E", "1.5", "3", "g"], ["r4", "R", "7", "3", "1"],
            ["d1", "VCVS", "0.25", "2", "g", "2", "1"]],
        "x_vcvs_null_control.csv": [    # c == d: -gain + gain == 0 -> DOK deletes the key
            ["r1", "R", "2", "1", "g"], ["r2", "R", "3", "2", "g"], ["r3", "R", "5", "1", "2"],
            ["a1", "A", "0.5", "1", "g"],
            ["d1", "VCVS", "3", "2", "g", "1", "1"]],
        "x_ccvs_overwrite.csv": [       # CCVS control coincides with its own leads: '=' overwrites +-1
            ["r1", "R", "4", "1", "2"], ["r2", "R", "2", "2", "g"], ["e1", "E", "3", "1", "g"],
            ["r3", "R", "8", "3", "g"],
            ["h1", "CCVS", "5", "3", "g", "1", "2", "r1"]],
        "x_parallel_and_selfloop.csv": [  # parallel resistors (sum order) + self loop (exact zero)
            ["r1", "R", "3", "1", "g"], ["r2", "R", "7", "1", "g"], ["r3", "R", "0.1", "1", "2"],
            ["r4", "R", "0.3", "2", "1"], ["r5", "R", "11", "2", "2"], ["r6", "R", "13", "2", "g"],
            ["a1", "A", "1", "1", "g"]],
        "x_no_g_ground.csv": [          # ground chosen by degree, ties -> first inserted
            ["r1", "R", "1", "a", "b"], ["r2", "R", "2", "b", "c"], ["r3", "R", "3", "c", "a"],
            ["r4", "R", "4", "b", "d"], ["a1", "A", "2", "a", "d"]],
        "x_cccs_mixed.csv": [
            ["r1", "R", "2", "1", "2"], ["r2", "R", "4", "2", "g"], ["e1", "E", "6", "1", "g"],
            ["f1", "CCCS", "3", "3", "g", "1", "2", "r1"], ["r3", "R", "5", "3", "g"],
            ["q1", "OPMODEL", "1000", "4", "g", "3", "5"], ["r4", "R", "500", "5", "g"],
            ["r5", "R", "100", "4", "g"]],
        "x_duplicate_name.csv": [      # components[key] is overwritten (nodal.py:243): last row wins,
            ["r1", "R", "10", "a", "g"], ["r1", "R", "20", "b", "a"],   # stamped once per occurrence
            ["r2", "R", "5", "b", "g"], ["a1", "A", "1", "b", "g"]],
        "x_negative_resistor.csv": [   # g + (-g) passes through exact zero then re-inserts
            ["r1", "R", "2", "1", "g"], ["r2", "R", "-2", "1", "g"], ["r3", "R", "4", "1", "g"],
            ["r4", "R", "1", "1", "2"], ["r5", "R", "1", "2", "g"], ["a1", "A", "1", "2", "g"]],
    }
    tmp = "/tmp/_nodal_golden"
    os.makedirs(tmp, exist_ok=True)
    for name, rows in extra.items():
        p = os.path.join(tmp, name)
        write_rows(rows, p)
        doc[name] = dump_netlist(p)

    with open(os.path.join(HERE, "doc_netlists.json"), "w") as fh:
        json.dump(doc, fh, indent=0, sort_keys=True)

    # --- expected strings pinned by the reference's own tests (tests.py:15,52-122)
    pinned = {
        "equiv": {"resistive_1.csv": 2.0, "resistive_2.csv": 1.0, "resistive_3.csv": 1.0},
        "printed": {
            "1.6.1.csv": "Ground node: g\ne(1) \t= 2.0\ne(2) \t= -1.0\ne(4) \t= 8.0\n"
                         "i(d1) \t= -1.9999999999999998\ni(e1) \t= 3.0",
            "netlist.csv": "Ground node: 1\ne(2) \t= -1.0\ne(3) \t= -2.0",
            "opmodel_amplifier.csv": "Ground node: g\ne(1) \t= 0.9998800143982737\n"
                                     "e(2) \t= 1.9997600287845492\ne(3) \t= 1.0000000000000002\n"
                                     "e(q1_internal_node) \t= 11.998560172647306\n"
                                     "i(q1_vcvs) \t= 0.9998800143862756\ni(v1) \t= 1.1998560172647305e-11",
            "test_1.csv": "Ground node: g\ne(1) \t= 1.0\ne(2) \t= 1.0\ne(3) \t= 1.0\ne(4) \t= 1.0\n"
                          "e(5) \t= 1.0\ne(6) \t= 1.0\ni(d1) \t= -0.0\ni(d2) \t= -0.0\ni(d3) \t= 1.0\n"
                          "i(d4) \t= 1.0\ni(e1) \t= -0.0",
            "buffer.csv": "Ground node: g\ne(1) \t= 9.999900000999991\ne(2) \t= 9.999900000899993\n"
                          "e(3) \t= 10.0\ni(d1) \t= -9.999889805101247e-12\ni(vs) \t= 9.999900000899993e-12",
            "opmodel_voltage_buffer.csv": "Ground node: g\ne(2) \t= 0.999990000099999\n"
                                          "e(3) \t= 0.9999999999999999\n"
                                          "e(q1_internal_node) \t= 0.9999900000899992\n"
                                          "i(q1_vcvs) \t= -9.999917560676863e-13\n"
                                          "i(v1) \t= 9.999900000899992e-13",
        },
    }
    with open(os.path.join(HERE, "pinned_by_reference_tests.json"), "w") as fh:
        json.dump(pinned, fh, indent=1, sort_keys=True)

    # --- grids (config C2 family) through the reference end to end
    grids = {}
    for N, modes in ((6, ("dense", "sparse")), (20, ("dense", "sparse")), (50, ("sparse",)), (100, ("sparse",))):
        p = os.path.join(tmp, f"grid{N}.csv")
        write_rows(grid2d_rows(N), p)
        net = ref.Netlist(p)
        entry = dict(N=N, kcl=net.nums["kcl"], ground=net.ground,
                     nodenum_head=dict(list(net.nodenum.items())[: 4 * N]))
        for mode in modes:
            r = ref.equiv.equivalent_resistance(ref.Netlist(p), "1", "g", sparse=(mode == "sparse"))
            entry[f"R_{mode}"] = float(r)
        c = ref.Circuit(ref.Netlist(p), sparse=True)
        G = c.G.copy()
        G.sort_indices()
        entry["nnz"] = int(G.nnz)
        if N <= 20:
            entry["csr_sorted"] = dict(indptr=G.indptr.tolist(), indices=G.indices.tolist(),
                                       data=G.data.tolist())
        else:  # checksums only
            entry["indptr_sum"] = int(G.indptr.astype(np.int64).sum())
            entry["indices_sum"] = int(G.indices.astype(np.int64).sum())
            entry["indices_wsum"] = int((G.indices.astype(np.int64) * (np.arange(G.nnz) % 1009)).sum())
            entry["data_sum"] = float(G.data.sum())
        grids[f"grid2d_{N}"] = entry
    for N in (5, 6):
        p = os.path.join(tmp, f"lat{N}.csv")
        write_rows(lattice3d_rows(N), p)
        net = ref.Netlist(p)
        c = ref.Circuit(ref.Netlist(p), sparse=True)
        G = c.G.copy()
        G.sort_indices()
        grids[f"lattice3d_{N}"] = dict(
            N=N, kcl=net.nums["kcl"], ground=net.ground, nnz=int(G.nnz),
            R_dense=float(ref.equiv.equivalent_resistance(ref.Netlist(p), "1", "g")),
            R_sparse=float(ref.equiv.equivalent_resistance(ref.Netlist(p), "1", "g", sparse=True)),
            csr_sorted=dict(indptr=G.indptr.tolist(), indices=G.indices.tolist(), data=G.data.tolist()))
    with open(os.path.join(HERE, "grids.json"), "w") as fh:
        json.dump(grids, fh, indent=0, sort_keys=True)

    # --- Component.check_input verdicts (nodal.py:150-178) on rows written for this repo
    good = ["r1,R,1,a,b", "a1,A,1e-3,a,b", "e1,E,-5,a,g", "v1,VCVS,5,1,2,3,4", "v1,VCCS,5,1,2,3,4",
            "c1,CCCS,2,1,2,3,4,r1", "c1,CCVS,2,1,2,3,4,r1", "q1,OPMODEL,1,2,g,3,1",
            "q1,OPAMP,1,2,g,3,1", "r1, R, 1e7, 1, 3", "", "# comment, with, commas", "#x"]
    bad = ["zzz", "r1,R,1,a", "r1,R,1,a,b,c", "r1,X,1,a,b", "r1,r,1,a,b", "r1,R,abc,a,b",
           "v1,VCVS,5,1,2", "v1,VCVS,5,1,2,3,4,5", "c1,CCCS,2,1,2,3,4", "c1,CCVS,2,1,2,3,4,r1,x",
           "q1,OPMODEL,1,2,g,3", "e1,E,,a,b", "a1,A,1 2,a,b"]
    verdicts = {}
    for line in good + bad:
        row = next(csv.reader(io.StringIO(line), skipinitialspace=True), [])
        try:
            ref.Component.check_input(None, row)
            verdicts[line] = "ok"
        except Exception as e:  # noqa: BLE001
            verdicts[line] = type(e).__name__
    with open(os.path.join(HERE, "check_input.json"), "w") as fh:
        json.dump(verdicts, fh, indent=1, sort_keys=True)
    print("golden files written to", HERE)


if __name__ == "__main__":
    main()
