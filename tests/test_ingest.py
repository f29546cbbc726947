"""Vectorised csv ingest (nodal_b200.ingest) == the row-by-row Netlist, which is checked against
the reference's numbering in test_host.py."""
import numpy as np
import pytest

import nodal_b200 as n
from helpers import golden, write_csv
from nodal_b200 import generators as gen
from nodal_b200.ingest import read_table_netlist
from oracle import mna_oracle as orc

DOC = golden("doc_netlists.json")


def same_numbering(path):
    slow = n.Netlist(path)
    fast = read_table_netlist(path)
    assert fast.ground == slow.ground
    assert dict(fast.nodenum) == slow.nodenum and list(fast.nodenum) == list(slow.nodenum)
    assert fast.anomnum == slow.anomnum
    assert fast.component_keys == slow.component_keys
    assert fast.degrees == slow.degrees
    for k in ("components", "anomalies", "be", "kcl"):
        assert fast.nums[k] == slow.nums[k], k
    return slow, fast


@pytest.mark.parametrize("name", sorted(DOC))
def test_doc_netlists(name, tmp_path):
    g = DOC[name]
    path = write_csv(g["rows"], tmp_path / name)
    slow, fast = same_numbering(path)
    assert fast.ground == g["ground"] and dict(fast.nodenum) == g["nodenum"] and fast.anomnum == g["anomnum"]
    if "G" not in g:
        return
    t1, c1 = slow.table_and_currents()
    t2, c2 = fast.table_and_currents()
    assert c1 == c2 == g["currents"]
    for col in ("type", "value", "a", "b", "c", "d", "drv", "branch"):
        assert np.array_equal(getattr(t1, col), getattr(t2, col)), col
    assert (t1.kcl, t1.be) == (t2.kcl, t2.be)


def test_comments_blank_lines_and_spacing(tmp_path):
    text = ("# header, with, commas\n\nRi, R, 1e7, 1, 3\n  Ro,R,1e1,1 ,2\n#x\nvs, E, 10, 3, g\n"
            "d1, VCVS, 1e5, 2, g, 3, 1 \nq1,OPMODEL,1,2,g,3,1\n")
    p = tmp_path / "s.csv"
    p.write_text(text)
    same_numbering(str(p))


def test_generated_netlists(tmp_path):
    rows = gen.random_opamp_network_rows(M=200, P=20, S=15, V=6, seed=5)
    same_numbering(write_csv(rows, tmp_path / "c3.csv"))
    path = write_csv(orc.grid2d_rows(30), tmp_path / "g.csv")
    slow, fast = same_numbering(path)
    tn = gen.grid2d(30)
    assert dict(fast.nodenum) == dict(tn.nodenum)
    for col in ("type", "value", "a", "b"):
        assert np.array_equal(getattr(fast.table(), col), getattr(tn.table(), col))
    # no-"g" netlist: ground by degree, first maximum
    rows = [["r1", "R", "1", "a", "b"], ["r2", "R", "2", "b", "c"], ["r3", "R", "3", "c", "a"],
            ["r4", "R", "4", "b", "d"], ["a1", "A", "2", "a", "d"]]
    same_numbering(write_csv(rows, tmp_path / "ng.csv"))


@pytest.mark.parametrize("rows,err", [
    ([["r1", "R", "1", "a"]], ValueError),
    ([["r1", "X", "1", "a", "b"]], ValueError),
    ([["r1", "R", "abc", "a", "b"]], ValueError),
    ([["v1", "VCVS", "5", "1", "2"]], ValueError),
    ([["r1", "R", "1", "1", "g"], ["h", "CCVS", "1", "2", "g", "1", "g", "nope"]], KeyError),
    ([["r1", "R", "1", "1", "g"], ["d", "VCVS", "1", "2", "g", "zz", "g"]], KeyError),
    ([["q", "OPAMP", "1", "2", "g", "3", "1"]], NotImplementedError),
])
def test_errors_match_row_by_row_path(rows, err, tmp_path):
    path = write_csv(rows, tmp_path / "bad.csv")
    with pytest.raises(err):
        read_table_netlist(path)
    with pytest.raises(err):
        n.Netlist(path).table()


def test_speed_on_a_large_file(tmp_path):
    import time
    path = write_csv(orc.grid2d_rows(250), tmp_path / "big.csv")      # 124 500 rows
    t0 = time.perf_counter(); fast = read_table_netlist(path); t_fast = time.perf_counter() - t0
    t0 = time.perf_counter(); slow = n.Netlist(path); t1 = time.perf_counter() - t0
    assert dict(fast.nodenum) == slow.nodenum
    assert t_fast < t1          # row-by-row python vs pandas + numpy
