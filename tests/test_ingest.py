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
    if name == "x_duplicate_name.csv":       # refused by the table ingest; covered row by row below
        with pytest.raises(ValueError):
            read_table_netlist(path)
        return
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


def test_cli_uses_the_vectorised_ingest_for_large_files(tmp_path):
    """nodal_b200.cli.load_netlist_or_exit: >= 1 MiB -> TableNetlist with the reference numbering;
    small or malformed files -> the row-by-row Netlist (which reports errors the reference's way)."""
    from nodal_b200 import cli
    from nodal_b200.generators import TableNetlist
    from oracle import mna_oracle as orc
    rows = orc.grid2d_rows(160)
    big = write_csv(rows, tmp_path / "big.csv")
    import os
    assert os.path.getsize(big) >= cli.FAST_INGEST_BYTES
    net = cli.load_netlist_or_exit(big)
    assert isinstance(net, TableNetlist)
    ref = orc.OracleNetlist(rows)
    assert net.ground == ref.ground and dict(net.nodenum) == ref.nodenum and net.nums["kcl"] == ref.kcl
    small = write_csv(rows[:50], tmp_path / "small.csv")
    assert not isinstance(cli.load_netlist_or_exit(small), TableNetlist)
    bad = write_csv(rows + [["rx", "R", "1", "n0_0"]], tmp_path / "bad.csv")      # a row with a missing lead
    with pytest.raises(ValueError):
        cli.load_netlist_or_exit(bad)
    with pytest.raises(SystemExit):
        cli.load_netlist_or_exit(str(tmp_path / "missing.csv"))


def test_value_strings_parse_like_python_float(tmp_path):
    """Arrow's float parser (fast path) and float() (fallback) give the same bits as the
    reference's float(value) for every spelling, including the ones Arrow rejects."""
    rng = np.random.default_rng(3)
    spellings = [repr(float(v)) for v in rng.uniform(-1e6, 1e6, 300)]
    spellings += [f"{v:.17g}" for v in 10.0 ** rng.uniform(-300, 300, 300)]
    spellings += ["1e3", ".5", "5.", "+2", "-0.0", "1E-3", "0.1", "4.9e-324", "1.7976931348623157e308"]
    rows = [[f"r{k}", "R", s, f"a{k}", "g"] for k, s in enumerate(spellings)]
    fast = read_table_netlist(write_csv(rows, tmp_path / "arrow.csv"))
    want = np.array([float(s) for s in spellings])
    assert np.array_equal(fast.table().value.view(np.uint64), want.view(np.uint64))
    odd = spellings[:20] + ["1_000", " 7", "8 ", "Infinity"]       # python-only spellings -> fallback for the file
    rows = [[f"r{k}", "R", s, f"a{k}", "g"] for k, s in enumerate(odd)]
    p = tmp_path / "python.csv"
    p.write_text("".join(",".join(r) + "\n" for r in rows))
    fast = read_table_netlist(str(p))
    want = np.array([float(s) for s in odd])
    assert np.array_equal(fast.table().value.view(np.uint64), want.view(np.uint64))
    same_numbering(str(p))
    p.write_text("r1,R,abc,1,g\n")
    with pytest.raises(ValueError, match="expected a number"):
        read_table_netlist(str(p))


@pytest.mark.parametrize("newline", ["\n", "\r\n"])
def test_line_endings_field_groups_and_lazy_maps(tmp_path, newline):
    """LF and CRLF files, rows of 5 / 7 / 8 fields interleaved with comments and blank lines, a
    last line without terminator; the label maps only become dicts when they are used."""
    lines = ["# c", "r1,R,10,1,2", "", "e1,E,5,2,g", "x1,VCVS,2.5,3,g,1,2", " r2, R, 20, 3, g", "#tail",
             "f1,CCCS,3,4,g,1,2,r1", "r3,R,7,4,g", "h1,CCVS,2,5,g,2,1,r1", "r4,R,1,5,g"]
    p = tmp_path / "mixed.csv"
    p.write_bytes((newline.join(lines)).encode())
    slow, fast = same_numbering(str(p))
    t1, t2 = slow.table_and_currents()[0], fast.table_and_currents()[0]
    for col in ("type", "value", "a", "b", "c", "d", "drv", "branch"):
        assert np.array_equal(getattr(t1, col), getattr(t2, col)), col
    big = write_csv(orc.grid2d_rows(40), tmp_path / "lazy.csv")
    net = read_table_netlist(big)
    assert net.nodenum._dict is None and net.degrees._dict is None       # nothing built yet
    assert len(net.nodenum) == net.nums["kcl"] == 40 * 40 - 1
    assert net.ground == "g" and "g" not in net.nodenum and net.nodenum["1"] >= 0
    assert net.nodenum._dict is not None


def test_ground_is_first_node_of_largest_degree_without_g(tmp_path):
    rows = [["r1", "R", "1", "a", "b"], ["r2", "R", "1", "c", "b"], ["r3", "R", "1", "c", "d"],
            ["r4", "R", "1", "d", "a"]] + [[f"s{k}", "R", "1", f"x{k}", f"x{k + 1}"] for k in range(300)]
    slow, fast = same_numbering(write_csv(rows, tmp_path / "nog.csv"))
    assert fast.ground == slow.ground == "a"


@pytest.mark.parametrize("source", ["doc", "c3", "grid_generator", "dict_netlist"])
def test_binary_round_trip(source, tmp_path):
    """save_table_netlist / load_table_netlist: same table, numbering, names and currents."""
    from nodal_b200.ingest import load_table_netlist, save_table_netlist
    if source == "doc":
        net = read_table_netlist(write_csv(DOC["test_1.csv"]["rows"], tmp_path / "t.csv"))
    elif source == "c3":
        net = read_table_netlist(write_csv(gen.random_opamp_network_rows(M=150, P=12, S=10, V=5, seed=2),
                                           tmp_path / "c3.csv"))
    elif source == "grid_generator":
        net = gen.grid2d(25)
    else:
        net = n.Netlist(write_csv(DOC["netlist.csv"]["rows"], tmp_path / "n.csv"))
    path = tmp_path / "net.npz"
    save_table_netlist(net, path)
    back = load_table_netlist(path)
    t1, c1 = net.table_and_currents()
    t2, c2 = back.table_and_currents()
    for col in ("type", "value", "a", "b", "c", "d", "drv", "branch"):
        assert np.array_equal(getattr(t1, col), getattr(t2, col)), col
    assert (t1.kcl, t1.be) == (t2.kcl, t2.be) and list(c1) == list(c2)
    assert back.ground == net.ground
    assert dict(back.nodenum) == dict(net.nodenum) and list(back.nodenum) == list(net.nodenum)
    assert back.anomnum == dict(net.anomnum)
    assert list(back.component_keys) == list(net.component_keys)
    assert back.nums["kcl"] == net.nums["kcl"] and back.nums["components"] == len(t1)


def test_convert_and_load_through_the_cli(tmp_path, capsys):
    from nodal_b200 import cli, ingest
    src = write_csv(orc.grid2d_rows(12), tmp_path / "g.csv")
    dst = str(tmp_path / "g.npz")
    ingest.main([src, dst])
    assert "ground 'g'" in capsys.readouterr().out
    net = cli.load_netlist_or_exit(dst)
    ref = n.Netlist(src)
    assert dict(net.nodenum) == ref.nodenum and net.ground == ref.ground
    assert net.is_resistive()


def test_duplicate_component_names_fall_back_to_the_row_by_row_netlist(tmp_path):
    """The reference keeps one record per name (last row wins, nodal/nodal.py:243) and stamps it
    once per occurrence.  The table ingest refuses such files and the CLI loader falls back to the
    row-by-row Netlist, whose table reproduces that: same file, same answer at any size."""
    from nodal_b200 import cli
    from nodal_b200.ingest import DuplicateNameError
    rows = [["r1", "R", "10", "a", "g"], ["r1", "R", "20", "b", "a"], ["a1", "A", "1", "b", "g"]]
    path = write_csv(rows, tmp_path / "dup.csv")
    with pytest.raises(DuplicateNameError):
        read_table_netlist(path)
    slow = n.Netlist(path)
    t = slow.table()
    assert list(t.value) == [20.0, 20.0, 1.0]            # the last definition, once per occurrence
    ref = orc.Netlist(rows) if hasattr(orc, "Netlist") else None
    if ref is not None and hasattr(ref, "components"):
        assert ref.components["r1"].value == 20.0
    # a large file takes the fast path unless it has duplicates
    big = rows + [[f"x{k}", "R", "1", f"n{k}", "g"] for k in range(70000)]
    big_path = write_csv(big, tmp_path / "dup_big.csv")
    import os
    assert os.path.getsize(big_path) >= cli.FAST_INGEST_BYTES
    net = cli.load_netlist_or_exit(big_path)
    assert isinstance(net, n.Netlist) and list(net.table().value[:3]) == [20.0, 20.0, 1.0]
