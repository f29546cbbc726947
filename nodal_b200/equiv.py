"""`nodal-resistance FILE [-s]` and equivalent_resistance() (mirror of reference nodal/equiv.py)."""
import argparse
from copy import deepcopy

import nodal_b200 as n

parser = argparse.ArgumentParser(
    description="Calculate equivalent resistance using nodal analysis"
    "\n"
    "Label nodes as '1' and 'g' to mark where to connect to the network.")
parser.add_argument("netlist_path", metavar="FILE", help="csv file describing the resistive network")
parser.add_argument("-s", "--sparse", action="store_true", help="use a sparse matrix")


def check_resistive(netlist):
    """True when every component of the netlist is a resistor (equiv.py:22-28)."""
    return netlist.is_resistive()


def equivalent_resistance(netlist, a, b, sparse=False, **options):
    """Equivalent resistance seen through nodes a and b (equiv.py:31-61).

    A 1 A source is connected from b to a, the circuit is solved on the GPU and
    R = e(a) - e(b).  Raises ValueError for non-resistive netlists and KeyError
    for unknown nodes.  `options` are forwarded to Circuit (rtol, maxit, ...).
    """
    if not check_resistive(netlist):
        raise ValueError("Network is not resistive")
    for node in (a, b):
        if node not in netlist.nodenum and node != netlist.ground:
            raise KeyError(f"Node `{node}` not found in netlist")
    probe = deepcopy(netlist)
    probe.process_component(["a1", "A", "1", a, b])
    solution = n.Circuit(probe, sparse=sparse, **options).solve()
    e = [0, 0]
    for k, node in enumerate((a, b)):
        if node != "g":          # literal "g", as the reference (equiv.py:57)
            e[k] = solution.result[solution.nodenum[node]]
    equivalent_resistance.last_stats = solution.stats
    return e[0] - e[1]


def main(argv=None):
    args = parser.parse_args(argv)
    try:
        netlist = n.Netlist(args.netlist_path)
    except FileNotFoundError:
        exit(1)
    try:
        r = equivalent_resistance(netlist, "1", "g", sparse=args.sparse)
    except ValueError:
        print("Invalid netlist\n")
        print("Resistors are the only component allowed in the circuit")
        exit(1)
    except KeyError as e:
        print("Invalid netlist\n")
        print(e.args[0])
        exit(1)
    print(f"R = {r}")


if __name__ == "__main__":
    main()
