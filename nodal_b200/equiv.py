"""`nodal-resistance FILE [-s]` and equivalent_resistance().

Mirror of the reference's nodal/equiv.py: same function signature, same errors, same command
line; the probe-source solve runs on the GPU.
"""
import copy
import sys

import nodal_b200 as n
from nodal_b200.cli import circuit_options, load_netlist_or_exit, make_parser
from nodal_b200.generators import TableNetlist

parser = make_parser(
    "Calculate equivalent resistance using nodal analysis\n"
    "Label nodes as '1' and 'g' to mark where to connect to the network.",
    "csv file describing the resistive network")

PROBE = ("a1", "A", "1")          # name, type, value of the 1 A source the method inserts (equiv.py:51)


def check_resistive(netlist):
    """True when the netlist contains resistors only (reference equiv.py:22-28)."""
    return netlist.is_resistive()


def equivalent_resistance(netlist, a, b, sparse=False, **options):
    """Resistance seen between nodes `a` and `b` (reference equiv.py:31-61).

    Method: connect a 1 A current source from b to a, solve, return e(a) - e(b).
    Raises ValueError if anything but resistors is present and KeyError if a or b is not a
    node of the netlist.  Extra keyword `options` go to Circuit (rtol, maxit, ...).
    """
    if not check_resistive(netlist):
        raise ValueError("Network is not resistive")
    missing = [node for node in (a, b) if node != netlist.ground and node not in netlist.nodenum]
    if missing:
        raise KeyError(f"Node `{missing[0]}` not found in netlist")

    if isinstance(netlist, TableNetlist):
        # Same linear system without touching the table: the 1 A source only contributes the
        # right-hand side entries +1 at a and -1 at b (models.py:27-32), which
        # Circuit.port_resistances writes directly.  Appending a row would copy, re-scan and
        # re-upload the whole table (1.1 GB at 16.7 M nodes) for every call.
        for node in (a, b):
            if node == netlist.ground and node != "g":
                raise KeyError(node)                     # reference: nodenum[ground label], equiv.py:57-59
        circuit = n.Circuit(netlist, sparse=sparse, **options)
        (resistance,) = circuit.port_resistances([(a, b)])
        equivalent_resistance.last_stats = circuit.stats[0]
        return resistance

    probed = copy.deepcopy(netlist)                      # the caller's netlist stays untouched
    probed.process_component([*PROBE, a, b])
    solution = n.Circuit(probed, sparse=sparse, **options).solve()
    equivalent_resistance.last_stats = solution.stats

    def potential(node):
        # the reference tests against the literal "g", not netlist.ground (equiv.py:57)
        return 0 if node == "g" else solution.result[solution.nodenum[node]]

    return potential(a) - potential(b)


def equivalent_resistances(netlist, pairs, sparse=True, **options):
    """Many-port form: [R(a, b) for (a, b) in pairs] with the matrix assembled once and, with
    precond="amg", one AMG hierarchy shared by all right-hand sides.
    Each value equals equivalent_resistance(netlist, a, b) to the solver tolerance; same errors."""
    if not check_resistive(netlist):
        raise ValueError("Network is not resistive")
    pairs = [tuple(p) for p in pairs]
    for pair in pairs:
        for node in pair:
            if node != netlist.ground and node not in netlist.nodenum:
                raise KeyError(f"Node `{node}` not found in netlist")
    circuit = n.Circuit(netlist, sparse=sparse, **options)
    values = circuit.port_resistances(pairs)
    equivalent_resistances.last_stats = circuit.stats
    return values


def main(argv=None):
    options = parser.parse_args(argv)
    netlist = load_netlist_or_exit(options.netlist_path)
    try:
        r = equivalent_resistance(netlist, "1", "g", sparse=options.sparse, **circuit_options(options))
    except n.UnconnectedCircuitError:
        sys.exit(1)
    except ValueError:
        print("Invalid netlist\n")
        print("Resistors are the only component allowed in the circuit")
        sys.exit(1)
    except KeyError as err:
        print("Invalid netlist\n")
        print(err.args[0])
        sys.exit(1)
    print(f"R = {r}")


if __name__ == "__main__":
    main()
