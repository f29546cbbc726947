"""`nodal-solver FILE [-s]`: print every node potential and branch current of a netlist.

Mirror of the reference's nodal/solver.py (same arguments, same exit codes, same output); the
solve itself runs on the GPU through nodal_b200.Circuit.
"""
import sys

import nodal_b200 as n
from nodal_b200.cli import circuit_options, load_netlist_or_exit, make_parser

parser = make_parser("Solve electrical circuits using nodal analysis", "csv file describing the netlist")


def solve_file(path, sparse=False, **options):
    """Returns the printable Solution, or None when the circuit has floating nodes."""
    netlist = load_netlist_or_exit(path)
    try:
        return n.Circuit(netlist, sparse=sparse, **options).solve()
    except n.UnconnectedCircuitError:
        return None


def main(argv=None):
    options = parser.parse_args(argv)
    solution = solve_file(options.netlist_path, sparse=options.sparse, **circuit_options(options))
    if solution is None:
        sys.exit(1)
    print(solution)


if __name__ == "__main__":
    main()
