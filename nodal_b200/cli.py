"""Shared pieces of the two console entry points (`nodal-solver`, `nodal-resistance`)."""
import argparse
import sys

import nodal_b200 as n


def make_parser(description, file_help):
    """FILE plus -s/--sparse: the whole command line of both tools (reference
    nodal/solver.py:7-13, nodal/equiv.py:11-19)."""
    parser = argparse.ArgumentParser(description=description)
    parser.add_argument("netlist_path", metavar="FILE", help=file_help)
    parser.add_argument("-s", "--sparse", action="store_true", help="use a sparse matrix")
    return parser


def load_netlist_or_exit(path):
    """Exit status 1 when the file does not exist (the error itself is logged by Netlist)."""
    try:
        return n.Netlist(path)
    except FileNotFoundError:
        sys.exit(1)
