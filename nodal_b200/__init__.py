"""nodal_b200 -- B200-native MNA assembly and node-voltage solve behind nodal's Python surface.

Drop-in for the hot path of EnricoMiccoli/nodal 1.3.0 (`import nodal_b200 as nodal`,
or `nodal_b200.install_as_nodal()` to register the `nodal` module names).
"""
__version__ = "1.3.0+b200.r1"
from .nodal import *  # noqa: F401,F403  (the reference re-exports nodal.nodal wholesale, __init__.py:3)
from .nodal import Circuit, Component, Netlist, Solution, UnconnectedCircuitError  # noqa: F401
from .table import ComponentTable  # noqa: F401


def install_as_nodal():
    """Register this package under the reference's module names (nodal, nodal.nodal,
    nodal.models, nodal.constants, nodal.equiv, nodal.solver) in sys.modules."""
    import importlib
    import sys
    pkg = sys.modules[__name__]
    sys.modules["nodal"] = pkg
    for sub in ("nodal", "models", "constants", "equiv", "solver"):
        sys.modules[f"nodal.{sub}"] = importlib.import_module(f"{__name__}.{sub}")
    return pkg
