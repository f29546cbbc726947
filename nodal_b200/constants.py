"""Constants shared between modules (mirror of reference nodal/constants.py:1-35).

The numeric values are part of the parity contract: the CSV column map feeds
``Netlist``; the type lists decide which components get a branch unknown; the
op-amp macro-model values (RI / RO / GAIN) end up, through ``str()`` and
``float()``, in the component table the stamp kernel reads.
"""

# CSV columns (nodal/constants.py:4-12)
NCOL = 0  # component name
TCOL = 1  # component type
VCOL = 2  # component value
ACOL = 3  # first lead (currents enter here)
BCOL = 4  # second lead
CCOL = 5  # first control node (dependent sources)
DCOL = 6  # second control node
PCOL = 7  # driving component (current-controlled sources)

# Type lists (nodal/constants.py:15-18)
NODE_TYPES_CC = ["CCCS", "CCVS"]
NODE_TYPES_DEP = ["VCVS", "VCCS"] + NODE_TYPES_CC
NODE_TYPES_ANOM = ["E"] + NODE_TYPES_DEP
NODE_TYPES = ["A", "R"] + NODE_TYPES_ANOM + ["OPAMP", "OPMODEL"]

# Number of csv fields per type (nodal/constants.py:20-30)
NODE_ARGS_NUMBER = {
    "OPAMP": 7, "OPMODEL": 7,
    "R": 5, "A": 5, "E": 5,
    "VCCS": 7, "VCVS": 7,
    "CCCS": 8, "CCVS": 8,
}

# Op-amp macro model (nodal/constants.py:33-35)
OPMODEL_RI = 1e7   # ohm
OPMODEL_RO = 10    # ohm
OPMODEL_GAIN = 1e5

# ---- B200 build: type codes of the struct-of-arrays component table (device side
# mirror is csrc/common.cuh enum CompType; keep in sync).
T_R, T_A, T_E, T_VCVS, T_VCCS, T_CCVS, T_CCCS = range(7)
TYPE_CODE = {"R": T_R, "A": T_A, "E": T_E, "VCVS": T_VCVS, "VCCS": T_VCCS,
             "CCVS": T_CCVS, "CCCS": T_CCCS}
TYPE_NAME = {v: k for k, v in TYPE_CODE.items()}
GROUND = -1   # lead / control index meaning "ground node" (row dropped)
UNUSED = -2   # control index of a component type that has no control nodes
