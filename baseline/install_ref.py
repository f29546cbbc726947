#!/usr/bin/env python
"""Install the UNMODIFIED reference (EnricoMiccoli/nodal) into baseline/_ref for the CPU arm of
bench.py (`--impl reference`, `cpu_baseline.kind == "reference"`).

    python baseline/install_ref.py            # build container only (/root/reference is not on the GPU box)

The reference builds with flit (pyproject.toml:1-3), which is not in this image and cannot be
fetched (no network), so `pip install /root/reference` fails in the build step.  The package is
pure Python: what flit would install is the `nodal/` directory as it is.  This script therefore
installs from a scratch copy under /tmp whose ONLY change is the build metadata (a setuptools
pyproject in place of the flit one); every file of the `nodal` package lands in baseline/_ref
byte for byte (checked below).  baseline/_ref is git-ignored (no reference source enters the
history) and travels to the GPU box with the gpurun snapshot.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NODAL_REFERENCE", "/root/reference")
TARGET = os.path.join(ROOT, "baseline", "_ref")

PYPROJECT = """[build-system]
requires = ["setuptools>=61"]
build-backend = "setuptools.build_meta"

[project]
name = "nodal"
version = "1.3.0"
requires-python = ">=3"

[project.scripts]
nodal-solver = "nodal.solver:main"
nodal-resistance = "nodal.equiv:main"

[tool.setuptools]
packages = ["nodal"]
"""


def installed():
    return os.path.isfile(os.path.join(TARGET, "nodal", "nodal.py"))


def main():
    if not os.path.isdir(os.path.join(REF, "nodal")):
        print(f"reference not found at {REF}: nothing installed", file=sys.stderr)
        return 1
    tmp = tempfile.mkdtemp(prefix="nodal_ref_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        with open(os.path.join(src, "pyproject.toml"), "w") as fh:
            fh.write(PYPROJECT)
        shutil.rmtree(TARGET, ignore_errors=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    cmp = filecmp.dircmp(os.path.join(REF, "nodal"), os.path.join(TARGET, "nodal"), ignore=["__pycache__"])
    assert not cmp.diff_files and not cmp.left_only, (cmp.diff_files, cmp.left_only)
    print(f"reference installed unmodified into {TARGET}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
