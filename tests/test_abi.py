"""The C-ABI library loads and exports every symbol include/igneous_b200.h
declares (no compute calls: runs without a GPU)."""
import os
import re

from igneous_b200 import _shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
  text = open(os.path.join(ROOT, "include", "igneous_b200.h")).read()
  return sorted(set(re.findall(r"IGN_API\s+[\w\s\*]+?\b(ign_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
  lib = _shim.load()
  names = declared_symbols()
  assert len(names) >= 45
  missing = [n for n in names if not hasattr(lib, n)]
  assert not missing, missing


def test_version_and_error_string():
  lib = _shim.load()
  assert lib.ign_version() >= 100
  assert isinstance(lib.ign_last_error(), bytes)


def test_no_oracle_import_in_product():
  # the product must never route through the CPU oracle
  pkg = os.path.join(ROOT, "igneous_b200")
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h")):
        src = open(os.path.join(dirpath, f), errors="replace").read()
        assert "import oracle" not in src and "from oracle" not in src, f
        assert "liboracle" not in src, f
