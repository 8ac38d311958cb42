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


def test_fails_loudly_without_a_gpu():
  """No CPU fallback: on a box without a CUDA device every compute entry point
  must raise instead of silently computing on the host."""
  import numpy as np
  import pytest
  if _shim.device_count() > 0:
    pytest.skip("a CUDA device is visible")
  with pytest.raises(_shim.IgneousB200Error):
    _shim.Context(0)
  from igneous_b200 import tinybrain, cc3d, zmesh, fastremap
  img = np.zeros((8, 8, 8), dtype=np.uint32)
  for call in (lambda: tinybrain.downsample_segmentation(img, (2, 2, 1)),
               lambda: cc3d.connected_components(img, connectivity=6),
               lambda: fastremap.renumber(img + 1),
               lambda: zmesh.Mesher((1, 1, 1)).mesh(img)):
    with pytest.raises(_shim.IgneousB200Error):
      call()


def test_missing_library_is_an_error(monkeypatch, tmp_path):
  import importlib
  import pytest
  monkeypatch.setenv("IGNEOUS_B200_LIB", str(tmp_path / "nope.so"))
  monkeypatch.setattr(_shim, "_lib", None)
  with pytest.raises(_shim.NativeLibraryMissing):
    _shim.load()
  monkeypatch.delenv("IGNEOUS_B200_LIB")
  monkeypatch.setattr(_shim, "_lib", None)
  assert _shim.load() is not None
