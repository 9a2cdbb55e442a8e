"""Import shim: the package directory is named `indexed-merkle-tree-halo2_b200` (not a Python identifier), so this
module loads it under the importable name `imt_b200`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "indexed-merkle-tree-halo2_b200")
_spec = importlib.util.spec_from_file_location("imt_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["imt_b200"] = _mod
_spec.loader.exec_module(_mod)
