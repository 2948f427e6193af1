"""Import shim: the package directory is named `halo2-lasso_b200` (not a Python identifier), so
`import halo2_lasso_b200` loads it from that directory under this importable name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "halo2-lasso_b200")
_spec = importlib.util.spec_from_file_location("halo2_lasso_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["halo2_lasso_b200"] = _mod
_spec.loader.exec_module(_mod)
