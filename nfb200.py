"""Import alias: `import nfb200` loads the package in ./normalizing-flows-study_b200/ (whose directory name is
not a valid Python identifier) under the module name `nfb200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "normalizing-flows-study_b200")
_spec = importlib.util.spec_from_file_location("nfb200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["nfb200"] = _mod
_spec.loader.exec_module(_mod)
