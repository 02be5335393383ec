"""Import shim: the package directory is named `fdtd-maxwell-microwave-oven_b200` (hyphens, as the
project is called), which Python cannot import by name.  `import fdtd_b200` loads it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fdtd-maxwell-microwave-oven_b200")
_spec = importlib.util.spec_from_file_location("fdtd_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fdtd_b200"] = _mod
_spec.loader.exec_module(_mod)
