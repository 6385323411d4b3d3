"""Import shim: the package directory is named ``yolo-pose-cpp_b200`` (not a valid Python
identifier), so it is loaded by path and re-exported under the name ``posebyte_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "yolo-pose-cpp_b200")
_spec = importlib.util.spec_from_file_location("_posebyte_b200_pkg", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["_posebyte_b200_pkg"] = _mod
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
