"""Import shim: the product package lives in the directory `gridapembedded.jl_b200/` (a name Python cannot
import directly because of the dot); `import gecut` loads it under the module name `gecut`."""
import importlib.util
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_root, "gridapembedded.jl_b200")
_spec = importlib.util.spec_from_file_location("gecut", os.path.join(_pkg, "__init__.py"),
                                               submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gecut"] = _mod
_spec.loader.exec_module(_mod)
