"""Import alias for the package directory ``modelpredictivecontrol.jl_b200/`` (named after the
reference, hence not a valid Python identifier): ``import mpc_b200`` loads that package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "modelpredictivecontrol.jl_b200")
_spec = importlib.util.spec_from_file_location("mpc_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mpc_b200"] = _mod
_spec.loader.exec_module(_mod)
