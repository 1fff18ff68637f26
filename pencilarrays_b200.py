"""Import shim: the package directory is named ``pencilarrays.jl_b200`` (a dot
is not importable), so this module loads it under the name ``pencilarrays_b200``.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pencilarrays.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "pencilarrays_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pencilarrays_b200"] = _mod
_spec.loader.exec_module(_mod)
