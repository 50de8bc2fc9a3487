"""`import nsb200` -> the package in navier-stokes_equations_b200/ (its directory name is not a Python identifier,
so it is loaded by path).  bench.py, __graft_entry__.py, tools/ and tests/ all go through this module."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "navier-stokes_equations_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
