"""Importable name of the `mp-block-preconditioners_b200/` package directory (a hyphen cannot appear in a Python module
name).  All code lives there: this module loads that directory's package under this name with the standard importlib
machinery and takes its place in `sys.modules`, so `import mp_block_preconditioners_b200` IS that package."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mp-block-preconditioners_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_real, "__init__.py"), submodule_search_locations=[_real])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
