"""Importable alias of the `mp-block-preconditioners_b200/` package directory (a hyphen cannot
appear in a Python module name).  All code lives there; this file only redirects the package path."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                          "mp-block-preconditioners_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
