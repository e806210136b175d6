"""Importable alias of the `a.i.gar_b200/` package directory.

The framework's package directory is literally named ``a.i.gar_b200`` (the dots make it
unimportable with a plain ``import`` statement), so this shim points its ``__path__`` at that
directory: ``import aigar_b200.layout`` loads ``a.i.gar_b200/layout.py``.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "a.i.gar_b200")
__path__.insert(0, _REAL)
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
