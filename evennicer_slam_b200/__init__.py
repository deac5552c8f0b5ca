"""Importable alias for the ``evennicer-slam_b200/`` package directory.

The repo layout names the package ``evennicer-slam_b200`` (with a hyphen), which is
not a valid Python identifier.  This shim makes ``import evennicer_slam_b200``
resolve to that directory: it points ``__path__`` at it and executes its
``__init__.py`` in this module's namespace.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "evennicer-slam_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py"), "r") as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
