"""Importable alias of the package directory `bts-fully-tf_b200/` (a hyphen cannot appear in a Python
module name).  All code lives there; this file only redirects the package path."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "bts-fully-tf_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
