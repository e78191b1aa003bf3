"""Importable alias of the product package.

The product package lives in the directory `async-ev-cnn_b200/` (the name the build contract
fixes); a hyphen is not a valid Python identifier, so this shim makes the same code importable as
`async_ev_cnn_b200` by pointing the package search path at that directory.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "async-ev-cnn_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
