"""Importable alias of the package directory `realtime-vulkan-hair_b200/` (a hyphen is not a
valid Python identifier).  All code lives there; this module only redirects the import."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "realtime-vulkan-hair_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
