"""Import shim: the package directory is named ``adaptive-volume-rendering_b200`` (not a
valid Python identifier), so ``import avr_b200`` loads it under this name."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "adaptive-volume-rendering_b200")
_spec = _ilu.spec_from_file_location(
    "avr_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = _ilu.module_from_spec(_spec)
_sys.modules["avr_b200"] = _mod
_spec.loader.exec_module(_mod)
