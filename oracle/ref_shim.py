"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY — used by ``oracle/make_golden.py`` and by
``tests/test_oracle_vs_reference.py`` to pin ``oracle/avr_oracle.py``.  The
reference does not exist on the GPU box, so nothing marked ``gpu``, nor
``smoke()`` nor ``bench.py`` touches this module.

``/root/reference/utils.py:7-31`` star-imports eight third-party packages that
are not installed here (matplotlib, lpips, gdown, h5py, imageio, skimage,
dotmap, pyhocon); none is used by the sampling / compositing path, so empty
stand-ins are registered before the import.  No reference source is copied.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AVR_REFERENCE_ROOT", "/root/reference")

_MISSING = (
    "matplotlib", "matplotlib.pyplot", "lpips", "gdown", "h5py", "imageio",
    "skimage", "skimage.metrics", "skimage.transform", "dotmap", "pyhocon", "configargparse",
)


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "renderers.py"))


class _Anything(types.ModuleType):
    """Module whose every attribute is a harmless placeholder."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _placeholder


def _placeholder(*_a, **_k):  # pragma: no cover - never reached on the hot path
    raise RuntimeError("stubbed third-party symbol used; not part of the hot path")


def load():
    """Return the reference's ``renderers`` module (imports utils/models too)."""
    if not available():
        raise FileNotFoundError(f"reference not present at {REFERENCE_ROOT}")
    for name in _MISSING:
        try:
            importlib.import_module(name)
        except Exception:
            mod = _Anything(name)
            mod.__path__ = []  # looks like a package so 'from x.y import z' resolves
            sys.modules[name] = mod
    # 'from dotmap import DotMap', 'from pyhocon import ConfigFactory' etc. resolve to placeholders
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("renderers")


class Conf(dict):
    """Minimal stand-in for a pyhocon ConfigTree (get_int/get_float/... with defaults)."""

    def _get(self, key, default=None):
        return self[key] if key in self else default

    get_int = get_float = get_bool = get_string = get_list = _get

    def get_config(self, key, default=None):
        v = self._get(key, default)
        return Conf(v) if isinstance(v, dict) else v

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        return Conf(v) if isinstance(v, dict) and not isinstance(v, Conf) else v
