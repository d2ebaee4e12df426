"""Configuration keys the SPH path reads.

The reference reads an INI cascade (pynbody/configuration.py:13-106,
pynbody/default_config.ini:28-43,70-90).  Only the keys this path consults are
kept here, with the reference's defaults; when pynbody itself is importable
its live ``config`` object is used instead so user settings carry over.
"""
from __future__ import annotations

import os

_defaults = {
    "number_of_threads": os.cpu_count() or 1,   # [general] number_of_threads=-1 -> cpu count (no-op on GPU)
    "image-default-resolution": 1000,
    "image-default-nside": 64,
    "sph": {
        "smooth-particles": 32,
        "tree-leafsize": 16,
        "kernel": "CubicSplineKernel",
        "threaded-image": True,
        "approximate-fast-images": True,
    },
}


class _LiveConfig:
    """Mapping that resolves every lookup at call time: pynbody's own ``config`` once pynbody has been imported
    (by the user, or by ``install()``), the defaults above otherwise.  Resolving lazily matters: this module is
    usually imported before pynbody is."""

    @staticmethod
    def _target():
        import sys
        ref = sys.modules.get("pynbody")
        cfg = getattr(ref, "config", None) if ref is not None else None
        return cfg if cfg is not None else _defaults

    def __getitem__(self, key):
        return self._target()[key]

    def __setitem__(self, key, value):
        self._target()[key] = value

    def __contains__(self, key):
        return key in self._target()

    def get(self, key, default=None):
        t = self._target()
        return t.get(key, default) if hasattr(t, "get") else (t[key] if key in t else default)


config = _LiveConfig()
