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


def _load():
    try:
        import pynbody  # noqa: F401
        from pynbody import config as ref_config
        return ref_config
    except Exception:
        return _defaults


config = _load()
