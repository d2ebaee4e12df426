"""Route an importable pynbody through the GPU path.

``install()`` swaps exactly the two native modules of the reference that sit on the SPH smoothing path
-- ``pynbody.kdtree.kdmain`` (C++) and the two functions of ``pynbody.sph._render`` (Cython) -- plus the
``KDTree`` host class, and nothing else.  pynbody's derived arrays (``sim['smooth']``, ``sim['rho']``,
``v_mean`` ...), ``build_tree``, filters and the render pipeline keep calling the same names
(pynbody/sph/__init__.py:40-94, derived.py:126-154, snapshot/simsnap.py:1729-1750,
sph/renderers.py:690-697, 711-718) and end up on the GPU.
"""
from __future__ import annotations

import importlib

_saved = {}


def install(pynbody=None, replace_kdtree_class=True):
    """Patch ``pynbody`` (imported if not given) in place; returns the patched module."""
    from .kdtree import KDTree, kdmain
    from .sph import _render
    if pynbody is None:
        pynbody = importlib.import_module("pynbody")
    kdt = importlib.import_module(pynbody.__name__ + ".kdtree")
    rnd = importlib.import_module(pynbody.__name__ + ".sph._render")
    if not _saved:
        _saved.update(kdmain=getattr(kdt, "kdmain", None), KDTree=getattr(kdt, "KDTree", None),
                      render_image=getattr(rnd, "render_image", None), to_3d_grid=getattr(rnd, "to_3d_grid", None))
    kdt.kdmain = kdmain                      # the reference KDTree class resolves `kdmain.*` at call time
    if replace_kdtree_class:
        kdt.KDTree = KDTree                  # lazily materialised kdnodes / particle_offsets
    rnd.render_image = _render.render_image  # looked up as `_render.render_image` (renderers.py:691)
    rnd.to_3d_grid = _render.to_3d_grid      # (renderers.py:712)
    # the exact renderer replaces the multi-resolution approximation (renderers.py:576-611)
    try:
        pynbody.config["sph"]["approximate-fast-images"] = False
        pynbody.config["sph"]["threaded-image"] = False
    except Exception:
        pass
    return pynbody


def uninstall(pynbody=None):
    if not _saved:
        return
    if pynbody is None:
        pynbody = importlib.import_module("pynbody")
    kdt = importlib.import_module(pynbody.__name__ + ".kdtree")
    rnd = importlib.import_module(pynbody.__name__ + ".sph._render")
    kdt.kdmain, kdt.KDTree = _saved["kdmain"], _saved["KDTree"]
    rnd.render_image, rnd.to_3d_grid = _saved["render_image"], _saved["to_3d_grid"]
    _saved.clear()
