"""Route an importable pynbody through the GPU path.

``install()`` swaps exactly the native modules of the reference that sit on the SPH smoothing path
-- ``pynbody.kdtree.kdmain`` (C++), the three functions of ``pynbody.sph._render`` (Cython) and
``pynbody.filt.geometry_selection.selection`` (Cython/C++) -- plus the ``KDTree`` host class, and nothing
else.  pynbody's derived arrays (``sim['smooth']``, ``sim['rho']``, ``v_mean`` ...), ``build_tree``,
filters and the render pipeline keep calling the same names (pynbody/sph/__init__.py:40-94,
derived.py:126-154, snapshot/simsnap.py:1729-1750, sph/renderers.py:690-697, 711-718, 752-766,
filt/__init__.py:232,257,299) and end up on the GPU.
"""
from __future__ import annotations

import gc
import importlib

_saved = {}


def _optional(name):
    try:
        return importlib.import_module(name)
    except ImportError:
        return None


def install(pynbody=None, replace_kdtree_class=True, healpix=True, filters=True):
    """Patch ``pynbody`` (imported if not given) in place; returns the patched module."""
    from .filt import geometry_selection
    from .kdtree import KDTree, kdmain
    from .sph import _render
    if pynbody is None:
        pynbody = importlib.import_module("pynbody")
    kdt = importlib.import_module(pynbody.__name__ + ".kdtree")
    rnd = importlib.import_module(pynbody.__name__ + ".sph._render")
    sel = _optional(pynbody.__name__ + ".filt.geometry_selection") if filters else None
    if not _saved:
        _saved.update(kdmain=getattr(kdt, "kdmain", None), KDTree=getattr(kdt, "KDTree", None),
                      render_image=getattr(rnd, "render_image", None), to_3d_grid=getattr(rnd, "to_3d_grid", None),
                      healpix=getattr(rnd, "render_spherical_image_core", None),
                      selection=getattr(sel, "selection", None) if sel else None)
    # trees built so far belong to the extension that built them: let unreferenced ones die first, and route the
    # ``free`` of any survivor (KDTree.__del__ resolves ``kdmain`` at call time) back to the reference module
    gc.collect()
    if _saved["kdmain"] is not None and _saved["kdmain"] is not kdmain:
        kdmain._foreign_free = getattr(_saved["kdmain"], "free", None)
    kdt.kdmain = kdmain                      # the reference KDTree class resolves `kdmain.*` at call time
    if replace_kdtree_class:
        kdt.KDTree = KDTree                  # lazily materialised kdnodes / particle_offsets
    rnd.render_image = _render.render_image  # looked up as `_render.render_image` (renderers.py:691)
    rnd.to_3d_grid = _render.to_3d_grid      # (renderers.py:712)
    if healpix:
        rnd.render_spherical_image_core = _render.render_spherical_image_core      # (renderers.py:752-766)
    if sel is not None:
        sel.selection = geometry_selection.selection  # `geometry_selection.selection(...)` (filt/__init__.py:232)
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
    sel = _optional(pynbody.__name__ + ".filt.geometry_selection")
    gc.collect()                             # GPU trees nobody references any more are freed by the module that owns them
    kdt.kdmain, kdt.KDTree = _saved["kdmain"], _saved["KDTree"]
    rnd.render_image, rnd.to_3d_grid = _saved["render_image"], _saved["to_3d_grid"]
    if _saved["healpix"] is not None:
        rnd.render_spherical_image_core = _saved["healpix"]
    if sel is not None and _saved["selection"] is not None:
        sel.selection = _saved["selection"]
    _saved.clear()
