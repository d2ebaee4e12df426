"""Array-level front end of the SPH path: what ``pynbody.sph`` does for a snapshot, for plain arrays.

pynbody's snapshot layer (``sim['smooth']``, ``sim['rho']``, ``pynbody.sph.render_image(sim, ...)``) reaches
the native code through a little host glue: neighbour counts and kernel from the config
(pynbody/sph/__init__.py:40-94), image geometry, periodic tile offsets and the projected-kernel switch
(pynbody/sph/renderers.py:45-104, 616-629, 650-718, 921-938).  This module restates that glue for numpy /
torch arrays so the path is usable (and testable) without a pynbody installation; with pynbody installed,
``pynbody_b200.install.install()`` makes pynbody's own glue call the same native entry points.

Units never reach the native layer: arrays are taken as plain numbers.
"""
from __future__ import annotations

import numpy as np

from ..configuration import config
from . import _render, kernels


def _default_nn():
    try:
        return int(config["sph"]["smooth-particles"])
    except Exception:
        return 32


def _tree(pos, mass, boxsize, kernel):
    from ..kdtree import KDTree
    try:
        leafsize = int(config["sph"]["tree-leafsize"])
    except Exception:
        leafsize = 16
    kd = KDTree(pos, mass, leafsize=leafsize, boxsize=boxsize)         # simsnap.py:1729-1750
    kd.set_kernel(kernel)
    return kd


def _empty_like(a, n):
    from .. import _lib
    if _lib._is_torch(a):
        import torch
        return torch.empty(n, dtype=a.dtype, device=a.device)
    return np.empty(n, dtype=a.dtype)


def smooth(pos, mass, nn=None, boxsize=None, kernel=None, kdtree=None):
    """Smoothing lengths h = half the distance to the nn-th neighbour (pynbody/sph/__init__.py:40-57)."""
    nn = _default_nn() if nn is None else int(nn)
    kd = kdtree or _tree(pos, mass, boxsize, kernel)
    sm = _empty_like(mass, len(pos))
    kd.set_array_ref("smooth", sm)
    kd.populate("hsm", nn)
    return sm


def rho(pos, mass, nn=None, boxsize=None, kernel=None, smooth_array=None, kdtree=None):
    """SPH density (pynbody/sph/__init__.py:71-94); returns ``(rho, smooth)``."""
    nn = _default_nn() if nn is None else int(nn)
    kd = kdtree or _tree(pos, mass, boxsize, kernel)
    sm = smooth(pos, mass, nn, boxsize, kernel, kd) if smooth_array is None else smooth_array
    out = _empty_like(mass, len(pos))
    kd.set_array_ref("smooth", sm)
    kd.set_array_ref("mass", mass)
    kd.set_array_ref("rho", out)
    kd.populate("rho", nn)
    return out, sm


# ---- image geometry (pynbody/sph/renderers.py:45-104) ------------------------------------------------
class ImageGeometry:
    """Frame of an image: x/y extent centred on the origin, optional depth range, slice plane, camera."""

    def __init__(self, width=200.0, resolution=None):
        self.z1, self.z2 = -np.inf, np.inf
        self.z_plane = 0.0
        self.z_camera = None
        res = resolution if resolution is not None else config["image-default-resolution"]
        self.nx = self.ny = self.nz = int(res)
        self.set_width(width)

    @property
    def width(self):
        return self.x2 - self.x1

    def set_width(self, width):
        self.x2 = self.y2 = width / 2
        self.x1 = self.y1 = -width / 2

    def set_resolution(self, resolution):
        self.nx = self.ny = self.nz = int(resolution)

    def restrict_z_range(self):
        self.z1, self.z2 = self.x1, self.x2

    def set_pixels(self, nx=None, ny=None, nz=None):
        """Non-square frames keep square pixels: the y (z) extent scales with ny/nx (renderers.py:931-942)."""
        base = self.nx if nx is None else int(nx)
        if nx is not None:
            self.nx = int(nx)
        if ny is not None:
            self.ny = int(ny)
            self.y1 *= ny / base
            self.y2 *= ny / base
        if nz is not None:
            self.nz = int(nz)
            self.z1 *= nz / base
            self.z2 *= nz / base


def wrapping_repeat_array(x1, x2, boxsize):
    """Periodic tile offsets that make an image of extent [x1,x2] wrap (renderers.py:616-629)."""
    if boxsize:
        num_repeats = int(round((x2 - x1) / (2 * boxsize))) + 1
        return np.linspace(-num_repeats * boxsize, num_repeats * boxsize, num_repeats * 2 + 1)
    return np.array([0.0])


def _columns(pos):
    return pos[:, 0], pos[:, 1], pos[:, 2]


def render_image(pos, mass, rho, smooth, quantity=None, width=None, resolution=None, nx=None, ny=None, kernel=None,
                 projected=True, z_plane=0.0, z_camera=None, smooth_range=(0.0, np.inf), smooth_floor=0.0,
                 restrict_depth=False, boxsize=None, geometry=None):
    """SPH image of ``quantity`` (default: the density) -- ImageRenderer.render + _call_c_renderer
    (renderers.py:650-697).  ``projected=True`` integrates along z with the 2-D kernel; otherwise a slice at
    ``z_plane``.  Always the exact renderer (the reference's ``approximate_fast=False``)."""
    g = geometry or ImageGeometry(width if width is not None else 200.0, resolution)
    if geometry is None:
        g.set_pixels(nx, ny)
        g.z_plane, g.z_camera = z_plane, z_camera
        if restrict_depth:
            g.restrict_z_range()
    k = kernels.create_kernel(kernel)
    if projected:
        k = k.projection()
    q = rho if quantity is None else quantity
    x, y, z = _columns(pos)
    return _render.render_image(g.nx, g.ny, x, y, z, smooth, g.x1, g.x2, g.y1, g.y2, g.z_camera or 0.0, g.z_plane,
                                q, mass, rho, smooth_range[0], smooth_range[1], g.z1, g.z2, smooth_floor, k,
                                wrapping_repeat_array(g.x1, g.x2, boxsize), wrapping_repeat_array(g.y1, g.y2, boxsize))


def render_3d_grid(pos, mass, rho, smooth, quantity=None, width=None, resolution=None, nx=None, ny=None, nz=None,
                   kernel=None, smooth_range=(0.0, np.inf), boxsize=None, geometry=None):
    """SPH interpolation onto a 3-D grid -- Grid3dRenderer (renderers.py:700-718): a cube of side ``width``
    centred on the origin, float32 ``(nx, ny, nz)``."""
    g = geometry or ImageGeometry(width if width is not None else 200.0, resolution)
    if geometry is None:
        g.restrict_z_range()
        g.set_pixels(nx, ny, nz)
    k = kernels.create_kernel(kernel)
    q = rho if quantity is None else quantity
    x, y, z = _columns(pos)
    return _render.to_3d_grid(g.nx, g.ny, g.nz, x, y, z, smooth, g.x1, g.x2, g.y1, g.y2, g.z1, g.z2, q, mass, rho,
                              smooth_range[0], smooth_range[1], k, wrapping_repeat_array(g.x1, g.x2, boxsize),
                              wrapping_repeat_array(g.y1, g.y2, boxsize), wrapping_repeat_array(g.z1, g.z2, boxsize))


to_3d_grid = render_3d_grid
