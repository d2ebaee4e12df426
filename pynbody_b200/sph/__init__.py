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


class RenderPipelineLogicError(RuntimeError):
    """Inconsistent combination of projection / weighting / denoising (renderers.py:42-43)."""


def _ones_like(a):
    from .. import _lib
    if _lib._is_torch(a):
        import torch
        return torch.ones_like(a)
    return np.ones_like(np.asarray(a))


def _divide(num, den):
    from .. import _lib
    if _lib._is_torch(num):
        return num / den
    with np.errstate(divide="ignore", invalid="ignore"):
        return num / den


def _multipass(render_one, quantity, projected, weight, denoise):
    """The two-pass decorators of the reference's render pipeline, in the order make_render_pipeline applies
    them (renderers.py:985-992):

    * ``weight`` (True = volume weighting, or an array): projected average
      ``render(q * w, projected) / render(w, projected)`` -- ProjectionAverageImageRenderer (renderers.py:508-556);
    * ``denoise``: ``render(q) / render(ones)`` for slices and grids -- DenoisedImageRenderer (renderers.py:488-506).

    ``render_one(quantity, projected)`` performs one native render."""
    if weight is not None and weight is not False:
        if projected:
            raise RenderPipelineLogicError("Cannot take a projected average of an already projected image.")
        w = _ones_like(quantity) if weight is True else weight
        if len(w) != len(quantity):
            raise ValueError("Weighting array must have the same length as the snapshot")
        if denoise:
            raise RenderPipelineLogicError("Denoising not supported with projected images")
        return _divide(render_one(quantity * w, True), render_one(w, True))
    if denoise:
        if projected:
            raise RenderPipelineLogicError("Denoising not supported with projected images")
        return _divide(render_one(quantity, False), render_one(_ones_like(quantity), False))
    return render_one(quantity, projected)


def render_image(pos, mass, rho, smooth, quantity=None, width=None, resolution=None, nx=None, ny=None, kernel=None,
                 projected=True, z_plane=0.0, z_camera=None, smooth_range=(0.0, np.inf), smooth_floor=0.0,
                 restrict_depth=False, boxsize=None, geometry=None, weight=None, denoise=False):
    """SPH image of ``quantity`` (default: the density) -- ImageRenderer.render + _call_c_renderer
    (renderers.py:650-697).  ``projected=True`` integrates along z with the 2-D kernel; otherwise a slice at
    ``z_plane``.  ``weight`` / ``denoise`` add the reference's two-pass decorators (see :func:`_multipass`;
    ``weight`` needs ``projected=False`` as its passes are projections themselves).  Always the exact renderer
    (the reference's ``approximate_fast=False``)."""
    g = geometry or ImageGeometry(width if width is not None else 200.0, resolution)
    if geometry is None:
        g.set_pixels(nx, ny)
        g.z_plane, g.z_camera = z_plane, z_camera
        if restrict_depth:
            g.restrict_z_range()
    k3 = kernels.create_kernel(kernel)
    k2 = None
    x, y, z = _columns(pos)
    wx, wy = wrapping_repeat_array(g.x1, g.x2, boxsize), wrapping_repeat_array(g.y1, g.y2, boxsize)

    def one(q, proj):
        nonlocal k2
        if proj and k2 is None:
            k2 = k3.projection()
        return _render.render_image(g.nx, g.ny, x, y, z, smooth, g.x1, g.x2, g.y1, g.y2, g.z_camera or 0.0, g.z_plane,
                                    q, mass, rho, smooth_range[0], smooth_range[1], g.z1, g.z2, smooth_floor,
                                    k2 if proj else k3, wx, wy)

    return _multipass(one, rho if quantity is None else quantity, projected, weight, denoise)


def render_healpix(pos, mass, rho, smooth, quantity=None, nside=None, kernel=None, shell_distance=None, weight=None,
                   denoise=False):
    """Healpix (RING) SPH map around the origin -- HealpixRenderer (renderers.py:720-766).  By default the
    line-of-sight projection per solid angle; with ``shell_distance`` the quantity sampled on a thin shell at
    that radius with the 3-D kernel.  float32 ``(12 * nside**2,)``."""
    if nside is None:
        nside = config["image-default-nside"]
    k3 = kernels.create_kernel(kernel)
    k2 = None
    x, y, z = _columns(pos)
    is_shell = shell_distance is not None

    def one(q, proj):
        nonlocal k2
        if proj and k2 is None:
            k2 = k3.projection()
        return _render.render_spherical_image_core(rho, mass, q, x, y, z, smooth, int(nside), k2 if proj else k3,
                                                   float(shell_distance or 0.0))

    # by default spherical images are projected onto the sky; a weighted map starts from an unprojected base
    # and a thin shell is not a projection (renderers.py:946-960)
    projected = not is_shell and not weight
    return _multipass(one, rho if quantity is None else quantity, projected, weight, denoise)


def render_3d_grid(pos, mass, rho, smooth, quantity=None, width=None, resolution=None, nx=None, ny=None, nz=None,
                   kernel=None, smooth_range=(0.0, np.inf), boxsize=None, geometry=None, denoise=False):
    """SPH interpolation onto a 3-D grid -- Grid3dRenderer (renderers.py:700-718): a cube of side ``width``
    centred on the origin, float32 ``(nx, ny, nz)``; ``denoise`` divides by the grid of ones (renderers.py:488-506)."""
    g = geometry or ImageGeometry(width if width is not None else 200.0, resolution)
    if geometry is None:
        g.restrict_z_range()
        g.set_pixels(nx, ny, nz)
    k = kernels.create_kernel(kernel)
    x, y, z = _columns(pos)
    wraps = [wrapping_repeat_array(a, b, boxsize) for a, b in ((g.x1, g.x2), (g.y1, g.y2), (g.z1, g.z2))]

    def one(q, proj):
        return _render.to_3d_grid(g.nx, g.ny, g.nz, x, y, z, smooth, g.x1, g.x2, g.y1, g.y2, g.z1, g.z2, q, mass, rho,
                                  smooth_range[0], smooth_range[1], k, *wraps)

    return _multipass(one, rho if quantity is None else quantity, False, None, denoise)


def render(pos, mass, rho, smooth, quantity=None, target="image", nside=None, shell_distance=None, **kwargs):
    """Array-level counterpart of ``make_render_pipeline(...).render()`` (renderers.py:770-994): dispatch on
    ``target`` ('image', 'volume' or 'healpix') with the reference's argument checks."""
    if target not in ("image", "volume", "healpix"):
        raise ValueError("Unknown render target; options are 'image', 'volume' or 'healpix'")
    if nside is not None and target != "healpix":
        raise ValueError("nside is only valid in combination with target='healpix'")
    if shell_distance is not None and target != "healpix":
        raise ValueError("shell_distance is only valid in combination with target='healpix'")
    if target == "image":
        return render_image(pos, mass, rho, smooth, quantity, **kwargs)
    if target == "volume":
        return render_3d_grid(pos, mass, rho, smooth, quantity, **kwargs)
    return render_healpix(pos, mass, rho, smooth, quantity, nside=nside, shell_distance=shell_distance, **kwargs)


to_3d_grid = render_3d_grid
