"""TEST INFRASTRUCTURE -- not part of the product path.

Thin driver around the *unmodified* reference extension modules compiled by
``oracle/Makefile`` into ``oracle/_ref/`` (``kdmain`` from
pynbody/kdtree/{kdmain,kd,smooth}.cpp and ``_render`` from
pynbody/sph/_render.pyx).  It calls them with raw numpy arrays in exactly the
sequence the reference host code does:

* ``RefKDTree``     follows pynbody/kdtree/__init__.py:79-128, 167-227, 249-375
* ``render_image``  follows pynbody/sph/renderers.py:690-697
* ``to_3d_grid``    follows pynbody/sph/renderers.py:711-718

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.  Nothing
here reads /root/reference at run time: the compiled modules travel with the
repo snapshot.
"""
from __future__ import annotations

import os
import sys
import threading
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")

kdmain = None
_render = None


def available() -> bool:
    return _load(quiet=True)


def _load(quiet=False) -> bool:
    global kdmain, _render
    if kdmain is not None:
        return True
    if not os.path.isdir(_REF):
        if quiet:
            return False
        raise ImportError("oracle/_ref not built: run `make -C oracle ref` where /root/reference exists")
    sys.path.insert(0, _REF)
    try:
        import kdmain as _k  # type: ignore
        import _render as _r  # type: ignore
    except ImportError:
        if quiet:
            return False
        raise
    finally:
        sys.path.remove(_REF)
    kdmain, _render = _k, _r
    return True


# numpy mirror of kd.h:29-40 (pynbody/kdtree/__init__.py:17-30)
Boundary = np.dtype([("fMin", np.float64, (3,)), ("fMax", np.float64, (3,))])
KDNode = np.dtype([("fSplit", np.float64), ("bnd", Boundary), ("iDim", np.int32),
                   ("_pad", np.int32), ("pLower", np.intp), ("pUpper", np.intp)])

_ARRAY_ID = {"smooth": 0, "rho": 1, "mass": 2, "qty": 3, "qty_sm": 4}
_PROPID = {"hsm": 1, "rho": 2, "qty_mean_1d": 3, "qty_mean_nd": 4, "qty_disp_1d": 5,
           "qty_disp_nd": 6, "qty_div": 7, "qty_curl": 8}


def _thread_map(func, *args):
    """pynbody/util/__init__.py:468 -- one fresh Python thread per call."""
    excs = []

    def run(*a):
        try:
            func(*a)
        except BaseException as e:  # noqa: BLE001
            excs.append(e)

    ts = [threading.Thread(target=run, args=a) for a in zip(*args)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if excs:
        raise excs[0]


class RefKDTree:
    """Reference kd-tree driven exactly as pynbody.kdtree.KDTree drives kdmain."""

    def __init__(self, pos, mass, leafsize=16, boxsize=None, num_threads=None, kernel_id=0):
        _load()
        if num_threads is None:
            num_threads = os.cpu_count() or 1
        self.num_threads = int(num_threads)
        self.leafsize = int(leafsize)
        self.boxsize = -1.0 if boxsize is None else float(boxsize)
        self._pos, self._mass = pos, mass
        self.kernel_id = int(kernel_id)
        self.kdtree = kdmain.init(pos, mass, self.leafsize)
        n_nodes = kdmain.get_node_count(self.kdtree)
        self.kdnodes = np.zeros(n_nodes, dtype=KDNode)
        self.particle_offsets = np.empty(len(pos), dtype=np.intp)
        kdmain.build(self.kdtree, self.kdnodes, self.particle_offsets,
                     2 ** int(np.log2(self.num_threads)))

    def set_array_ref(self, name, ar):
        kdmain.set_arrayref(self.kdtree, _ARRAY_ID[name], ar)

    def populate(self, mode, nn):
        smx = kdmain.nn_start(self.kdtree, int(nn), self.boxsize)
        try:
            propid = _PROPID[mode]
            if propid == 1:
                kdmain.domain_decomposition(self.kdtree, self.num_threads)
            if self.num_threads == 1:
                kdmain.populate(self.kdtree, smx, propid, 0, self.kernel_id)
            else:
                n = self.num_threads
                _thread_map(kdmain.populate, [self.kdtree] * n, [smx] * n, [propid] * n,
                            list(range(n)), [self.kernel_id] * n)
        finally:
            kdmain.nn_stop(self.kdtree, smx)

    # -- conveniences mirroring sph/__init__.py:40-94 and derived.py:126-154 ------------
    def smooth(self, nn):
        sm = np.empty(len(self._pos), dtype=self._pos.dtype)
        self.set_array_ref("smooth", sm)
        self.populate("hsm", nn)
        self._smooth = sm
        return sm

    def rho(self, nn, smooth=None):
        if smooth is None:
            smooth = getattr(self, "_smooth", None)
            if smooth is None:
                smooth = self.smooth(nn)
        rho = np.empty(len(self._pos), dtype=self._pos.dtype)
        self.set_array_ref("smooth", smooth)
        self.set_array_ref("mass", self._mass)
        self.set_array_ref("rho", rho)
        self.populate("rho", nn)
        self._rho = rho
        return rho

    def qty_op(self, op, qty, nn, smooth=None, rho=None):
        """op in mean|disp|div|curl; qty is N or N x 3."""
        smooth = self._smooth if smooth is None else smooth
        rho = self._rho if rho is None else rho
        nd = qty.ndim == 2
        if op in ("mean", "curl"):
            out = np.empty_like(qty)
        else:
            out = np.empty(len(qty), dtype=qty.dtype)
        mode = {"mean": "qty_mean_nd" if nd else "qty_mean_1d",
                "disp": "qty_disp_nd" if nd else "qty_disp_1d",
                "div": "qty_div", "curl": "qty_curl"}[op]
        self.set_array_ref("rho", rho)
        self.set_array_ref("smooth", smooth)
        self.set_array_ref("mass", self._mass)
        self.set_array_ref("qty", qty)
        self.set_array_ref("qty_sm", out)
        self.populate(mode, nn)
        return out

    def all_nn(self, nn):
        """kdtree/__init__.py:191-231 -- list of [i, h, [j...], [d2...]] in snake order."""
        sm = np.empty(len(self._pos), dtype=self._pos.dtype)
        self.set_array_ref("smooth", sm)
        smx = kdmain.nn_start(self.kdtree, int(nn), self.boxsize)
        kdmain.domain_decomposition(self.kdtree, 1)
        out = []
        while True:
            r = kdmain.nn_next(self.kdtree, smx)
            if r is None:
                break
            out.append(r)
        kdmain.nn_stop(self.kdtree, smx)
        return out

    def particles_in_sphere(self, center, radius):
        smx = kdmain.nn_start(self.kdtree, 1, self.boxsize)
        ids = kdmain.particles_in_sphere(self.kdtree, smx, center[0], center[1], center[2], radius)
        kdmain.nn_stop(self.kdtree, smx)
        return ids

    def __del__(self):
        if getattr(self, "kdtree", None) is not None and kdmain is not None:
            kdmain.free(self.kdtree)
            self.kdtree = None


class LutKernel:
    """Object with the three attributes _render reads (_render.pyx:245-250)."""

    def __init__(self, samples, h_power, max_d=2):
        self._samples = np.ascontiguousarray(samples, dtype=np.float32)
        self.h_power = int(h_power)
        self.max_d = max_d

    def get_samples(self, dtype=np.float32):
        return self._samples.astype(dtype, copy=False)


def render_image(nx, ny, x, y, z, sm, x1, x2, y1, y2, z_camera, z0, qty, mass, rho,
                 smooth_lo, smooth_hi, z_lo, z_hi, min_smooth, kernel,
                 wrap_x=(0.0,), wrap_y=(0.0,), num_threads=1):
    """_render.render_image, optionally split i::n over threads and summed as
    ThreadedImageRenderer does (renderers.py:558-573)."""
    _load()
    if num_threads <= 1:
        return _render.render_image(nx, ny, x, y, z, sm, x1, x2, y1, y2, z_camera, z0, qty, mass, rho,
                                    smooth_lo, smooth_hi, z_lo, z_hi, min_smooth, kernel,
                                    list(wrap_x), list(wrap_y))
    res = [None] * num_threads

    def work(i):
        s = slice(i, None, num_threads)
        res[i] = _render.render_image(nx, ny, x[s], y[s], z[s], sm[s], x1, x2, y1, y2, z_camera, z0,
                                      qty[s], mass[s], rho[s], smooth_lo, smooth_hi, z_lo, z_hi,
                                      min_smooth, kernel, list(wrap_x), list(wrap_y))

    _thread_map(work, list(range(num_threads)))
    return sum(res)


def to_3d_grid(nx, ny, nz, x, y, z, sm, x1, x2, y1, y2, z1, z2, qty, mass, rho,
               smooth_lo, smooth_hi, kernel, wrap_x=(0.0,), wrap_y=(0.0,), wrap_z=(0.0,), num_threads=1):
    _load()
    if num_threads <= 1:
        return _render.to_3d_grid(nx, ny, nz, x, y, z, sm, x1, x2, y1, y2, z1, z2, qty, mass, rho,
                                  smooth_lo, smooth_hi, kernel, list(wrap_x), list(wrap_y), list(wrap_z))
    res = [None] * num_threads

    def work(i):
        s = slice(i, None, num_threads)
        res[i] = _render.to_3d_grid(nx, ny, nz, x[s], y[s], z[s], sm[s], x1, x2, y1, y2, z1, z2,
                                    qty[s], mass[s], rho[s], smooth_lo, smooth_hi, kernel,
                                    list(wrap_x), list(wrap_y), list(wrap_z))

    _thread_map(work, list(range(num_threads)))
    return sum(res)


def render_spherical_image_core(rho, mass, qty, x, y, z, h, nside, kernel, shell_radius=0.0):
    """The reference's healpix renderer (pynbody/sph/_render.pyx:51-179), as HealpixRenderer calls it
    (pynbody/sph/renderers.py:720-766)."""
    _load()
    return _render.render_spherical_image_core(rho, mass, qty, x, y, z, h, nside, kernel, shell_radius)


def selection(pos_ar, region, parameters, wrap):
    """The reference's compiled ``pynbody.filt.geometry_selection.selection``
    (pynbody/filt/geometry_selection.pyx:31-103), built by oracle/Makefile into ``_ref/pnbref/filt``."""
    if not os.path.isdir(os.path.join(_REF, "pnbref")):
        raise ImportError("oracle/_ref/pnbref not built: run `make -C oracle ref` where /root/reference exists")
    sys.path.insert(0, _REF)
    try:
        from pnbref.filt import geometry_selection  # type: ignore
    finally:
        sys.path.remove(_REF)
    return geometry_selection.selection(pos_ar, region, parameters, wrap)
