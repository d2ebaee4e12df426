"""TEST INFRASTRUCTURE -- ctypes front-end of liboracle.so (oracle/oracle.c).

A CPU restatement of the reference's SPH smoothing hot path, used as the parity
checker for the CUDA implementation.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product path (``pynbody_b200``) never does.

Interface mirrors ``oracle.ref.RefKDTree`` (and therefore pynbody.kdtree.KDTree,
pynbody/kdtree/__init__.py:32-538) so tests can swap one for the other.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_vp, _i64, _i32, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle_kd.inc", "oracle_sph.inc")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_tree_build.restype = _vp
        L.orc_tree_build.argtypes = [_vp, _i64, _i32, _i32]
        L.orc_tree_free.argtypes = [_vp]
        L.orc_tree_nnodes.restype = _i64
        L.orc_tree_nnodes.argtypes = [_vp]
        L.orc_tree_export.argtypes = [_vp] * 8
        L.orc_smooth.argtypes = [_vp, _i32, _dbl, _vp, _i32]
        L.orc_nn.argtypes = [_vp, _i32, _dbl, _vp, _vp, _vp, _vp]
        L.orc_populate.argtypes = [_vp, _i32, _i32, _dbl, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32]
        L.orc_gather_counts.restype = None
        L.orc_gather_counts.argtypes = [_vp, _dbl, _vp, _vp, _i32]
        L.orc_ball.restype = _i64
        L.orc_ball.argtypes = [_vp, _dbl, _dbl, _dbl, _dbl, _dbl, _vp, _i64]
        L.orc_render_image.restype = None
        L.orc_render_image.argtypes = ([_i32, _i32, _i64, _vp, _vp, _vp, _i64, _i32] + [_vp, _i32] * 4
                                       + [_dbl] * 11 + [_vp, _i32, _i32, _dbl, _vp, _i32, _vp, _i32, _vp])
        L.orc_to_3d_grid.restype = None
        L.orc_to_3d_grid.argtypes = ([_i32, _i32, _i32, _i64, _vp, _vp, _vp, _i64, _i32] + [_vp, _i32] * 4
                                     + [_dbl] * 8 + [_vp, _i32, _i32, _dbl, _vp, _i32, _vp, _i32, _vp, _i32, _vp])
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


_PROPID = {"rho": 2, "qty_mean_1d": 3, "qty_mean_nd": 4, "qty_disp_1d": 5, "qty_disp_nd": 6,
           "qty_div": 7, "qty_curl": 8}


class OracleKDTree:
    def __init__(self, pos, mass, leafsize=16, boxsize=None, num_threads=None, kernel_id=0):
        pos = np.ascontiguousarray(pos)
        if pos.dtype not in (np.float32, np.float64):
            raise ValueError("Unsupported array dtype for kdtree")
        mass = np.ascontiguousarray(mass, dtype=pos.dtype)
        self._pos, self._mass = pos, mass
        self.leafsize = int(leafsize)
        self.boxsize = -1.0 if boxsize is None else float(boxsize)
        self.num_threads = int(num_threads or os.cpu_count() or 1)
        self.kernel_id = int(kernel_id)
        self.periodic_honoured = None
        self._h = lib().orc_tree_build(_p(pos), len(pos), int(pos.dtype == np.float64), self.leafsize)

    def export(self):
        L = lib()
        nn = L.orc_tree_nnodes(self._h)
        out = dict(split=np.empty(nn), bmin=np.empty((nn, 3)), bmax=np.empty((nn, 3)),
                   dim=np.empty(nn, np.int32), lo=np.empty(nn, np.int64), hi=np.empty(nn, np.int64),
                   order=np.empty(len(self._pos), np.int64))
        L.orc_tree_export(self._h, *[_p(out[k]) for k in ("split", "bmin", "bmax", "dim", "lo", "hi", "order")])
        return out

    def smooth(self, nn):
        sm = np.empty(len(self._pos), dtype=self._pos.dtype)
        r = lib().orc_smooth(self._h, int(nn), self.boxsize, _p(sm), self.num_threads)
        if r < 0:
            raise ValueError("Number of smoothing particles exceeds number of particles in tree")
        self.periodic_honoured = bool(r)
        self._smooth = sm
        return sm

    def all_nn(self, nn):
        n = len(self._pos)
        sm = np.empty(n, dtype=self._pos.dtype)
        idx = np.empty((n, nn), np.int64)
        d2 = np.empty((n, nn), self._pos.dtype)
        visit = np.empty(n, np.int64)
        r = lib().orc_nn(self._h, int(nn), self.boxsize, _p(sm), _p(idx), _p(d2), _p(visit))
        if r < 0:
            raise ValueError("Number of smoothing particles exceeds number of particles in tree")
        return visit, sm, idx, d2

    def _populate(self, mode, nn, smooth, rho, qty, out):
        qf64 = int(qty is not None and qty.dtype == np.float64)
        r = lib().orc_populate(self._h, _PROPID[mode], int(nn), self.boxsize, self.kernel_id,
                               _p(smooth), _p(self._mass), _p(rho), _p(qty), _p(out), qf64, self.num_threads)
        if r == -2:
            raise RuntimeError("Buffer overflow in smoothing operation.")
        if r != 0:
            raise ValueError("bad populate arguments")

    def rho(self, nn, smooth=None):
        smooth = self._smooth if smooth is None else np.ascontiguousarray(smooth, dtype=self._pos.dtype)
        rho = np.empty(len(self._pos), dtype=self._pos.dtype)
        self._populate("rho", nn, smooth, rho, None, None)
        self._rho = rho
        return rho

    def qty_op(self, op, qty, nn, smooth=None, rho=None):
        smooth = self._smooth if smooth is None else smooth
        rho = self._rho if rho is None else rho
        qty = np.ascontiguousarray(qty)
        nd = qty.ndim == 2
        out = np.empty_like(qty) if op in ("mean", "curl") else np.empty(len(qty), dtype=qty.dtype)
        mode = {"mean": "qty_mean_nd" if nd else "qty_mean_1d",
                "disp": "qty_disp_nd" if nd else "qty_disp_1d",
                "div": "qty_div", "curl": "qty_curl"}[op]
        self._populate(mode, nn, smooth, rho, qty, out)
        return out

    def gather_counts(self, smooth=None):
        """nCnt of smBallGather(4 h^2) for every particle (kdmain.cpp:779, smooth.h:270-324), file order."""
        smooth = self._smooth if smooth is None else np.ascontiguousarray(smooth, dtype=self._pos.dtype)
        out = np.empty(len(self._pos), np.int32)
        lib().orc_gather_counts(self._h, self.boxsize, _p(smooth), _p(out), self.num_threads)
        return out

    def particles_in_sphere(self, center, radius):
        cap = len(self._pos)
        out = np.empty(cap, np.int64)
        n = lib().orc_ball(self._h, self.boxsize, float(center[0]), float(center[1]), float(center[2]),
                           float(radius), _p(out), cap)
        return out[:n].copy()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tree_free(self._h)
            self._h = None


def _arr(a):
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.ascontiguousarray(a)


def _pos_cols(x, y, z):
    """x, y, z must share dtype; returns (px, py, pz, stride_in_elements, is_f64, keepalive)."""
    x, y, z = np.asarray(x), np.asarray(y), np.asarray(z)
    assert x.dtype == y.dtype == z.dtype
    s = {a.strides[0] for a in (x, y, z)}
    if len(s) != 1 or x.strides[0] % x.itemsize:
        x, y, z = (np.ascontiguousarray(a) for a in (x, y, z))
    return x, y, z, x.strides[0] // x.itemsize, int(x.dtype == np.float64)


def render_image(nx, ny, x, y, z, sm, x1, x2, y1, y2, z_camera, z0, qty, mass, rho,
                 smooth_lo, smooth_hi, z_lo, z_hi, min_smooth, kernel, wrap_x=(0.0,), wrap_y=(0.0,)):
    x, y, z, st, pf = _pos_cols(x, y, z)
    sm, qty, mass, rho = (_arr(a) for a in (sm, qty, mass, rho))
    lut = np.ascontiguousarray(kernel.get_samples(dtype=np.float32))
    wx, wy = np.asarray(wrap_x, np.float64), np.asarray(wrap_y, np.float64)
    out = np.empty((ny, nx), np.float32)
    f = lambda a: int(a.dtype == np.float64)
    lib().orc_render_image(nx, ny, len(x), _p(x), _p(y), _p(z), st, pf, _p(sm), f(sm), _p(qty), f(qty),
                           _p(mass), f(mass), _p(rho), f(rho), x1, x2, y1, y2, z_camera, z0,
                           smooth_lo, smooth_hi, z_lo, z_hi, min_smooth, _p(lut), len(lut),
                           int(kernel.h_power), float(kernel.max_d), _p(wx), len(wx), _p(wy), len(wy), _p(out))
    return out


def to_3d_grid(nx, ny, nz, x, y, z, sm, x1, x2, y1, y2, z1, z2, qty, mass, rho,
               smooth_lo, smooth_hi, kernel, wrap_x=(0.0,), wrap_y=(0.0,), wrap_z=(0.0,)):
    if kernel.h_power < 3:
        raise ValueError("Cannot render to 3D grid without 3-dimensional kernel or greater")
    x, y, z, st, pf = _pos_cols(x, y, z)
    sm, qty, mass, rho = (_arr(a) for a in (sm, qty, mass, rho))
    lut = np.ascontiguousarray(kernel.get_samples(dtype=np.float32))
    wx, wy, wz = (np.asarray(w, np.float64) for w in (wrap_x, wrap_y, wrap_z))
    out = np.empty((nx, ny, nz), np.float32)
    f = lambda a: int(a.dtype == np.float64)
    lib().orc_to_3d_grid(nx, ny, nz, len(x), _p(x), _p(y), _p(z), st, pf, _p(sm), f(sm), _p(qty), f(qty),
                         _p(mass), f(mass), _p(rho), f(rho), x1, x2, y1, y2, z1, z2, smooth_lo, smooth_hi,
                         _p(lut), len(lut), int(kernel.h_power), float(kernel.max_d),
                         _p(wx), len(wx), _p(wy), len(wy), _p(wz), len(wz), _p(out))
    return out


def render_spherical_image_core(rho, mass, qty, x, y, z, h, nside, kernel, shell_radius=0.0):
    """_render.render_spherical_image_core (pynbody/sph/_render.pyx:51-179) restated."""
    if nside & (nside - 1):
        raise ValueError("nside value must be a power of 2")
    if kernel.h_power not in (2, 3):
        raise ValueError("Only kernels of dimension 2 (projected) or 3 (thin shell) are supported for healpix maps")
    if kernel.h_power == 3 and shell_radius <= 0.0:
        raise ValueError("A positive shell_radius is required to render a 3D (thin shell) healpix map")
    x, y, z, st, pf = _pos_cols(x, y, z)
    rho, mass, qty, h = (_arr(a) for a in (rho, mass, qty, h))
    lut = np.ascontiguousarray(kernel.get_samples(dtype=np.float32))
    out = np.empty(12 * nside * nside, np.float32)
    f = lambda a: int(a.dtype == np.float64)
    L = lib()
    L.orc_render_healpix.restype = None
    L.orc_render_healpix.argtypes = [_i64, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _vp, _i32, _i32,
                                     _vp, _i32, _i32, _dbl, _dbl, _vp]
    L.orc_render_healpix(len(x), _p(rho), f(rho), _p(mass), f(mass), _p(qty), f(qty), _p(x), _p(y), _p(z), st, pf,
                         _p(h), f(h), int(nside), _p(lut), len(lut), int(kernel.h_power), float(kernel.max_d),
                         float(shell_radius), _p(out))
    return out


def selection(pos_ar, region, parameters, wrap):
    """filt.geometry_selection.selection (pynbody/filt/geometry_selection.pyx:31-103) restated in numpy.

    Element-wise numpy arithmetic in the array's own dtype performs exactly the reference's operations
    (geometry_selection.hpp:11-28 offsets and wrap, :47-62 sphere, :64-77 cuboid), one IEEE operation
    at a time, so the flags are bit-identical; pinned against the compiled reference module in
    tests/test_oracle.py."""
    if not isinstance(pos_ar, np.ndarray) or pos_ar.dtype not in (np.float32, np.float64) or pos_ar.ndim != 2:
        raise TypeError("No matching signature found")
    out = np.empty(len(pos_ar), np.uint8)
    if len(pos_ar) == 0:
        return out
    if pos_ar.shape[1] != 3 or pos_ar.strides != (3 * pos_ar.itemsize, pos_ar.itemsize):
        raise ValueError("Input array must be C-contiguous and have 3 columns")
    T = pos_ar.dtype.type
    wrap = T(wrap)
    half = wrap / T(2)

    def offsets(centre):
        ds = []
        for c in range(3):
            d = pos_ar[:, c] - centre[c]
            if wrap > 0:
                d = np.where(d > half, d - wrap, d)
                d = np.where(d < -half, d + wrap, d)
            ds.append(d)
        return ds

    if region == "sphere":
        if len(parameters) != 4:
            raise ValueError("Sphere selection requires 4 parameters: (x0, y0, z0, r_max)")
        x0, y0, z0, r = (T(p) for p in parameters)
        dx, dy, dz = offsets((x0, y0, z0))
        out[:] = (dx * dx + dy * dy) + dz * dz < r * r
    elif region == "cube":
        if len(parameters) != 6:
            raise ValueError("Cube selection requires 6 parameters: (x0, y0, z0, x1, y1, z1)")
        x0, y0, z0, x1, y1, z1 = (T(p) for p in parameters)
        two = T(2)
        dx, dy, dz = offsets(((x0 + x1) / two, (y0 + y1) / two, (z0 + z1) / two))
        out[:] = (np.abs(dx) < (x1 - x0) / two) & (np.abs(dy) < (y1 - y0) / two) & (np.abs(dz) < (z1 - z0) / two)
    else:
        raise ValueError("Region must be either 'sphere' or 'cube'")
    return out
