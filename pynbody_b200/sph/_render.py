"""Drop-in for the reference's Cython module ``pynbody.sph._render`` (render_image / to_3d_grid).

Same positional signatures and return types as pynbody/sph/_render.pyx:209-223 and :373-385; the
work is done by ``pnb_render_image`` / ``pnb_render_grid`` in ``libpnb_b200.so`` (tile-binned gather
renderer, csrc/render.cu).  ``kernel`` is any object with ``h_power``, ``max_d`` and
``get_samples(dtype=np.float32)`` -- pynbody's own kernel objects or ``pynbody_b200.sph.kernels``.

Arrays may be numpy arrays (host) or torch CUDA tensors (device); the seven particle arrays must
live in the same memory space.  The image is returned as a numpy float32 array for host inputs and
as a torch CUDA tensor for device inputs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib


def _common_pos(x, y, z):
    """x, y, z as 1-D arrays with one dtype and one byte stride (columns of an N x 3 array qualify)."""
    if _lib._is_torch(x):
        es = x.element_size()
        if not (x.dtype == y.dtype == z.dtype) or not (x.stride() == y.stride() == z.stride()):
            x, y, z = (a.to(x.dtype).contiguous() for a in (x, y, z))
        return x, y, z, x.stride(0) * es
    x, y, z = (np.asarray(a) for a in (x, y, z))
    if not (x.dtype == y.dtype == z.dtype):
        dt = np.result_type(x.dtype, y.dtype, z.dtype)
        x, y, z = (np.ascontiguousarray(a, dtype=dt) for a in (x, y, z))
    if x.dtype not in (np.float32, np.float64):
        x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    if not (x.strides == y.strides == z.strides) or x.strides[0] % x.itemsize:
        x, y, z = (np.ascontiguousarray(a) for a in (x, y, z))
    return x, y, z, x.strides[0]


def _one(a):
    if _lib._is_torch(a):
        return a, a.stride(0) * a.element_size()
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    if a.ndim != 1:
        raise ValueError("particle arrays must be one-dimensional")
    if a.strides[0] % a.itemsize:
        a = np.ascontiguousarray(a)
    return a, a.strides[0]


def _arrays(x, y, z, sm, qty, mass, rho):
    x, y, z, ps = _common_pos(x, y, z)
    n = int(x.shape[0])
    keep = [x, y, z]
    ra = _lib.RenderArrays()
    mems = set()

    def ptr(a):
        p, _, _, _, code, mem, _ = _lib.describe(a)
        mems.add(mem)
        return p, code

    ra.x, ra.pos_dtype = ptr(x)
    ra.y, _ = ptr(y)
    ra.z, _ = ptr(z)
    ra.pos_stride = ps
    for name, arr in (("sm", sm), ("qty", qty), ("mass", mass), ("rho", rho)):
        arr, st = _one(arr)
        if int(arr.shape[0]) != n:
            raise ValueError("particle arrays must all have the same length")
        keep.append(arr)
        p, code = ptr(arr)
        setattr(ra, name, p)
        setattr(ra, name + "_stride", st)
        setattr(ra, name + "_dtype", code)
    if len(mems) > 1:
        raise ValueError("particle arrays must all be host arrays or all be device tensors")
    ra.n = n
    ra.mem = mems.pop() if mems else _lib.HOST
    return ra, keep


def _kernel_table(kernel):
    lut = np.ascontiguousarray(kernel.get_samples(dtype=np.float32), dtype=np.float32)
    return lut, int(kernel.h_power), float(kernel.max_d)


def _wrap(w):
    w = np.ascontiguousarray(np.atleast_1d(np.asarray(w, dtype=np.float64)))
    return w, w.ctypes.data, len(w)


def _output(shape, mem):
    if mem == _lib.DEVICE:
        import torch
        out = torch.empty(shape, dtype=torch.float32, device="cuda")
        return out, out.data_ptr()
    out = np.empty(shape, dtype=np.float32)
    return out, out.ctypes.data


def render_image(nx, ny, x, y, z, sm, x1, x2, y1, y2, z_camera, z0, qty, mass, rho,
                 smooth_lo, smooth_hi, z_lo, z_hi, min_smooth, kernel,
                 wrap_offsets_x=(0.0,), wrap_offsets_y=(0.0,)):
    """SPH image, float32 ``(ny, nx)`` (pynbody/sph/_render.pyx:209-365)."""
    ra, keep = _arrays(x, y, z, sm, qty, mass, rho)
    lut, h_power, max_d = _kernel_table(kernel)
    wx, wxp, nwx = _wrap(wrap_offsets_x)
    wy, wyp, nwy = _wrap(wrap_offsets_y)
    out, outp = _output((int(ny), int(nx)), ra.mem)
    _lib.check(_lib.lib().pnb_render_image(
        int(nx), int(ny), C.byref(ra), float(x1), float(x2), float(y1), float(y2), float(z_camera), float(z0),
        float(smooth_lo), float(smooth_hi), float(z_lo), float(z_hi), float(min_smooth),
        lut.ctypes.data, len(lut), h_power, max_d, wxp, nwx, wyp, nwy, outp, ra.mem))
    return out


def to_3d_grid(nx, ny, nz, x, y, z, sm, x1, x2, y1, y2, z1, z2, qty, mass, rho,
               smooth_lo, smooth_hi, kernel, wrap_offsets_x=(0.0,), wrap_offsets_y=(0.0,), wrap_offsets_z=(0.0,)):
    """SPH-interpolated 3-D grid, float32 ``(nx, ny, nz)`` (pynbody/sph/_render.pyx:373-497)."""
    if kernel.h_power < 3:
        raise ValueError("Cannot render to 3D grid without 3-dimensional kernel or greater")
    ra, keep = _arrays(x, y, z, sm, qty, mass, rho)
    lut, h_power, max_d = _kernel_table(kernel)
    wx, wxp, nwx = _wrap(wrap_offsets_x)
    wy, wyp, nwy = _wrap(wrap_offsets_y)
    wz, wzp, nwz = _wrap(wrap_offsets_z)
    out, outp = _output((int(nx), int(ny), int(nz)), ra.mem)
    _lib.check(_lib.lib().pnb_render_grid(
        int(nx), int(ny), int(nz), C.byref(ra), float(x1), float(x2), float(y1), float(y2), float(z1), float(z2),
        float(smooth_lo), float(smooth_hi), lut.ctypes.data, len(lut), h_power, max_d,
        wxp, nwx, wyp, nwy, wzp, nwz, outp, ra.mem))
    return out


def render_spherical_image_core(rho, mass, qtyar, x, y, z, h, nside, kernel, shell_radius=0.0):
    """Healpix (RING) SPH image around the origin, float32 ``(12 * nside**2,)``
    (pynbody/sph/_render.pyx:51-179): projected column per solid angle for a 2-D kernel, thin shell at
    ``shell_radius`` for a 3-D kernel."""
    nside = int(nside)
    if nside & (nside - 1) != 0:
        raise ValueError("nside value must be a power of 2")
    if kernel.h_power not in (2, 3):
        raise ValueError("Only kernels of dimension 2 (projected) or 3 (thin shell) are supported for healpix maps")
    if kernel.h_power == 3 and shell_radius <= 0.0:
        raise ValueError("A positive shell_radius is required to render a 3D (thin shell) healpix map")
    ra, keep = _arrays(x, y, z, h, qtyar, mass, rho)
    lut, h_power, max_d = _kernel_table(kernel)
    out, outp = _output((12 * nside * nside,), ra.mem)
    _lib.check(_lib.lib().pnb_render_healpix(C.byref(ra), nside, lut.ctypes.data, len(lut), h_power, max_d,
                                             float(shell_radius), outp, ra.mem))
    return out


def last_render_ms():
    """(device ms of the whole last render call, ms spent in the tile render kernel launches)."""
    L = _lib.lib()
    return L.pnb_last_device_ms(2), L.pnb_last_kernel_ms(2)
