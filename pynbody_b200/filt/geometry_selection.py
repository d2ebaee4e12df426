"""Drop-in for the reference's Cython module ``pynbody.filt.geometry_selection``.

``selection(pos_ar, region, parameters, wrap)`` has the reference's signature, argument meaning and
errors (pynbody/filt/geometry_selection.pyx:31-103); the work is done by ``pnb_select_region`` in
``libpnb_b200.so`` (csrc/select.cu).  Callers in the reference: ``Sphere.__call__`` /
``Sphere.cubic_cell_intersection`` / ``Cuboid._get_mask`` (pynbody/filt/__init__.py:232,257,299).

``pos_ar`` may be a numpy array (host; a uint8 numpy array is returned, as the reference does) or a
torch CUDA tensor (device; a uint8 CUDA tensor is returned).
"""
from __future__ import annotations

import numpy as np

from .. import _lib

_REGIONS = {"sphere": (0, 4, "Sphere selection requires 4 parameters: (x0, y0, z0, r_max)"),
            "cube": (1, 6, "Cube selection requires 6 parameters: (x0, y0, z0, x1, y1, z1)")}


def selection(pos_ar, region, parameters, wrap):
    """uint8 flags of the particles inside a sphere ``(x0, y0, z0, r_max)`` or a cuboid
    ``(x0, y0, z0, x1, y1, z1)``; ``wrap > 0`` is the box size of a periodic volume (offsets of at most
    one box length are folded, as in the reference), ``wrap <= 0`` means no wrapping."""
    on_device = _lib._is_torch(pos_ar)
    if on_device:
        import torch
        if pos_ar.dtype not in (torch.float32, torch.float64) or pos_ar.dim() != 2:
            raise TypeError("No matching signature found")
        n = int(pos_ar.shape[0])
        out = torch.empty(n, dtype=torch.uint8, device=pos_ar.device)
        if n == 0:
            return out
        if pos_ar.shape[1] != 3 or not pos_ar.is_contiguous():
            raise ValueError("Input array must be C-contiguous and have 3 columns")
        dtype = _lib.F64 if pos_ar.dtype == torch.float64 else _lib.F32
        ptr, outp, mem = pos_ar.data_ptr(), out.data_ptr(), _lib.DEVICE
    else:
        if not isinstance(pos_ar, np.ndarray) or pos_ar.dtype not in (np.float32, np.float64) or pos_ar.ndim != 2:
            raise TypeError("No matching signature found")
        n = len(pos_ar)
        out = np.empty(n, dtype=np.uint8)
        if n == 0:
            return out
        item = pos_ar.itemsize
        if pos_ar.shape[1] != 3 or pos_ar.strides != (3 * item, item):
            raise ValueError("Input array must be C-contiguous and have 3 columns")
        dtype = _lib.F64 if pos_ar.dtype == np.float64 else _lib.F32
        ptr, outp, mem = pos_ar.ctypes.data, out.ctypes.data, _lib.HOST
    if region not in _REGIONS:
        raise ValueError("Region must be either 'sphere' or 'cube'")
    code, count, message = _REGIONS[region]
    if len(parameters) != count:
        raise ValueError(message)
    prm = np.array([float(p) for p in parameters], dtype=np.float64)
    _lib.check(_lib.lib().pnb_select_region(ptr, n, dtype, mem, code, prm.ctypes.data, count, float(wrap), outp, mem))
    return out


def last_selection_ms():
    """(device ms of the whole last selection call incl. copies, ms of the selection kernel launch)."""
    L = _lib.lib()
    return L.pnb_last_device_ms(3), L.pnb_last_kernel_ms(3)
