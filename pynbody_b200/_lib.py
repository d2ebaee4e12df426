"""ctypes binding of ``libpnb_b200.so`` (C ABI declared in ``include/pnb_b200.h``).

The shared library holds the hand-written sm_100a CUDA kernels.  It is required:
there is no CPU fallback, and importing a compute entry point without the built
library raises ``ImportError`` with the build command.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PNB_B200_LIB selects another build of the same library (developer builds with kernel statistics: `make stats`)
LIB_PATH = os.environ.get("PNB_B200_LIB") or os.path.join(_HERE, "libpnb_b200.so")
CSRC = os.path.join(_HERE, "csrc")

OK, ERR_VALUE, ERR_TYPE, ERR_RUNTIME, ERR_CUDA, ERR_MEMORY = 0, -1, -2, -3, -4, -5
F32, F64 = 0, 1
HOST, DEVICE = 0, 1

_vp, _i64, _i32, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
_pi32 = C.POINTER(C.c_int)
_pi64 = C.POINTER(C.c_int64)


class RenderArrays(C.Structure):
    """``pnb_render_arrays`` (include/pnb_b200.h)."""
    _fields_ = [("x", _vp), ("y", _vp), ("z", _vp), ("pos_stride", _i64), ("pos_dtype", _i32),
                ("sm", _vp), ("sm_stride", _i64), ("sm_dtype", _i32),
                ("qty", _vp), ("qty_stride", _i64), ("qty_dtype", _i32),
                ("mass", _vp), ("mass_stride", _i64), ("mass_dtype", _i32),
                ("rho", _vp), ("rho_stride", _i64), ("rho_dtype", _i32),
                ("n", _i64), ("mem", _i32)]


# name -> (restype, argtypes); every symbol include/pnb_b200.h declares
SIGNATURES = {
    "pnb_last_error": (C.c_char_p, []),
    "pnb_version": (C.c_char_p, []),
    "pnb_device_count": (_i32, []),
    "pnb_set_device": (_i32, [_i32]),
    "pnb_set_option": (_i32, [C.c_char_p, _i32]),
    "pnb_tree_create": (_vp, [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32]),
    "pnb_tree_destroy": (None, [_vp]),
    "pnb_tree_num_particles": (_i64, [_vp]),
    "pnb_tree_bounds": (_i32, [_vp, _vp]),
    "pnb_tree_node_count": (_i64, [_vp]),
    "pnb_tree_order": (_i32, [_vp, _vp]),
    "pnb_tree_export": (_i32, [_vp, _vp, _i64, _vp]),
    "pnb_tree_num_buckets": (_i64, [_vp]),
    "pnb_tree_buckets": (_i32, [_vp, _vp, _vp]),
    "pnb_tree_node_bounds": (_i32, [_vp, _vp]),
    "pnb_set_array": (_i32, [_vp, _i32, _vp, _i64, _i64, _i32, _i32, _i32]),
    "pnb_set_query_mask": (_i32, [_vp, _vp, _i32]),
    "pnb_set_query_mask2": (_i32, [_vp, _vp, _i32, _i32, _i32]),
    "pnb_mark_near_faces": (_i32, [_vp, _i32, _vp, _vp, _vp, _i32]),
    "pnb_mark_in_balls": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _dbl, _vp, _i32]),
    "pnb_route_plan": (_i32, [_vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    "pnb_route_pack": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _vp, _i32, _vp, _vp]),
    "pnb_populate": (_i32, [_vp, _i32, _i32, _i32, _dbl, _pi32]),
    "pnb_knn": (_i32, [_vp, _i32, _dbl, _vp, _vp, _vp, _pi32]),
    "pnb_knn_start": (_i32, [_vp, _i32, _dbl, _pi32]),
    "pnb_knn_rows": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "pnb_gather_counts": (_i32, [_vp, _dbl, _vp, _i32]),
    "pnb_ball_query": (_i32, [_vp, _dbl, _dbl, _dbl, _dbl, _dbl, _vp, _i64, _pi64]),
    "pnb_render_image": (_i32, [_i32, _i32, C.POINTER(RenderArrays)] + [_dbl] * 11
                         + [_vp, _i32, _i32, _dbl, _vp, _i32, _vp, _i32, _vp, _i32]),
    "pnb_render_grid": (_i32, [_i32, _i32, _i32, C.POINTER(RenderArrays)] + [_dbl] * 8
                        + [_vp, _i32, _i32, _dbl, _vp, _i32, _vp, _i32, _vp, _i32, _vp, _i32]),
    "pnb_render_healpix": (_i32, [C.POINTER(RenderArrays), _i32, _vp, _i32, _i32, _dbl, _dbl, _vp, _i32]),
    "pnb_select_region": (_i32, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _dbl, _vp, _i32]),
    "pnb_release_cached_memory": (None, []),
    "pnb_last_device_ms": (_dbl, [_i32]),
    "pnb_last_launch_count": (_i64, [_i32]),
    "pnb_last_kernel_ms": (_dbl, [_i32]),
    "pnb_total_launches": (_i64, []),
}

_LIB = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources in ``csrc/`` for sm_100a into ``libpnb_b200.so`` (needs nvcc)."""
    args = ["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (or __graft_entry__.build()). "
                "pynbody_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def check(rc: int):
    """Map a status code to the exception class the reference raises for that condition."""
    if rc == OK:
        return
    msg = lib().pnb_last_error().decode("utf-8", "replace")
    if rc == ERR_VALUE:
        raise ValueError(msg)
    if rc == ERR_TYPE:
        raise TypeError(msg)
    if rc == ERR_MEMORY:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch") and hasattr(a, "data_ptr")


def dtype_code(a) -> int:
    """PNB_F32 / PNB_F64 for a numpy array or torch tensor; ValueError otherwise (kdmain.cpp:136-182)."""
    if _is_torch(a):
        name = str(a.dtype)
        if name == "torch.float32":
            return F32
        if name == "torch.float64":
            return F64
        raise ValueError("Unsupported array dtype for kdtree")
    dt = np.asarray(a).dtype
    if dt == np.float32:
        return F32
    if dt == np.float64:
        return F64
    raise ValueError("Unsupported array dtype for kdtree")


def describe(a):
    """(pointer, stride0_bytes, stride1_bytes, ncols, dtype_code, mem, keepalive) of a 1-D or N x C array.

    numpy arrays are host memory, torch CUDA tensors device memory (PyTorch is only the
    device-buffer container here).
    """
    if _is_torch(a):
        es = a.element_size()
        st = [s * es for s in a.stride()]
        mem = DEVICE if a.is_cuda else HOST
        ptr = a.data_ptr()
        shape = tuple(a.shape)
        keep = a
    else:
        if not isinstance(a, np.ndarray):
            a = np.asarray(a)
        st = list(a.strides)
        mem = HOST
        ptr = a.ctypes.data
        shape = a.shape
        keep = a
    code = dtype_code(keep)
    if len(shape) == 1:
        return ptr, st[0], 0, 1, code, mem, keep
    if len(shape) == 2:
        return ptr, st[0], st[1], shape[1], code, mem, keep
    raise ValueError("array must be one- or two-dimensional")
