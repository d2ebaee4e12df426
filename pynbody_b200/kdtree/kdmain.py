"""Drop-in for the reference's native module ``pynbody.kdtree.kdmain``.

Same fourteen functions with the same argument meaning and error behaviour as the
CPython extension's method table (pynbody/kdtree/kdmain.cpp:63-85), implemented
on ``libpnb_b200.so``.  Assigning this module to ``pynbody.kdtree.kdmain`` makes
the reference's own ``KDTree`` class (pynbody/kdtree/__init__.py:32-538) run on
the GPU unchanged; ``pynbody_b200.kdtree.KDTree`` is the same class restated on
top of it.

Arrays may be numpy arrays (host memory) or torch CUDA tensors (device memory).
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from .. import _lib

# numpy mirrors of kd.h:29-40 (pynbody/kdtree/__init__.py:17-30)
Boundary = np.dtype([("fMin", np.float64, (3,)), ("fMax", np.float64, (3,))])
KDNode = np.dtype([("fSplit", np.float64), ("bnd", Boundary), ("iDim", np.int32), ("_pad", np.int32),
                   ("pLower", np.intp), ("pUpper", np.intp)])

_PERIOD_WARNING = ("\nThe particles span a region larger than the specified boxsize; disabling periodicity.\n\n"
                   "For more information about this warning, see the module documentation for KDTree, \n"
                   "https://pynbody.readthedocs.io/latest/reference/_autosummary/pynbody.kdtree.html")


class _KDContext:
    """What the reference keeps in its ``KDContext`` capsule (kd.h:42-66)."""

    def __init__(self, pos, mass, leafsize):
        self.pos, self.mass, self.leafsize = pos, mass, int(leafsize)
        self.n = int(pos.shape[0])
        self.handle = None
        self.arrays = [None] * 5
        self.dtype_code = _lib.dtype_code(pos)

    def require_built(self):
        if self.handle is None:
            raise ValueError("Invalid KDContext object: tree has not been built")
        return self.handle


class _SmoothingContext:
    """``SmoothingContext`` capsule (smooth.h:23-68): neighbour count, period, nn() iterator state."""

    def __init__(self, k, period):
        self.k, self.period = int(k), float(period)
        self.nn_result = None       # the chunk of neighbour-list rows currently on the host
        self.nn_first = 0           # file index of its first row
        self.nn_started = False     # the device search of this nn() iteration has run
        self.nn_cursor = 0


def _check_shape(pos, mass):
    if len(pos.shape) != 2 or pos.shape[1] != 3:
        raise ValueError("pos must have shape (N,3)")
    if mass is not None and (len(mass.shape) != 1 or mass.shape[0] != pos.shape[0]):
        raise ValueError("mass must have shape (N,)")


def init(pos, mass, leafsize):
    """kdinit (kdmain.cpp:189-240): validate dtypes/shapes and create the context."""
    code = _lib.dtype_code(pos)                       # ValueError for unsupported dtypes
    _check_shape(pos, mass)
    if mass is not None and _lib.dtype_code(mass) != code:
        raise ValueError("pos and mass arrays must have matching dtypes")
    return _KDContext(pos, mass, leafsize)


def get_node_count(kd):
    """kdCountNodes (kd.cpp:98-111)."""
    n, l = kd.n, 1
    while n > kd.leafsize:
        n >>= 1
        l <<= 1
    return l << 1


def _create_device_tree(kd):
    L = _lib.lib()
    p, s0, s1, ncols, code, mem, _ = _lib.describe(kd.pos)
    if kd.mass is not None:
        mp, ms0, _, _, _, mmem, _ = _lib.describe(kd.mass)
        if mmem != mem:
            raise ValueError("pos and mass must live in the same memory space")
    else:
        mp, ms0 = None, 0
    h = L.pnb_tree_create(p, s0, s1, mp, ms0, kd.n, code, kd.leafsize, mem)
    if not h:
        msg = L.pnb_last_error().decode()
        if "no CUDA device" in msg or "cuda" in msg.lower():
            raise RuntimeError(msg)
        raise ValueError(msg)
    kd.handle = h


def build(kd, kdnodes, particle_offsets, num_threads=1):
    """build_or_import (kdmain.cpp:250-302) + kdBuildTree (kd.cpp:113-152) on the GPU.

    Fills ``particle_offsets`` (tree order -> file order) and the reference-layout ``kdnodes``; either
    may be ``None`` to skip the host copy (the device tree the kernels use is independent of them).
    """
    L = _lib.lib()
    _create_device_tree(kd)
    if particle_offsets is not None:
        if particle_offsets.dtype != np.intp or not particle_offsets.flags.c_contiguous or len(particle_offsets) != kd.n:
            raise ValueError("particle offsets must be a C-contiguous intp array of length N")
    if kdnodes is not None:
        if kdnodes.dtype.itemsize != KDNode.itemsize or not kdnodes.flags.c_contiguous:
            raise ValueError("kdnodes must be a C-contiguous KDNode array")
        if len(kdnodes) != get_node_count(kd):
            raise ValueError("kdnodes array has the wrong length")
    if kdnodes is not None or particle_offsets is not None:
        export(kd, kdnodes, particle_offsets)


def export(kd, kdnodes, particle_offsets):
    """Fill reference-layout ``kdnodes`` / ``particle_offsets`` (either may be None) from a tree with the
    reference's shape for ``kd.leafsize``, built on the GPU on demand (pnb_tree_export)."""
    _lib.check(_lib.lib().pnb_tree_export(kd.require_built(),
                                          None if kdnodes is None else kdnodes.ctypes.data,
                                          0 if kdnodes is None else len(kdnodes),
                                          None if particle_offsets is None else particle_offsets.ctypes.data))


def import_prebuilt(kd, kdnodes, particle_offsets, _unused=0):
    """kdmain.import_prebuilt: the reference adopts a serialized tree (kdmain.cpp:250-302).

    Neighbour sets do not depend on tree shape and a device build takes milliseconds, so the
    serialized structure is validated and the device tree is rebuilt from the positions.
    """
    if len(kdnodes) != get_node_count(kd):
        raise ValueError("Number of nodes in serialized state does not match number of nodes in kdtree")
    if len(particle_offsets) != kd.n:
        raise ValueError("Number of particle offsets in serialized state does not match number of particles")
    _create_device_tree(kd)


_foreign_free = None     # set by install(): the reference module's ``free`` for trees it built before the swap


def free(kd):
    if kd is not None and not isinstance(kd, _KDContext):
        # a tree of the reference's own extension (a PyCapsule) being collected after install()
        if _foreign_free is not None:
            _foreign_free(kd)
        return
    if kd is not None and kd.handle is not None:
        _lib.lib().pnb_tree_destroy(kd.handle)
        kd.handle = None


def nn_start(kd, nn, period):
    """smInit (smooth.h:71-124, kdmain.cpp:344-382): k > N -> ValueError; period check -> RuntimeWarning."""
    handle = kd.require_built()
    if not isinstance(period, (int, float, np.floating, np.integer)):
        raise SystemError("bad argument to internal function")     # the extension needs a float here
    if int(nn) > kd.n:
        raise ValueError("Number of smoothing particles exceeds number of particles in tree")
    period = float(period)
    if period > 0:
        b = np.empty(6)
        _lib.lib().pnb_tree_bounds(handle, b.ctypes.data)
        if np.any(b[3:] - b[:3] > period):
            warnings.warn(_PERIOD_WARNING, RuntimeWarning)
            period = -1.0
    return _SmoothingContext(nn, period)


def nn_stop(kd, smx):
    smx.nn_result = None
    smx.nn_started = False


def nn_rewind(smx):
    smx.nn_cursor = 0


NN_CHUNK_ROWS = 1 << 16       # rows of the neighbour lists fetched per device round trip by nn_next


def _start_knn(kd, smx):
    """One device search for all particles; the lists stay on the device (pnb_knn_start).  The registered
    smooth array is overwritten with h, as the reference's nn iteration does (smooth.h:429)."""
    L = _lib.lib()
    sm = kd.arrays[0]
    if sm is not None:
        p, s0, s1, ncols, code, mem, _ = _lib.describe(sm)
        _lib.check(L.pnb_set_array(kd.require_built(), 0, p, s0, s1, ncols, code, mem))
    _lib.check(L.pnb_knn_start(kd.require_built(), smx.k, smx.period, None))
    smx.nn_started = True
    smx.nn_result = None
    smx.nn_first = 0


def _fetch_rows(kd, smx, first, count):
    np_dtype = np.float64 if kd.dtype_code == _lib.F64 else np.float32
    idx = np.empty((count, smx.k), np.int64)
    d2 = np.empty((count, smx.k), np_dtype)
    h = np.empty(count, np_dtype)
    _lib.check(_lib.lib().pnb_knn_rows(kd.require_built(), first, count, idx.ctypes.data, d2.ctypes.data, h.ctypes.data))
    return idx, d2, h


def nn_next(kd, smx):
    """typed_nn_next (kdmain.cpp:390-440): [file index, h, [neighbour ids], [d2]] or None when exhausted.

    Particles are yielded in file order; the reference's order (its thread-0 'snake' walk) is an
    artefact no caller relies on.  Lists are unsorted and include the particle itself at d2 = 0.
    Rows come from the device NN_CHUNK_ROWS at a time, so iterating over 10^7 particles never holds
    more than one chunk on the host.
    """
    if not getattr(smx, "nn_started", False):
        _start_knn(kd, smx)
    i = smx.nn_cursor
    if i >= kd.n:
        return None
    if smx.nn_result is None or not (smx.nn_first <= i < smx.nn_first + len(smx.nn_result[2])):
        smx.nn_first = i
        smx.nn_result = _fetch_rows(kd, smx, i, min(NN_CHUNK_ROWS, kd.n - i))
    idx, d2, h = smx.nn_result
    r = i - smx.nn_first
    smx.nn_cursor = i + 1
    return [i, float(h[r]), idx[r].tolist(), d2[r].tolist()]


def all_neighbours(kd, smx):
    """Array form of the full nn() output: (idx[N,k] int64, d2[N,k], h[N]) -- N*k*(8+sizeof T) bytes on the host."""
    _start_knn(kd, smx)
    return _fetch_rows(kd, smx, 0, kd.n)


def gather_counts(kd, smx):
    """Number of particles the reference's gather finds around every particle for the registered smooth array
    (nCnt = smBallGather(4 h^2), kdmain.cpp:779), int32[N] in file order (pnb_gather_counts)."""
    L = _lib.lib()
    sm = kd.arrays[0]
    if sm is None:
        raise ValueError("smooth array has not been set")
    p, s0, s1, ncols, code, mem, _ = _lib.describe(sm)
    _lib.check(L.pnb_set_array(kd.require_built(), 0, p, s0, s1, ncols, code, mem))
    out = np.empty(kd.n, np.int32)
    _lib.check(L.pnb_gather_counts(kd.require_built(), smx.period, out.ctypes.data, _lib.HOST))
    return out


def particles_in_sphere(kd, smx, x, y, z, r):
    """typed_particles_in_sphere (kdmain.cpp:655-679): intp array of file-order ids (ascending)."""
    L = _lib.lib()
    h = kd.require_built()
    found = C.c_int64(0)
    _lib.check(L.pnb_ball_query(h, float(x), float(y), float(z), float(r), smx.period, None, 0, C.byref(found)))
    out = np.empty(found.value, dtype=np.intp)
    if found.value:
        _lib.check(L.pnb_ball_query(h, float(x), float(y), float(z), float(r), smx.period,
                                    out.ctypes.data, found.value, C.byref(found)))
    return out


def set_arrayref(kd, array_id, arr):
    """set_arrayref (kdmain.cpp:511-579): borrow ``arr`` for slot 0 smooth, 1 rho, 2 mass, 3 qty, 4 qty_sm."""
    if array_id not in (0, 1, 2, 3, 4):
        raise ValueError("Unknown array to set for KD tree")
    code = _lib.dtype_code(arr)
    if arr.shape[0] != kd.n:
        raise ValueError("Incorrect array shape")
    if array_id <= 2:
        if len(arr.shape) != 1:
            raise ValueError("Incorrect array shape")
        if code != kd.dtype_code:
            raise TypeError("array dtype must match the dtype of the positions")
    elif len(arr.shape) == 2 and arr.shape[1] != 3 or len(arr.shape) > 2:
        raise ValueError("Incorrect array shape")
    kd.arrays[array_id] = arr


def get_arrayref(kd, array_id):
    return kd.arrays[array_id]


def domain_decomposition(kd, nproc):
    """kdmain.cpp:622-653: the reference validates the smooth array here; the GPU needs no decomposition."""
    if kd.arrays[0] is None:
        raise ValueError("Smoothing array not set")
    if int(nproc) < 1:
        raise ValueError("Invalid number of processors")


def populate(kd, smx, propid, procid, kernel_id=0):
    """typed_populate (kdmain.cpp:681-806).  The reference is called once per host thread with a
    distinct ``procid``; the GPU does the whole pass for ``procid == 0`` and the rest return."""
    if int(procid) != 0:
        return
    L = _lib.lib()
    h = kd.require_built()
    propid = int(propid)
    needed = {1: (0,), 2: (0, 2, 1)}.get(propid, (0, 2, 1, 3, 4))
    for slot in needed:
        arr = kd.arrays[slot]
        if arr is None:
            if slot == 2 and kd.mass is not None and propid != 1:
                arr = kd.mass
            else:
                raise ValueError("%s array has not been set" % ("smooth", "rho", "mass", "qty", "qty_sm")[slot])
        p, s0, s1, ncols, code, mem, _ = _lib.describe(arr)
        if slot == 4 and ncols == 3 and propid in (5, 6, 7):
            # scalar-per-particle results go through the reference's 1-D accessor (kd.h:178-180), i.e.
            # into the first element of each row of a 2-D output (KDTree.sph_dispersion passes one)
            s1, ncols = 0, 1
        _lib.check(L.pnb_set_array(h, slot, p, s0, s1, ncols, code, mem))
    _lib.check(L.pnb_populate(h, propid, smx.k, int(kernel_id), smx.period, None))


def set_option(name, value):
    """Process-wide switches of the library (pnb_set_option): "neighbour_lists" (-1 auto / 0 / 1), "gather_walk" (0 / 1)."""
    _lib.check(_lib.lib().pnb_set_option(name.encode(), int(value)))


def last_device_ms(stage):
    """Device time (ms) of the last build (0) / populate-knn (1) / render (2) call on this thread."""
    return _lib.lib().pnb_last_device_ms(int(stage))


def last_launch_count(stage):
    return _lib.lib().pnb_last_launch_count(int(stage))


def last_kernel_ms(which):
    """Duration (ms) of the last kNN (0) / gather (1) / render (2) kernel launch on this thread."""
    return _lib.lib().pnb_last_kernel_ms(int(which))


def total_launches():
    """Cumulative kernel launches issued through the library by this thread."""
    return _lib.lib().pnb_total_launches()
