"""GPU drop-in for ``pynbody.kdtree`` (pynbody/kdtree/__init__.py).

``KDTree`` keeps the reference's constructor, attributes and public methods
(pynbody/kdtree/__init__.py:79-128, 136-163, 167-231, 249-375, 377-538) and drives
``pynbody_b200.kdtree.kdmain`` -- the drop-in for the native module -- in the same
sequence the reference drives its C++ extension.  Threads are not used: one call
runs the whole pass on the GPU.
"""
from __future__ import annotations

import logging
import time
import warnings

import numpy as np

from ..configuration import config
from . import kdmain
from .kdmain import Boundary, KDNode  # noqa: F401  (re-exported like the reference)

logger = logging.getLogger("pynbody_b200.kdtree")


class KDTree:
    """KDTree for smoothing, interpolation and geometrical queries, on the GPU."""

    PROPID_HSM = 1
    PROPID_RHO = 2
    PROPID_QTYMEAN_1D = 3
    PROPID_QTYMEAN_ND = 4
    PROPID_QTYDISP_1D = 5
    PROPID_QTYDISP_ND = 6
    PROPID_QTYDIV = 7
    PROPID_QTYCURL = 8

    def __init__(self, pos, mass, leafsize=32, boxsize=None, num_threads=None, shared_mem=False):
        """Create a KDTree (same signature as the reference, pynbody/kdtree/__init__.py:79).

        ``num_threads`` is accepted for compatibility and ignored.  ``shared_mem`` trees (a
        cross-process host feature of the reference) are not supported.
        """
        if shared_mem:
            raise NotImplementedError("shared_mem trees are a host-side feature outside the GPU path")
        self._set_num_threads(num_threads)
        self.leafsize = int(leafsize)
        self.kdtree = kdmain.init(pos, mass, self.leafsize)
        self._kdnodes = None
        self._particle_offsets = None
        kdmain.build(self.kdtree, None, None, 1)
        self.boxsize = boxsize
        self._pos = pos
        self.set_kernel(_config_kernel())

    # -- reference-layout views, materialised together on first access ------------------------
    def _materialise(self):
        if self._kdnodes is None or self._particle_offsets is None:
            nodes = np.zeros(kdmain.get_node_count(self.kdtree), dtype=KDNode)
            offsets = np.empty(self.kdtree.n, dtype=np.intp)
            kdmain.export(self.kdtree, nodes, offsets)
            self._kdnodes, self._particle_offsets = nodes, offsets

    @property
    def particle_offsets(self):
        self._materialise()
        return self._particle_offsets

    @property
    def kdnodes(self):
        self._materialise()
        return self._kdnodes

    def _set_num_threads(self, num_threads):
        if num_threads is None:
            num_threads = int(config["number_of_threads"])
        self.num_threads = num_threads
        return num_threads

    def serialize(self):
        """Produce a serialized description of the tree (pynbody/kdtree/__init__.py:136-138)."""
        return self.leafsize, self.boxsize, self.kdnodes, self.particle_offsets, self._kernel_id

    @classmethod
    def deserialize(cls, pos, mass, serialized_state, num_threads=None, boxsize=None):
        """A KDTree for ``pos`` from what :meth:`serialize` returned (pynbody/kdtree/__init__.py:140-163).  The
        serialized node records are checked against the particle set and kept as this tree's ``kdnodes`` /
        ``particle_offsets``; the device tree itself is rebuilt from the positions (milliseconds; neighbour sets do not
        depend on the tree's shape)."""
        leafsize, saved_boxsize, kdnodes, particle_offsets, kernel_id = serialized_state
        if None not in (boxsize, saved_boxsize) and abs(boxsize - saved_boxsize) > 1e-6:
            warnings.warn("Boxsize in serialized state does not match boxsize passed to deserialize")
        self = object.__new__(cls)
        self._set_num_threads(num_threads)
        self.leafsize = leafsize
        self.kdtree = kdmain.init(pos, mass, leafsize)
        kdmain.import_prebuilt(self.kdtree, kdnodes, particle_offsets, 0)     # ValueError on a size mismatch
        self._kdnodes, self._particle_offsets = kdnodes, particle_offsets
        self.boxsize, self._pos, self._kernel_id = boxsize, pos, kernel_id
        return self

    def _period(self):
        return -1.0 if self.boxsize is None else float(self.boxsize)

    def particles_in_sphere(self, center, radius):
        """Indices of the particles within ``radius`` of ``center`` (pynbody/kdtree/__init__.py:167-189)."""
        smx = kdmain.nn_start(self.kdtree, 1, self._period())
        try:
            return kdmain.particles_in_sphere(self.kdtree, smx, center[0], center[1], center[2], radius)
        finally:
            kdmain.nn_stop(self.kdtree, smx)

    def nn(self, nn=None):
        """Generator of ``[i, h_i, [neighbour ids], [d2]]`` (pynbody/kdtree/__init__.py:191-227)."""
        if nn is None:
            nn = 64
        smx = kdmain.nn_start(self.kdtree, int(nn), self._period())
        kdmain.domain_decomposition(self.kdtree, 1)
        while True:
            nbr_list = kdmain.nn_next(self.kdtree, smx)
            if nbr_list is None:
                break
            yield nbr_list
        kdmain.nn_stop(self.kdtree, smx)

    def all_nn(self, nn=None):
        """A list of neighbours information for all the particles."""
        return [x for x in self.nn(nn)]

    def all_nn_arrays(self, nn=None):
        """Array form of :meth:`all_nn`: ``(idx[N,k], d2[N,k], h[N])`` in file order (GPU-side addition)."""
        if nn is None:
            nn = 64
        smx = kdmain.nn_start(self.kdtree, int(nn), self._period())
        try:
            return kdmain.all_neighbours(self.kdtree, smx)
        finally:
            smx.nn_result = None

    def gather_neighbour_counts(self):
        """How many particles the reference's fixed-radius gather (d2 <= 4 h^2) finds around each particle for
        the registered smooth array: int32[N] (GPU-side addition; the reference keeps nCnt internal)."""
        smx = kdmain.nn_start(self.kdtree, 1, self._period())
        try:
            return kdmain.gather_counts(self.kdtree, smx)
        finally:
            kdmain.nn_stop(self.kdtree, smx)

    @staticmethod
    def array_name_to_id(name):
        try:
            return {"smooth": 0, "rho": 1, "mass": 2, "qty": 3, "qty_sm": 4}[name]
        except KeyError:
            raise ValueError("Unknown KDTree array") from None

    def set_array_ref(self, name, ar):
        """Give the native code access to an array (pynbody/kdtree/__init__.py:249-269)."""
        if self.array_name_to_id(name) < 3:
            if _np_dtype(self._pos) != _np_dtype(ar):
                raise TypeError("KDTree requires matching dtypes for %s (%s) and pos (%s) arrays"
                                % (name, ar.dtype, self._pos.dtype))
        kdmain.set_arrayref(self.kdtree, self.array_name_to_id(name), ar)
        assert self.get_array_ref(name) is ar

    def get_array_ref(self, name):
        return kdmain.get_arrayref(self.kdtree, self.array_name_to_id(name))

    # operation name -> (property id for 1-D input, property id for N x 3 input); None = shape not accepted
    _OPERATIONS = {"hsm": (PROPID_HSM, PROPID_HSM), "rho": (PROPID_RHO, PROPID_RHO),
                   "qty_mean": (PROPID_QTYMEAN_1D, PROPID_QTYMEAN_ND), "qty_disp": (PROPID_QTYDISP_1D, PROPID_QTYDISP_ND),
                   "qty_div": (None, PROPID_QTYDIV), "qty_curl": (None, PROPID_QTYCURL)}

    def smooth_operation_to_id(self, name):
        """Property id of a smoothing operation, chosen by the shape of the registered ``qty`` array
        (pynbody/kdtree/__init__.py:275-320: same names, same ValueErrors)."""
        if name not in self._OPERATIONS:
            raise ValueError("Unknown smoothing request %s" % name)
        one, nd = self._OPERATIONS[name]
        if one == nd:
            return one
        shape = tuple(self.get_array_ref("qty").shape)
        is_nd = len(shape) == 2 and shape[1] == 3
        if name in ("qty_div", "qty_curl"):
            if not is_nd:
                raise ValueError("Can only compute %s of 3D arrays" % ("divergence" if name == "qty_div" else "curl"))
            return nd
        if len(shape) == 1:
            return one
        if not is_nd:
            raise ValueError("Currently only able to smooth 3D or 1D arrays")
        return nd

    def set_kernel(self, kernel=None):
        """Set the kernel for smoothing operations (pynbody/kdtree/__init__.py:322-334)."""
        from ..sph import kernels
        self._kernel_id = kernels.create_kernel(kernel).get_c_kernel_id()

    def populate(self, mode, nn):
        """Run the operation ``mode`` with ``nn`` neighbours (pynbody/kdtree/__init__.py:338-375)."""
        if nn is None:
            nn = 64
        smx = kdmain.nn_start(self.kdtree, int(nn), self._period())
        try:
            propid = self.smooth_operation_to_id(mode)
            if propid == self.PROPID_HSM:
                kdmain.domain_decomposition(self.kdtree, 1)
            kdmain.populate(self.kdtree, smx, propid, 0, self._kernel_id)
        finally:
            kdmain.nn_stop(self.kdtree, smx)

    def _output_like(self, array, shape=None, units=None):
        if kdmain._lib._is_torch(array):
            import torch
            return torch.empty(shape if shape is not None else tuple(array.shape), dtype=array.dtype, device=array.device)
        out = np.empty_like(array) if shape is None else np.empty(shape, dtype=array.dtype)
        if hasattr(array, "units"):
            out = out.view(type(array))
            out.units = array.units if units is None else units
        return out

    def _sph_reduce(self, array, mode, label, nsmooth, scalar_out=False, per_length=False):
        """One SPH reduction of ``array`` over each particle's neighbours: registers input and output, runs
        ``populate(mode)`` and returns the output (pynbody/kdtree/__init__.py:377-534).  ``scalar_out``: one value per
        particle; ``per_length``: the result carries the units of ``array`` divided by those of the positions."""
        units = None
        if per_length and hasattr(array, "units") and hasattr(self._pos, "units"):
            units = array.units / self._pos.units
        output = self._output_like(array, shape=(len(array),) if scalar_out else None, units=units)
        self.set_array_ref("qty", array)
        self.set_array_ref("qty_sm", output)
        logger.info("%s of array with %d nearest neighbours" % (label, nsmooth))
        start = time.time()
        self.populate(mode, nsmooth)
        logger.info("%s done in %5.3g s" % (label, time.time() - start))
        return output

    def sph_mean(self, array, nsmooth=64):
        """SPH mean of a 1-D or N x 3 array (pynbody/kdtree/__init__.py:377-419)."""
        return self._sph_reduce(array, "qty_mean", "SPH mean", nsmooth)

    def sph_dispersion(self, array, nsmooth=64):
        """SPH dispersion (pynbody/kdtree/__init__.py:421-470).  As in the reference the output has the input's shape;
        for N x 3 input the single dispersion per particle lands in column 0 (smDispQtyND writes through a 1-D accessor)."""
        return self._sph_reduce(array, "qty_disp", "SPH dispersion", nsmooth)

    def sph_curl(self, array, nsmooth=64):
        """SPH curl of an N x 3 array (pynbody/kdtree/__init__.py:472-534)."""
        return self._sph_reduce(array, "qty_curl", "SPH curl", nsmooth, per_length=True)

    def sph_divergence(self, array, nsmooth=64):
        """SPH divergence of an N x 3 array (pynbody/kdtree/__init__.py:472-534)."""
        return self._sph_reduce(array, "qty_div", "SPH divergence", nsmooth, scalar_out=True, per_length=True)

    def __del__(self):
        if hasattr(self, "kdtree"):
            try:
                kdmain.free(self.kdtree)
            except Exception:
                pass


def _np_dtype(a):
    from .. import _lib
    return _lib.dtype_code(a)


def _config_kernel():
    try:
        return config["sph"].get("kernel", "CubicSplineKernel")
    except Exception:
        return "CubicSplineKernel"
