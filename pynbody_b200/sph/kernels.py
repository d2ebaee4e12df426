"""SPH kernels on the host side of the render path.

The native renderer takes the kernel *as data*: a 201-entry float32 table of
``W(sqrt(s))`` for ``s = q^2 = 0, 0.02, ..., 4.00`` (h = 1, normalisation
included), exactly what the reference hands to its Cython renderer
(pynbody/sph/kernels.py:30-40; consumed at pynbody/sph/_render.pyx:186-193).
The 2-D (projected) table is the line-of-sight integral of the 3-D kernel
(pynbody/sph/kernels.py:93-111) evaluated with ``scipy.integrate.quad`` as in
the reference.

Class names, ``h_power`` / ``max_d`` / ``get_samples`` / ``projection`` /
``get_c_kernel_id`` and ``create_kernel`` mirror the reference interface
(pynbody/sph/kernels.py:9-151) so either implementation's kernel objects can
be passed to either renderer.
"""
from __future__ import annotations

import math

import numpy as np

_N_SAMPLES_STOP = 4.01
_SAMPLE_STEP = 0.02


class KernelBase:
    """Common behaviour: table caching keyed on the kernel identity."""

    _table_cache: dict = {}

    h_power = 3   #: power of h dividing the tabulated value (3 for 3-D, 2 for projected)
    max_d = 2     #: support radius in units of h

    def _cache_key(self):
        return type(self)

    def __hash__(self):
        return hash(self._cache_key())

    def __eq__(self, other):
        return isinstance(other, KernelBase) and self._cache_key() == other._cache_key()

    def get_value(self, d, h=1):
        raise NotImplementedError

    def get_samples(self, dtype=np.float32):
        key = self._cache_key()
        tab = KernelBase._table_cache.get(key)
        if tab is None:
            s = np.arange(0, _N_SAMPLES_STOP, _SAMPLE_STEP)
            tab = np.array([self.get_value(v ** 0.5) for v in s], dtype=np.float64)
            KernelBase._table_cache[key] = tab
        return tab.astype(dtype)

    def projection(self):
        return Kernel2D(self)

    @classmethod
    def get_c_kernel_id(cls):
        raise NotImplementedError


class CubicSplineKernel(KernelBase):
    """M4 cubic spline, support 2h (pynbody/sph/kernels.py:61-75; native id 0)."""

    def get_value(self, d, h=1):
        if d < 1:
            f = 1. - (3. / 2) * d ** 2 + (3. / 4.) * d ** 3
        elif d < 2:
            f = 0.25 * (2. - d) ** 3
        else:
            f = 0
        return f / (math.pi * h ** 3)

    @classmethod
    def get_c_kernel_id(cls):
        return 0


class WendlandC2Kernel(KernelBase):
    """Wendland C2, support 2h (pynbody/sph/kernels.py:77-90; native id 1)."""

    def get_value(self, d, h=1):
        f = (1. - (d / 2.)) ** 4 * (2. * d + 1) if d < 2 else 0
        return (21. * f) / (16. * math.pi * h ** 3)

    @classmethod
    def get_c_kernel_id(cls):
        return 1


class Kernel2D(KernelBase):
    """Line-of-sight projection of a 3-D kernel (pynbody/sph/kernels.py:93-111)."""

    h_power = 2

    def __init__(self, k_orig=None):
        self.k_orig = k_orig if k_orig is not None else CubicSplineKernel()
        self.max_d = self.k_orig.max_d

    def _cache_key(self):
        return (type(self), self.k_orig._cache_key())

    def projection(self):
        raise ValueError("Cannot project a 2D kernel")

    def get_value(self, d, h=1):
        import scipy.integrate as integrate
        return 2 * integrate.quad(lambda z: self.k_orig.get_value(math.sqrt(z ** 2 + d ** 2), h), 0, 2 * h)[0]

    def get_c_kernel_id(self):
        raise NotImplementedError("2D kernels are not supported in C")


_DEFAULT_KERNEL = "CubicSplineKernel"


def create_kernel(spec=None) -> KernelBase:
    """String / class / instance / None -> kernel object (pynbody/sph/kernels.py:114-151)."""
    if spec is None:
        from ..configuration import config
        return create_kernel(config["sph"].get("kernel", _DEFAULT_KERNEL))
    if isinstance(spec, type):
        return spec()
    if isinstance(spec, KernelBase):
        return spec
    if isinstance(spec, str):
        want = spec.lower()
        for cls in KernelBase.__subclasses__():
            name = cls.__name__.lower()
            if name == want or (name.endswith("kernel") and name[:-6] == want):
                return cls()
        raise ValueError("Unknown kernel '%s'" % spec)
    # duck-typed foreign kernel objects (e.g. pynbody's own) are accepted as they are
    if all(hasattr(spec, a) for a in ("h_power", "max_d", "get_samples")):
        return spec
    raise ValueError("Unknown kernel specification %r" % spec)
