"""pynbody_b200 -- B200-native SPH smoothing path behind pynbody's API.

Drop-in for ``pynbody.kdtree.KDTree`` and ``pynbody.sph._render`` (see
INTEGRATION.md).  The CUDA extension is required: there is no CPU fallback.
"""
__version__ = "0.1.0"
