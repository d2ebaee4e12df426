"""TEST INFRASTRUCTURE: import the UNMODIFIED reference package (pynbody 2.5.0) from ``baseline/_ref``.

``baseline/_ref`` is produced once by

    cp -r /root/reference /tmp/refbuild/src
    CC=/usr/bin/gcc CXX=/usr/bin/g++ LDSHARED="/usr/bin/gcc -shared" python -m pip install --no-index \
        --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/refbuild/src

(git-ignored; it travels to the GPU box with the repo snapshot).  ``h5py`` is not in the image and several
of the reference's snapshot loaders import it at module scope, so an empty stand-in
module is registered first; nothing on the SPH path touches it.
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "pynbody"))


def load():
    """Return the reference ``pynbody`` module (imported from baseline/_ref) or raise ImportError."""
    if "pynbody" in sys.modules:
        return sys.modules["pynbody"]
    if not available():
        raise ImportError("baseline/_ref/pynbody not installed (see tests/refpynbody.py)")
    try:
        import h5py  # noqa: F401
    except ImportError:
        stub = types.ModuleType("h5py")
        for name in ("File", "Group", "Dataset", "VirtualLayout", "VirtualSource"):
            setattr(stub, name, type(name, (), {}))
        sys.modules["h5py"] = stub
    sys.path.insert(0, REF)
    try:
        import pynbody
    finally:
        sys.path.remove(REF)
    return pynbody


def new_gas_snapshot(pynbody, pos, vel, mass, boxsize=None, temp=None):
    """``pynbody.new(gas=N)`` holding copies of the given arrays IN THEIR OWN DTYPE (kpc, km/s, Msol),
    optionally periodic.  (``pynbody.new`` creates float64 arrays; they are re-created so that float32
    inputs exercise the float32 tree.)"""
    n = len(pos)
    f = pynbody.new(gas=n)
    for name, src, ndim, unit in (("pos", pos, 3, "kpc"), ("vel", vel, 3, "km s^-1"), ("mass", mass, 1, "Msol")):
        del f[name]
        f._create_array(name, ndim, dtype=src.dtype)
        f[name][:] = src
        f[name].units = unit
    if temp is not None:
        f["temp"] = temp
        f["temp"].units = "K"
    if boxsize is not None:
        f.properties["boxsize"] = pynbody.units.Unit("%r kpc" % float(boxsize))
    return f
