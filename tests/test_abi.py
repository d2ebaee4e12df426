"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/pnb_b200.h declares, the ctypes signatures cover them, and the host layer keeps the reference's
interface (names, argument validation, error classes) -- no compute calls, no GPU needed."""
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pnb_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pnb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from pynbody_b200 import _lib
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pnb_b200.h but not exported by libpnb_b200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in pynbody_b200/_lib.py"
    assert set(_lib.SIGNATURES) <= set(names)
    assert b"sm_100a" in L.pnb_version()


def test_library_contains_only_sm100a_code():
    """The shipped .so carries sm_100a SASS (no multi-arch fat binary, no PTX-JIT fallback target)."""
    import subprocess
    from pynbody_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_device_means_loud_failure_not_fallback():
    """Without a CUDA device every compute entry point raises; nothing is computed on the CPU."""
    from pynbody_b200 import _lib
    if _lib.lib().pnb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from pynbody_b200.kdtree import KDTree
    from pynbody_b200.sph import _render, kernels
    pos = np.random.default_rng(0).uniform(size=(100, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        KDTree(pos, np.ones(100))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _render.render_image(8, 8, pos[:, 0], pos[:, 1], pos[:, 2], np.full(100, 0.1), 0, 1, 0, 1, 0, 0,
                             np.ones(100), np.ones(100), np.ones(100), 0, np.inf, -np.inf, np.inf, 0,
                             kernels.CubicSplineKernel().projection())


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pynbody_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r"#.*|//.*", "", text), f"{f} mentions the oracle"


def test_kdtree_interface_matches_reference():
    """Same public surface as pynbody.kdtree.KDTree (pynbody/kdtree/__init__.py:32-538)."""
    from pynbody_b200.kdtree import Boundary, KDNode, KDTree, kdmain
    sig = inspect.signature(KDTree.__init__)
    assert list(sig.parameters) == ["self", "pos", "mass", "leafsize", "boxsize", "num_threads", "shared_mem"]
    assert sig.parameters["leafsize"].default == 32 and sig.parameters["boxsize"].default is None
    for m in ("serialize", "deserialize", "particles_in_sphere", "nn", "all_nn", "array_name_to_id", "set_array_ref",
              "get_array_ref", "smooth_operation_to_id", "set_kernel", "populate", "sph_mean", "sph_dispersion",
              "sph_curl", "sph_divergence"):
        assert callable(getattr(KDTree, m)), m
    assert [KDTree.PROPID_HSM, KDTree.PROPID_RHO, KDTree.PROPID_QTYMEAN_1D, KDTree.PROPID_QTYMEAN_ND,
            KDTree.PROPID_QTYDISP_1D, KDTree.PROPID_QTYDISP_ND, KDTree.PROPID_QTYDIV, KDTree.PROPID_QTYCURL] == list(range(1, 9))
    assert KDNode.itemsize == 80 and Boundary.itemsize == 48          # kd.h:29-40
    assert KDNode.names == ("fSplit", "bnd", "iDim", "_pad", "pLower", "pUpper")
    # the native module's method table (kdmain.cpp:63-85)
    for f in ("init", "get_node_count", "build", "import_prebuilt", "free", "nn_start", "nn_next", "nn_stop",
              "nn_rewind", "particles_in_sphere", "set_arrayref", "get_arrayref", "domain_decomposition", "populate"):
        assert callable(getattr(kdmain, f)), f
    assert KDTree.array_name_to_id("smooth") == 0 and KDTree.array_name_to_id("qty_sm") == 4
    with pytest.raises(ValueError):
        KDTree.array_name_to_id("nope")


def test_kdmain_argument_validation_without_device():
    from pynbody_b200.kdtree import kdmain
    rng = np.random.default_rng(1)
    pos = rng.uniform(size=(1000, 3))
    with pytest.raises(ValueError):
        kdmain.init(pos.astype(np.int64), np.ones(1000), 16)           # kdmain.cpp:136-182
    with pytest.raises(ValueError):
        kdmain.init(pos[:, :2], np.ones(1000), 16)
    with pytest.raises(ValueError):
        kdmain.init(pos, np.ones(1000, np.float32), 16)                # kdmain.cpp:204-208
    kd = kdmain.init(pos, np.ones(1000), 16)
    assert kdmain.get_node_count(kd) == 128                            # kd.cpp:98-111: 1000>>6 = 15 <= 16
    kd32 = kdmain.init(pos.astype(np.float32), np.ones(1000, np.float32), 32)
    assert kdmain.get_node_count(kd32) == 64
    with pytest.raises(TypeError):
        kdmain.set_arrayref(kd, 0, np.empty(1000, np.float32))         # kdtree/__init__.py:262-267
    with pytest.raises(ValueError):
        kdmain.set_arrayref(kd, 3, np.empty((1000, 2)))
    with pytest.raises(ValueError):
        kdmain.set_arrayref(kd, 7, np.empty(1000))
    with pytest.raises(ValueError):
        kdmain.domain_decomposition(kd, 1)                             # smooth not set (kdmain.cpp:633-639)
    sm = np.empty(1000)
    kdmain.set_arrayref(kd, 0, sm)
    assert kdmain.get_arrayref(kd, 0) is sm
    with pytest.raises(ValueError):
        kdmain.nn_start(kd, 32, 1.0)                                   # tree not built


def test_render_interface_matches_reference():
    """Positional signatures of _render.pyx:209-223 and :373-385."""
    from pynbody_b200.sph import _render
    p = list(inspect.signature(_render.render_image).parameters)
    assert p == ["nx", "ny", "x", "y", "z", "sm", "x1", "x2", "y1", "y2", "z_camera", "z0", "qty", "mass", "rho",
                 "smooth_lo", "smooth_hi", "z_lo", "z_hi", "min_smooth", "kernel", "wrap_offsets_x", "wrap_offsets_y"]
    p = list(inspect.signature(_render.to_3d_grid).parameters)
    assert p == ["nx", "ny", "nz", "x", "y", "z", "sm", "x1", "x2", "y1", "y2", "z1", "z2", "qty", "mass", "rho",
                 "smooth_lo", "smooth_hi", "kernel", "wrap_offsets_x", "wrap_offsets_y", "wrap_offsets_z"]
    from pynbody_b200.sph import kernels
    with pytest.raises(ValueError, match="3D grid"):
        e = np.empty(0)
        _render.to_3d_grid(4, 4, 4, e, e, e, e, 0, 1, 0, 1, 0, 1, e, e, e, 0.0, np.inf,
                           kernels.CubicSplineKernel().projection())


def test_describe_handles_strides_and_dtypes():
    from pynbody_b200 import _lib
    a = np.zeros((10, 3), np.float32)
    ptr, s0, s1, ncols, code, mem, _ = _lib.describe(a)
    assert (s0, s1, ncols, code, mem) == (12, 4, 3, _lib.F32, _lib.HOST) and ptr == a.ctypes.data
    ptr, s0, s1, ncols, code, mem, _ = _lib.describe(a[:, 1])
    assert (s0, ncols) == (12, 1) and ptr == a.ctypes.data + 4
    with pytest.raises(ValueError):
        _lib.describe(np.zeros(4, np.int32))


def test_null_tree_handle_is_an_error_not_a_crash():
    """Every entry point that takes a tree handle reports PNB_ERR_VALUE for a null handle (no device needed)."""
    from pynbody_b200 import _lib
    L = _lib.lib()
    assert L.pnb_tree_num_particles(None) == _lib.ERR_VALUE
    assert L.pnb_tree_node_count(None) == _lib.ERR_VALUE
    assert L.pnb_tree_num_buckets(None) == _lib.ERR_VALUE
    assert L.pnb_populate(None, 1, 32, 0, 0.0, None) == _lib.ERR_VALUE
    assert b"null" in L.pnb_last_error()
    L.pnb_tree_destroy(None)                      # like free(NULL)
