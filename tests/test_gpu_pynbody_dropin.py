"""The drop-in claim, exercised on the REAL reference package (pynbody 2.5.0 from baseline/_ref).

The same synthetic gas snapshot is evaluated twice through pynbody's own user-level calls --
``f['smooth']``, ``f['rho']``, ``f['v_mean']``, ``f['v_disp']``, ``f['v_div']``, ``f['v_curl']``,
``pynbody.sph.render_image`` (slice, projected, weighted), ``render_3d_grid``, ``filt.Sphere`` / ``filt.Cuboid``
-- once unpatched (the reference's C++/Cython on the host) and once after ``pynbody_b200.install()``, which swaps
``pynbody.kdtree.kdmain`` / ``KDTree`` and ``pynbody.sph._render``'s three functions for the GPU library.  This runs
pynbody/sph/__init__.py:40-94, derived.py:126-154, snapshot/simsnap.py:1729-1769, sph/renderers.py:650-718 and
filt/__init__.py:244-249,324-337 against libpnb_b200.so with SimArray inputs (units, strided column views).

Bars: smoothing lengths bit-exact; rho / v_* within 1e-5 (float64) or 1e-4 (float32) relative; images (float32)
within 1e-4; selections identical sets.
"""
import numpy as np
import pytest

import refpynbody

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pyn():
    if not refpynbody.available():
        pytest.skip("baseline/_ref (the installed reference) is not present")
    return refpynbody.load()


def _positions(kind, n, dtype, seed):
    from pynbody_b200 import synth
    if kind == "uniform":            # BASELINE.json configs[0]: uniform-random periodic box
        rng = np.random.default_rng(seed)
        pos = rng.uniform(-0.5, 0.5, (n, 3))
        vel = rng.normal(0.0, 1.0, (n, 3)) + 3.0 * np.sin(2 * np.pi * pos[:, [1, 2, 0]])
        mass = rng.uniform(0.5, 1.5, n) / n
    else:                            # configs[1]: clustered Zel'dovich-like box
        side = int(round(n ** (1 / 3)))
        pos, vel, mass = synth.zeldovich_box(side, seed, np.float64)
        pos = pos - 0.5
    return pos.astype(dtype), vel.astype(dtype), mass.astype(dtype)


def _evaluate(pyn, pos, vel, mass, res):
    f = refpynbody.new_gas_snapshot(pyn, pos, vel, mass, boxsize=1.0, temp=1.0 + np.abs(vel[:, 0]))
    out = {}
    for name in ("smooth", "rho", "v_mean", "v_disp", "v_div", "v_curl"):
        out[name] = np.array(f[name])
    assert out["smooth"].dtype == pos.dtype and out["rho"].dtype == pos.dtype
    sph = pyn.sph
    out["slice"] = np.array(sph.render_image(f, quantity="rho", width=1.0, resolution=res, approximate_fast=False))
    out["projected"] = np.array(sph.render_image(f, quantity="rho", width=1.0, resolution=res,
                                                 out_units="Msol kpc^-2", approximate_fast=False))
    out["zoom"] = np.array(sph.render_image(f, quantity="rho", width=0.3, resolution=res, nx=res, ny=res // 2,
                                            out_units="Msol kpc^-2", approximate_fast=False))
    out["weighted"] = np.array(sph.render_image(f, quantity="temp", width=1.0, resolution=res // 2, weight="rho",
                                                approximate_fast=False))
    out["grid"] = np.array(sph.render_3d_grid(f, quantity="rho", width=1.0, nx=48, approximate_fast=False))
    sel = f[pyn.filt.Sphere(0.23, (0.4, -0.1, 0.2))]
    out["sphere"] = np.sort(sel.get_index_list(f))
    sel = f[pyn.filt.Cuboid(-0.1, -0.2, -0.3, 0.2, 0.1, 0.4)]
    out["cuboid"] = np.sort(sel.get_index_list(f))
    out["kdtree_class"] = type(f.kdtree).__module__
    return out


def _compare(ref, got, dtype):
    rtol = 1e-5 if dtype == np.float64 else 1e-4
    assert np.array_equal(ref["smooth"], got["smooth"]), "smoothing lengths are not bit-identical"
    np.testing.assert_allclose(got["rho"], ref["rho"], rtol=rtol)
    for name in ("v_mean", "v_disp", "v_div", "v_curl"):
        scale = np.abs(ref[name]).max()
        np.testing.assert_allclose(got[name], ref[name], rtol=rtol, atol=rtol * scale, err_msg=name)
    for name in ("slice", "projected", "zoom", "weighted", "grid"):
        assert got[name].shape == ref[name].shape and got[name].dtype == ref[name].dtype, name
        np.testing.assert_allclose(got[name], ref[name], rtol=1e-4, atol=1e-6 * np.abs(ref[name]).max(), err_msg=name)
    assert np.array_equal(ref["sphere"], got["sphere"]) and len(ref["sphere"]) > 0
    assert np.array_equal(ref["cuboid"], got["cuboid"]) and len(ref["cuboid"]) > 0


CASES = [
    # (id, kind, n, dtype, k, kernel, image resolution)
    ("C1_uniform_1M_f64", "uniform", 1_000_000, np.float64, 32, "CubicSplineKernel", 256),
    ("C1_uniform_1M_f32", "uniform", 1_000_000, np.float32, 32, "CubicSplineKernel", 256),
    ("C2like_zeldovich_64cubed_f64", "zeldovich", 64 ** 3, np.float64, 64, "WendlandC2Kernel", 200),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_real_pynbody_runs_on_the_gpu_library(pyn, case):
    from pynbody_b200 import _lib, install
    _, kind, n, dtype, k, kernel, res = case
    pos, vel, mass = _positions(kind, n, dtype, 5)
    saved = (pyn.config["sph"]["smooth-particles"], pyn.config["sph"]["kernel"],
             pyn.config["sph"].get("approximate-fast-images"), pyn.config["sph"].get("threaded-image"))
    pyn.config["sph"]["smooth-particles"] = k
    pyn.config["sph"]["kernel"] = kernel
    try:
        ref = _evaluate(pyn, pos, vel, mass, res)
        assert ref["kdtree_class"] == "pynbody.kdtree"
        launches0 = _lib.lib().pnb_total_launches()
        install.install(pyn)
        try:
            got = _evaluate(pyn, pos, vel, mass, res)
        finally:
            install.uninstall(pyn)
        assert got["kdtree_class"] == "pynbody_b200.kdtree"
        assert _lib.lib().pnb_total_launches() > launches0, "the patched run launched no kernel of the library"
        _compare(ref, got, dtype)
    finally:
        (pyn.config["sph"]["smooth-particles"], pyn.config["sph"]["kernel"]) = saved[:2]
        pyn.config["sph"]["approximate-fast-images"], pyn.config["sph"]["threaded-image"] = saved[2:]


def test_reference_kdtree_class_on_our_native_module(pyn):
    """Only ``kdmain`` swapped: the reference's OWN ``KDTree`` host class (pynbody/kdtree/__init__.py:32-538)
    drives the GPU library through the fourteen-function module interface, including ``kdnodes`` /
    ``particle_offsets`` arrays it allocates itself and its Python-thread fan-out of ``populate``."""
    from pynbody_b200 import install
    pos, vel, mass = _positions("uniform", 200_000, np.float64, 9)
    pyn.config["sph"]["smooth-particles"] = 32
    f = refpynbody.new_gas_snapshot(pyn, pos, vel, mass, boxsize=1.0)
    ref = {k: np.array(f[k]) for k in ("smooth", "rho", "v_mean", "v_div")}
    ref_nodes = np.array(f.kdtree.kdnodes)
    del f
    install.install(pyn, replace_kdtree_class=False)
    try:
        g = refpynbody.new_gas_snapshot(pyn, pos, vel, mass, boxsize=1.0)
        got = {k: np.array(g[k]) for k in ("smooth", "rho", "v_mean", "v_div")}
        assert type(g.kdtree).__module__ == "pynbody.kdtree"
        nodes = np.array(g.kdtree.kdnodes)
        offsets = np.array(g.kdtree.particle_offsets)
        near = g.kdtree.particles_in_sphere((0.1, 0.2, 0.3), 0.07)
        del g                        # its tree belongs to the GPU module: release it before the swap back
    finally:
        install.uninstall(pyn)
    assert np.array_equal(ref["smooth"], got["smooth"])
    np.testing.assert_allclose(got["rho"], ref["rho"], rtol=1e-5)
    np.testing.assert_allclose(got["v_mean"], ref["v_mean"], rtol=1e-5, atol=1e-5 * np.abs(ref["v_mean"]).max())
    np.testing.assert_allclose(got["v_div"], ref["v_div"], rtol=1e-5, atol=1e-5 * np.abs(ref["v_div"]).max())
    assert nodes.shape == ref_nodes.shape and np.array_equal(np.sort(offsets), np.arange(len(pos)))
    d = pos - np.array([0.1, 0.2, 0.3])
    d -= np.round(d)
    assert np.array_equal(np.sort(near), np.flatnonzero((d ** 2).sum(1) <= 0.07 ** 2))
