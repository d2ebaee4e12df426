"""BASELINE.json's configurations at FULL SIZE against the compiled reference (oracle/_ref = the reference's own
kdmain / _render built by oracle/Makefile), through the C ABI.

* C1  10^6 uniform periodic, k = 32, cubic spline, float64 and float32: smooth bit-exact, rho, v_mean, v_disp,
      v_div, v_curl within 1e-5 / 1e-4.
* C2  2^24 clustered (Zel'dovich) periodic box, k = 64, WendlandC2, float64: smooth bit-exact; rho (fused,
      equal masses), rho with unequal masses (neighbour-list path), v_mean, v_disp within 1e-5.
* neighbour INDEX SETS and the gather's neighbour COUNTS (nCnt) at 2*10^5 particles against the oracle, on the
      benchmarked kernel instantiations (float64 k = 64 Wendland; float32 k = 32).
* C3  2048^2 projected image (3 x 3 periodic images) and 256^3 grid of a 2 M-particle clustered box against the
      reference's _render on all host cores.
* the list reductions against the tree-walking gather for every operator and dtype pair.
"""
import os

import numpy as np
import pytest

from pynbody_b200 import synth
from pynbody_b200.kdtree import KDTree, kdmain

pytestmark = pytest.mark.gpu


def _ref():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref (compiled reference) not built")
    return ref


def _rtol(dtype):
    return 1e-5 if dtype == np.float64 else 1e-4


def _close(got, want, dtype, what):
    scale = np.abs(want).max()
    np.testing.assert_allclose(got, want, rtol=_rtol(dtype), atol=_rtol(dtype) * scale * (0 if what == "rho" else 1),
                               err_msg=what)


def _gpu_all(pos, vel, mass, k, kernel, ops=("mean", "disp", "div", "curl")):
    kd = KDTree(pos, mass, leafsize=16, boxsize=1.0)
    kd.set_kernel(kernel)
    n = len(pos)
    sm = np.empty(n, pos.dtype)
    rho = np.empty(n, pos.dtype)
    kd.set_array_ref("smooth", sm)
    kd.populate("hsm", k)
    kd.set_array_ref("mass", mass)
    kd.set_array_ref("rho", rho)
    kd.populate("rho", k)
    out = {"smooth": sm, "rho": rho}
    fn = {"mean": kd.sph_mean, "disp": kd.sph_dispersion, "div": kd.sph_divergence, "curl": kd.sph_curl}
    for op in ops:
        res = fn[op](vel, k)
        out[op] = res[:, 0] if op == "disp" else res          # (N x 3 output, dispersion in column 0)
    return kd, out


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_c1_million_uniform_against_compiled_reference(dtype):
    ref = _ref()
    n, k = 1_000_000, 32
    rng = np.random.default_rng(1)
    pos = rng.uniform(0, 1, (n, 3)).astype(dtype)
    vel = (rng.normal(0, 1, (n, 3)) + np.sin(2 * np.pi * pos[:, [2, 0, 1]])).astype(dtype)
    mass = np.full(n, 1.0 / n, dtype)
    _, got = _gpu_all(pos, vel, mass, k, "CubicSplineKernel")
    r = ref.RefKDTree(pos, mass, 16, 1.0, None, 0)
    assert np.array_equal(got["smooth"], r.smooth(k))
    _close(got["rho"], r.rho(k), dtype, "rho")
    for op in ("mean", "disp", "div", "curl"):
        _close(got[op], r.qty_op(op, vel, k), dtype, op)


def test_c2_sixteen_million_clustered_against_compiled_reference():
    """BASELINE.json configs[1] as bench.py runs it: 256^3 Zel'dovich box, k = 64, WendlandC2, float64."""
    ref = _ref()
    k = 64
    pos, vel, mass = synth.zeldovich_box(256, 2, np.float64)
    assert len(pos) == 1 << 24
    kd, got = _gpu_all(pos, vel, mass, k, "WendlandC2Kernel", ops=("mean", "disp"))
    r = ref.RefKDTree(pos, mass, 16, 1.0, None, 1)
    assert np.array_equal(got["smooth"], r.smooth(k)), "smoothing lengths are not bit-identical at 2^24"
    _close(got["rho"], r.rho(k), np.float64, "rho")
    _close(got["mean"], r.qty_op("mean", vel, k), np.float64, "mean")
    _close(got["disp"], r.qty_op("disp", vel, k), np.float64, "disp")
    # unequal masses: the density comes from the resident neighbour lists instead of the kNN epilogue
    m2 = mass * np.random.default_rng(3).uniform(0.5, 1.5, len(mass))
    rho2 = np.empty_like(mass)
    kd.set_array_ref("mass", m2)
    kd.set_array_ref("rho", rho2)
    kd.populate("rho", k)
    r2 = ref.RefKDTree(pos, m2, 16, 1.0, None, 1)
    _close(rho2, r2.rho(k, smooth=got["smooth"]), np.float64, "rho")


@pytest.mark.parametrize("dtype,k,kernel_id", [(np.float64, 64, 1), (np.float32, 32, 0)])
def test_neighbour_sets_and_gather_counts_at_2e5(dtype, k, kernel_id):
    """Index sets of nn() bit-exact (modulo entries tied with the k-th distance) and nCnt of the gather equal
    to the oracle's, on 200 000 clustered particles -- through the kernel instantiation populate('hsm') uses
    (ids kept in the resident lists), not a separate debugging variant."""
    from oracle import oracle
    n = 200_000
    pos, vel, mass = synth.clumpy_box(n, 31, dtype)
    o = oracle.OracleKDTree(pos, mass, 16, 1.0, None, kernel_id)
    visit, sm_o, idx_o, d2_o = o.all_nn(k)
    kd = KDTree(pos, mass, leafsize=16, boxsize=1.0)
    sm = np.empty(n, dtype)
    kd.set_array_ref("smooth", sm)
    idx, d2, h = kd.all_nn_arrays(k)
    assert np.array_equal(sm, h), "nn() must overwrite the registered smooth array (smooth.h:429)"
    inv = np.empty(n, np.int64)
    inv[visit] = np.arange(n)                       # oracle rows are in its visiting order
    idx_o, d2_o = idx_o[inv], d2_o[inv]           # (the smoothing lengths already are in file order)
    assert np.array_equal(h, sm_o)
    # distances: identical multisets
    assert np.array_equal(np.sort(d2, axis=1), np.sort(d2_o, axis=1))
    # ids: identical sets except where a neighbour is tied with the k-th distance (pq.h:110 keeps whichever came first)
    a, b = np.sort(idx, axis=1), np.sort(idx_o, axis=1)
    differ = np.flatnonzero((a != b).any(axis=1))
    for i in differ:
        dk = d2_o[i].max()
        only = np.setxor1d(idx[i], idx_o[i])
        dd = np.concatenate([d2[i][np.isin(idx[i], only)], d2_o[i][np.isin(idx_o[i], only)]])
        assert np.all(dd == dk), (i, only, dd, dk)
    assert len(differ) < n // 1000
    assert np.all((idx == np.arange(n)[:, None]).sum(axis=1) == 1)       # every list holds the particle itself once
    # chunked access returns the same rows
    smx = kdmain.nn_start(kd.kdtree, k, 1.0)
    kdmain._start_knn(kd.kdtree, smx)
    i2, dd2, h2 = kdmain._fetch_rows(kd.kdtree, smx, 12345, 4096)
    assert np.array_equal(i2, idx[12345:12345 + 4096]) and np.array_equal(dd2, d2[12345:12345 + 4096])
    # neighbour counts of the fixed-radius gather
    kd.set_array_ref("smooth", h)
    cnt = kd.gather_neighbour_counts()
    want = o.gather_counts(sm_o)
    assert set(np.unique(want)) >= {k - 1, k}
    # The reference prunes CELLS in double but measures particles in T (kd.h:68-154 vs smooth.h:300-310): a particle
    # whose rounded d2 equals the ball radius to the last bit can sit in a cell the double test skips (DESIGN.md
    # section 8; checked on the host for this very data set: the particle's float d2 EQUALS fl(4hh) while its exact
    # double distance exceeds it, 21 of 200 000 float32 particles, none in float64).  Such razor-edge cases are the
    # only differences allowed: at most 5e-4 of the particles, off by one, and for each of them OUR count must be
    # the exact evaluation of the reference's acceptance predicate d2 <= fl(4 h h) over ALL particles (brute force
    # in T arithmetic).
    from test_gpu_kdtree import _bruteforce_d2
    odd = np.flatnonzero(cnt != want)
    assert len(odd) <= n // 2000, (len(odd), cnt[odd][:10], want[odd][:10])
    for i in odd:
        assert cnt[i] == want[i] + 1
        ball2 = dtype(4) * h[i] * h[i]
        assert int((_bruteforce_d2(pos[i], pos, 1.0) <= ball2).sum()) == cnt[i], i
    for i in np.random.default_rng(0).integers(0, n, 50):
        ball2 = dtype(4) * h[i] * h[i]
        assert int((_bruteforce_d2(pos[i], pos, 1.0) <= ball2).sum()) == cnt[i], i


@pytest.mark.parametrize("tdtype,qdtype", [(np.float64, np.float64), (np.float32, np.float32),
                                           (np.float64, np.float32), (np.float32, np.float64)])
def test_list_reductions_equal_the_tree_walk(tdtype, qdtype):
    """Every operator over the resident neighbour lists vs the tree-walking fixed-radius gather (the formulation
    of the reference) on clustered, unequal-mass particles: to rounding."""
    n, k = 150_000, 40
    pos, vel, mass = synth.clumpy_box(n, 7, tdtype)
    mass = (mass * np.random.default_rng(1).uniform(0.5, 1.5, n)).astype(tdtype)
    vel = vel.astype(qdtype)
    res = {}
    for walk in (0, 1):
        kdmain.set_option("gather_walk", walk)
        try:
            for kern in ("CubicSplineKernel", "WendlandC2Kernel"):
                kd, out = _gpu_all(pos, vel, mass, k, kern)
                out["mean1"] = kd.sph_mean(vel[:, 0].copy(), k)
                out["disp1"] = kd.sph_dispersion(vel[:, 1].copy(), k)
                res[walk, kern] = out
        finally:
            kdmain.set_option("gather_walk", 0)
    tol = 1e-12 if tdtype == np.float64 and qdtype == np.float64 else 2e-5
    for kern in ("CubicSplineKernel", "WendlandC2Kernel"):
        a, b = res[0, kern], res[1, kern]
        assert np.array_equal(a["smooth"], b["smooth"])
        for name in ("rho", "mean", "disp", "div", "curl", "mean1", "disp1"):
            scale = np.abs(b[name]).max()
            np.testing.assert_allclose(a[name], b[name], rtol=tol, atol=tol * scale, err_msg=f"{kern} {name}")


def test_c3_2048_image_and_256_grid_against_compiled_reference():
    """BASELINE.json configs[2] geometry (2048^2 projected density with 3 x 3 periodic images; 256^3 grid) on a
    128^3 = 2.1 M-particle clustered box whose smoothing lengths and densities come from the GPU path."""
    ref = _ref()
    from pynbody_b200.sph import _render, kernels
    k = 64
    pos, vel, mass = synth.zeldovich_box(128, 2, np.float64)
    pos = pos - 0.5
    _, got = _gpu_all(pos, vel, mass, k, "WendlandC2Kernel", ops=())
    sm, rho = got["smooth"], got["rho"]
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
    wrap = [-1.0, 0.0, 1.0]
    threads = os.cpu_count() or 1
    kern2 = kernels.WendlandC2Kernel().projection()
    args = (2048, 2048, x, y, z, sm, -0.5, 0.5, -0.5, 0.5, 0.0, 0.0, rho, mass, rho,
            0.0, np.inf, -np.inf, np.inf, 0.0, kern2, wrap, wrap)
    img = _render.render_image(*args)
    want = ref.render_image(*args, num_threads=threads)
    assert img.shape == (2048, 2048) and img.dtype == np.float32
    np.testing.assert_allclose(img, want, rtol=1e-4, atol=1e-6 * want.max())
    # mass conservation: integral of the projected image = total mass (the box is fully covered) up to the
    # renderer's 201-entry nearest-lower kernel table (_render.pyx:186-193; the reference's image integrates
    # to the same 1.0125)
    assert abs(img.astype(np.float64).sum() / 2048 ** 2 - mass.sum()) < 0.02 * mass.sum()
    assert abs(img.astype(np.float64).sum() - want.astype(np.float64).sum()) < 1e-5 * want.astype(np.float64).sum()
    kern3 = kernels.WendlandC2Kernel()
    gargs = (256, 256, 256, x, y, z, sm, -0.5, 0.5, -0.5, 0.5, -0.5, 0.5, rho, mass, rho, 0.0, np.inf, kern3,
             wrap, wrap, wrap)
    grid = _render.to_3d_grid(*gargs)
    gwant = ref.to_3d_grid(*gargs, num_threads=threads)
    assert grid.shape == (256, 256, 256)
    np.testing.assert_allclose(grid, gwant, rtol=1e-4, atol=1e-6 * gwant.max())
