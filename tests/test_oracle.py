"""Pins the CPU oracle (oracle/oracle.c) to the reference.

* golden vectors produced by the reference's own compiled modules
  (tests/golden/make_golden.py, run where /root/reference exists);
* the reference's known-answer kernel tables;
* when oracle/_ref is present (it travels with the repo), a live bit-for-bit
  comparison on fresh inputs;
* the reference's own synthetic-data tests (tests/kdtree_test.py:116-121,
  258-282, 322-329 of the reference) restated against the oracle.
"""
import warnings

import numpy as np
import pytest

import cases
from oracle import oracle, ref
from pynbody_b200.sph import kernels


@pytest.fixture(scope="module")
def luts():
    return cases.load_luts()


# ---------------------------------------------------------------- kernel tables
KNOWN = {  # SURVEY.md section 8 row a17 (probe of the reference)
    "cubic3d": ([0.31830987, 0.3094358, 0.30112115], 11.300778),
    "cubic2d": ([0.47746482, 0.46448934, 0.4519395], 16.155329),
    "wendland3d": ([0.41778174, 0.3996931, 0.38374922], 12.147177),
    "wendland2d": ([0.5570423, 0.53696674, 0.51808804], 16.195751),
}


def _my_lut(name):
    k = {"cubic": kernels.CubicSplineKernel, "wendland": kernels.WendlandC2Kernel}[name[:-2]]()
    if name.endswith("2d"):
        k = k.projection()
    return k.get_samples(dtype=np.float32), k


@pytest.mark.parametrize("name", list(KNOWN))
def test_lut_known_answers(name, luts):
    mine, k = _my_lut(name)
    assert mine.shape == (201,) and mine.dtype == np.float32
    first, total = KNOWN[name]
    np.testing.assert_allclose(mine[:3], first, rtol=1e-6)
    np.testing.assert_allclose(mine.sum(dtype=np.float64), total, rtol=1e-6)
    assert mine[-1] == 0
    # bit-identical to the table the reference's kernels.py produced
    assert np.array_equal(mine, luts[name])
    assert k.h_power == (2 if name.endswith("2d") else 3) and k.max_d == 2


def test_create_kernel_specs():
    assert isinstance(kernels.create_kernel("wendlandc2"), kernels.WendlandC2Kernel)
    assert isinstance(kernels.create_kernel("CubicSplineKernel"), kernels.CubicSplineKernel)
    assert isinstance(kernels.create_kernel(kernels.WendlandC2Kernel), kernels.WendlandC2Kernel)
    assert kernels.create_kernel("cubicspline").get_c_kernel_id() == 0
    assert kernels.create_kernel("WendlandC2").get_c_kernel_id() == 1
    with pytest.raises(ValueError):
        kernels.create_kernel("nonsense")
    with pytest.raises(ValueError):
        kernels.CubicSplineKernel().projection().projection()


# ---------------------------------------------------------------- smoothing path vs golden
@pytest.mark.parametrize("name", list(cases.SMOOTH_CASES))
def test_oracle_matches_reference_golden(name):
    case = cases.SMOOTH_CASES[name]
    gold = cases.load_golden("smooth", name)
    pos, vel, mass = case.make()
    kd = oracle.OracleKDTree(pos, mass, case.leafsize, case.boxsize, 2, case.kernel_id)
    sm = kd.smooth(case.k)
    assert np.array_equal(sm, gold["smooth"])           # bit-exact smoothing lengths
    assert np.array_equal(kd.rho(case.k), gold["rho"])
    for op in ("mean", "disp", "div", "curl"):
        assert np.array_equal(kd.qty_op(op, vel, case.k), gold["v_" + op]), op
    assert np.array_equal(kd.qty_op("mean", np.ascontiguousarray(vel[:, 0]), case.k), gold["vx_mean"])
    assert np.array_equal(kd.qty_op("disp", np.ascontiguousarray(vel[:, 0]), case.k), gold["vx_disp"])
    other = np.float32 if vel.dtype == np.float64 else np.float64
    assert np.array_equal(kd.qty_op("mean", vel.astype(other), case.k), gold["v_mean_othertype"])
    # neighbour index sets
    visit, sm2, idx, d2 = kd.all_nn(case.k)
    assert np.array_equal(sm2, sm)
    by_particle = {int(v): i for i, v in enumerate(visit)}
    for i, want, d2max in zip(gold["nn_i"], gold["nn_idx"], gold["nn_d2max"]):
        row = by_particle[int(i)]
        assert np.array_equal(np.sort(idx[row]), want)
        assert d2[row].max() == d2max
        assert int(i) in idx[row] and d2[row][list(idx[row]).index(int(i))] == 0.0   # self at d2 = 0
    assert np.array_equal(np.sort(kd.particles_in_sphere(case.ball_centre, case.ball_radius)), gold["ball"])


@pytest.mark.parametrize("name", list(cases.RENDER_CASES))
def test_render_oracle_matches_reference_golden(name, luts):
    case = cases.RENDER_CASES[name]
    gold = cases.load_golden("render", name)["image"]
    kern = ref.LutKernel(luts[case.lut], case.h_power)
    a = case.make()
    img = oracle.to_3d_grid(*case.args(a, kern)) if case.grid else oracle.render_image(*case.args(a, kern))
    assert img.dtype == np.float32 and img.shape == gold.shape
    assert np.array_equal(img, gold, equal_nan=True)


# ---------------------------------------------------------------- live cross-check with oracle/_ref
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_oracle_vs_live_reference(dtype):
    rng = np.random.default_rng(99)
    n = 6000
    pos = rng.uniform(0, 1, (n, 3)).astype(dtype)
    mass = (rng.uniform(0.5, 1.5, n) / n).astype(dtype)
    vel = rng.normal(size=(n, 3)).astype(dtype)
    a = ref.RefKDTree(pos, mass, 16, 1.0, 2, 1)
    b = oracle.OracleKDTree(pos, mass, 16, 1.0, 2, 1)
    ex = b.export()
    assert np.array_equal(ex["order"], a.particle_offsets)
    assert np.array_equal(ex["lo"], a.kdnodes["pLower"]) and np.array_equal(ex["hi"], a.kdnodes["pUpper"])
    assert np.array_equal(ex["dim"], a.kdnodes["iDim"])
    assert np.array_equal(ex["bmin"], a.kdnodes["bnd"]["fMin"]) and np.array_equal(ex["bmax"], a.kdnodes["bnd"]["fMax"])
    assert np.array_equal(a.smooth(48), b.smooth(48))
    assert np.array_equal(a.rho(48), b.rho(48))
    for op in ("mean", "disp", "div", "curl"):
        assert np.array_equal(a.qty_op(op, vel, 48), b.qty_op(op, vel, 48))


# ---------------------------------------------------------------- reference's own synthetic tests, restated
def test_k_larger_than_n_raises():                       # kdtree_test.py:116-121
    pos = np.random.default_rng(0).uniform(size=(16, 3))
    kd = oracle.OracleKDTree(pos, np.ones(16), 16, None, 1)
    with pytest.raises(ValueError):
        kd.smooth(32)


def test_boxsize_too_small_disables_periodicity():       # kdtree_test.py:322-329
    rng = np.random.default_rng(0)
    pos = rng.normal(scale=1.0, size=(1000, 3))
    kd = oracle.OracleKDTree(pos, np.ones(1000), 16, 0.1, 1)
    kd_open = oracle.OracleKDTree(pos, np.ones(1000), 16, None, 1)
    assert np.array_equal(kd.smooth(32), kd_open.smooth(32))
    assert kd.periodic_honoured is False


@pytest.mark.parametrize("npart", [100, 1000, 100000])
@pytest.mark.parametrize("offset", [0.0, 0.2, 0.5])
@pytest.mark.parametrize("radius", [0.1, 0.3])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_particles_in_sphere_vs_bruteforce(npart, offset, radius, dtype):   # kdtree_test.py:258-282
    np.random.seed(1337)
    pos = np.random.uniform(low=-0.5, high=0.5, size=(npart, 3)).astype(dtype)
    kd = oracle.OracleKDTree(pos, np.ones(npart, dtype), 16, 1.0, 1)
    got = np.sort(kd.particles_in_sphere([offset, 0.0, 0.0], radius))
    p = pos.copy()
    p[:, 0] -= dtype(offset)
    p[p > 0.5] -= 1.0
    p[p < -0.5] += 1.0
    want = np.where(np.sqrt((p.astype(dtype) ** 2).sum(axis=1)) < radius)[0]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name", list(cases.HEALPIX_CASES))
def test_healpix_oracle_matches_reference_golden(name, luts):
    """sph/_render.pyx:51-179 + sph/healpix.c restated: bit-exact against the reference-generated golden."""
    case = cases.HEALPIX_CASES[name]
    gold = cases.load_golden("healpix", name)["image"]
    kern = ref.LutKernel(luts[case.lut], case.h_power)
    img = oracle.render_spherical_image_core(*case.args(case.make(), kern))
    assert img.dtype == np.float32 and img.shape == (12 * case.nside ** 2,)
    assert np.array_equal(img, gold)
    with pytest.raises(ValueError):
        oracle.render_spherical_image_core(*case.args(case.make(), kern)[:7], 12, kern)      # nside not a power of 2


@pytest.mark.parametrize("name", list(cases.SELECT_CASES))
def test_selection_oracle_matches_reference_golden(name):
    """filt/geometry_selection.pyx:31-103 restated: flags bit-identical to the compiled reference's, on
    inputs that include points within a few ulps of the surface and exactly on the wrap boundary."""
    case = cases.SELECT_CASES[name]
    gold = cases.load_golden("select", name)
    flags = oracle.selection(*case.args(case.make()))
    assert flags.dtype == np.uint8 and flags.shape == (case.n,)
    assert int(flags.sum()) == int(gold["count"])
    assert np.array_equal(np.packbits(flags), gold["flags"])
    if ref.available():
        assert np.array_equal(flags, ref.selection(*case.args(case.make())))


def test_selection_oracle_errors_and_empty():
    pos = np.zeros((4, 3))
    with pytest.raises(ValueError, match="4 parameters"):
        oracle.selection(pos, "sphere", (0, 0, 0), -1.0)
    with pytest.raises(ValueError, match="6 parameters"):
        oracle.selection(pos, "cube", (0, 0, 0, 1), -1.0)
    with pytest.raises(ValueError, match="sphere' or 'cube"):
        oracle.selection(pos, "disc", (0, 0, 0, 1), -1.0)
    with pytest.raises(ValueError, match="C-contiguous"):
        oracle.selection(np.zeros((4, 6))[:, ::2], "sphere", (0, 0, 0, 1), -1.0)
    assert oracle.selection(np.zeros((0, 3)), "sphere", (0, 0, 0, 1), -1.0).shape == (0,)


@pytest.mark.skipif(not ref.available(), reason="compiled reference not built")
@pytest.mark.parametrize("seed", range(6))
def test_selection_oracle_equals_compiled_reference_on_random_regions(seed):
    """Random regions, box sizes and dtypes, including regions that straddle the periodic boundary and cuboids
    whose upper corner lies beyond the box (Cuboid._get_boundaries adds one box length, filt/__init__.py:301-310)."""
    rng = np.random.default_rng(100 + seed)
    dtype = (np.float32, np.float64)[seed % 2]
    box = float(rng.choice([1.0, 37.5, 1000.0]))
    pos = np.ascontiguousarray(rng.uniform(-box / 2, box / 2, (5000, 3)).astype(dtype))
    for _ in range(8):
        c = rng.uniform(-box / 2, box / 2, 3)
        r = float(rng.uniform(0.01, 0.6) * box)
        wrap = dtype(box if rng.random() < 0.7 else -1.0)
        assert np.array_equal(oracle.selection(pos, "sphere", (c[0], c[1], c[2], r), wrap),
                              ref.selection(pos, "sphere", (c[0], c[1], c[2], r), wrap))
        lo = rng.uniform(-box / 2, box / 2, 3)
        hi = lo + rng.uniform(0.01, 0.9, 3) * box
        prm = (lo[0], lo[1], lo[2], hi[0], hi[1], hi[2])
        assert np.array_equal(oracle.selection(pos, "cube", prm, wrap), ref.selection(pos, "cube", prm, wrap))
