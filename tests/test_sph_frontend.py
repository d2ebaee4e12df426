"""The array-level front end (pynbody_b200.sph): the host glue of pynbody/sph/renderers.py restated.

CPU part: geometry and periodic tile offsets against the reference's formulas.  GPU part: the front end
gives the same images as calling the native entry point with hand-derived arguments, and the same
smooth / rho as the oracle."""
import numpy as np
import pytest

from pynbody_b200 import sph


def test_wrapping_repeat_array_matches_reference_formula():          # renderers.py:616-629
    assert np.array_equal(sph.wrapping_repeat_array(-0.5, 0.5, None), [0.0])
    assert np.allclose(sph.wrapping_repeat_array(-0.5, 0.5, 1.0), [-1.0, 0.0, 1.0])       # round(0.5) = 0 -> 1 repeat
    assert np.allclose(sph.wrapping_repeat_array(-1.5, 1.5, 1.0), [-3, -2, -1, 0, 1, 2, 3])  # round(1.5) = 2 -> 3 repeats
    assert np.allclose(sph.wrapping_repeat_array(-0.05, 0.05, 1.0), [-1.0, 0.0, 1.0])
    assert np.allclose(sph.wrapping_repeat_array(-10, 10, 4.0), np.linspace(-12, 12, 7))              # round(2.5) = 2 -> 3


def test_image_geometry_matches_reference():                          # renderers.py:45-104, 931-942
    g = sph.ImageGeometry(width=2.0, resolution=100)
    assert (g.x1, g.x2, g.y1, g.y2, g.nx, g.ny, g.nz) == (-1.0, 1.0, -1.0, 1.0, 100, 100, 100)
    assert g.z1 == -np.inf and g.z2 == np.inf and g.z_camera is None and g.width == 2.0
    g.set_pixels(ny=50)
    assert (g.y1, g.y2, g.ny) == (-0.5, 0.5, 50)
    g.restrict_z_range()
    assert (g.z1, g.z2) == (-1.0, 1.0)
    g.set_pixels(nz=25)
    assert (g.z1, g.z2, g.nz) == (-0.25, 0.25, 25)


@pytest.mark.gpu
def test_frontend_images_equal_native_calls():
    from oracle import oracle
    from pynbody_b200.sph import kernels
    rng = np.random.default_rng(3)
    n = 20000
    pos = rng.uniform(-0.5, 0.5, (n, 3))
    sm = rng.lognormal(np.log(0.02), 0.5, n)
    mass = np.full(n, 1.0 / n)
    rho = rng.uniform(0.5, 2.0, n)
    temp = rng.uniform(1.0, 2.0, n)
    k2 = kernels.CubicSplineKernel().projection()
    wrap = [-1.0, 0.0, 1.0]
    img = sph.render_image(pos, mass, rho, sm, temp, width=1.0, nx=128, ny=64, kernel="CubicSplineKernel", boxsize=1.0)
    want = oracle.render_image(128, 64, pos[:, 0], pos[:, 1], pos[:, 2], sm, -0.5, 0.5, -0.25, 0.25, 0.0, 0.0, temp, mass,
                               rho, 0.0, np.inf, -np.inf, np.inf, 0.0, k2, wrap, wrap)
    np.testing.assert_allclose(img, want, rtol=1e-4, atol=1e-6 * want.max())
    # slice through z = 0.1 with the 3-D kernel, open box
    k3 = kernels.WendlandC2Kernel()
    img = sph.render_image(pos, mass, rho, sm, None, width=0.6, resolution=96, kernel=k3, projected=False, z_plane=0.1)
    want = oracle.render_image(96, 96, pos[:, 0], pos[:, 1], pos[:, 2], sm, -0.3, 0.3, -0.3, 0.3, 0.0, 0.1, rho, mass, rho,
                               0.0, np.inf, -np.inf, np.inf, 0.0, k3, [0.0], [0.0])
    np.testing.assert_allclose(img, want, rtol=1e-4, atol=1e-6 * want.max())
    grid = sph.render_3d_grid(pos, mass, rho, sm, width=1.0, resolution=24, kernel=k3, boxsize=1.0)
    want = oracle.to_3d_grid(24, 24, 24, pos[:, 0], pos[:, 1], pos[:, 2], sm, -0.5, 0.5, -0.5, 0.5, -0.5, 0.5, rho, mass, rho,
                             0.0, np.inf, k3, wrap, wrap, wrap)
    np.testing.assert_allclose(grid, want, rtol=1e-4, atol=1e-6 * want.max())


@pytest.mark.gpu
def test_frontend_smooth_rho_equal_oracle():
    from oracle import oracle
    from pynbody_b200 import synth
    pos, vel, mass = synth.zeldovich_box(20, 8, np.float64)
    r, sm = sph.rho(pos, mass, nn=32, boxsize=1.0, kernel="WendlandC2Kernel")
    ref = oracle.OracleKDTree(pos, mass, 16, 1.0, None, 1)
    assert np.array_equal(sm, ref.smooth(32))
    np.testing.assert_allclose(r, ref.rho(32), rtol=1e-5)
    assert np.array_equal(sph.smooth(pos, mass, boxsize=1.0), ref.smooth(32))     # config default: 32 neighbours
