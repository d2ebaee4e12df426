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


def test_multipass_argument_checks():
    """The reference's pipeline logic errors (renderers.py:488-521, 939-944) -- raised before any native call."""
    n = 8
    pos, one = np.zeros((n, 3)), np.ones(n)
    with pytest.raises(sph.RenderPipelineLogicError, match="already projected"):
        sph.render_image(pos, one, one, one, one, width=1.0, resolution=4, projected=True, weight=True)
    with pytest.raises(sph.RenderPipelineLogicError, match="Denoising not supported"):
        sph.render_image(pos, one, one, one, one, width=1.0, resolution=4, projected=True, denoise=True)
    with pytest.raises(sph.RenderPipelineLogicError, match="Denoising not supported"):
        sph.render_image(pos, one, one, one, one, width=1.0, resolution=4, projected=False, weight=True, denoise=True)
    with pytest.raises(ValueError, match="same length"):
        sph.render_image(pos, one, one, one, one, width=1.0, resolution=4, projected=False, weight=np.ones(3))
    with pytest.raises(ValueError, match="nside is only valid"):
        sph.render(pos, one, one, one, target="image", nside=8)
    with pytest.raises(ValueError, match="shell_distance is only valid"):
        sph.render(pos, one, one, one, target="volume", shell_distance=1.0)
    with pytest.raises(ValueError, match="Unknown render target"):
        sph.render(pos, one, one, one, target="cube")


@pytest.mark.gpu
def test_weighted_denoised_and_healpix_pipelines():
    """ProjectionAverageImageRenderer / DenoisedImageRenderer / HealpixRenderer compositions
    (renderers.py:488-556, 720-766) against the same compositions of oracle renders."""
    from oracle import oracle
    from pynbody_b200.sph import kernels
    rng = np.random.default_rng(4)
    n = 20000
    pos = rng.uniform(-0.5, 0.5, (n, 3))
    sm = rng.lognormal(np.log(0.03), 0.4, n)
    mass = np.full(n, 1.0 / n)
    rho = rng.uniform(0.5, 2.0, n)
    temp = rng.uniform(1.0, 2.0, n)
    k3 = kernels.CubicSplineKernel()
    k2 = k3.projection()
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]

    def ora(q, kern, z0=0.0):
        return oracle.render_image(64, 64, x, y, z, sm, -0.5, 0.5, -0.5, 0.5, 0.0, z0, q, mass, rho, 0.0, np.inf, -np.inf,
                                   np.inf, 0.0, kern, [0.0], [0.0])

    # rho-weighted projected average of temp
    img = sph.render_image(pos, mass, rho, sm, temp, width=1.0, resolution=64, kernel=k3, projected=False, weight=rho)
    want = ora(temp * rho, k2) / ora(rho, k2)
    np.testing.assert_allclose(img, want, rtol=2e-4)
    assert 1.0 < np.nanmin(img) and np.nanmax(img) < 2.0
    # volume-weighted
    img = sph.render(pos, mass, rho, sm, temp, target="image", width=1.0, resolution=64, kernel=k3, projected=False, weight=True)
    np.testing.assert_allclose(img, ora(temp, k2) / ora(np.ones(n), k2), rtol=2e-4)
    # denoised slice
    img = sph.render_image(pos, mass, rho, sm, temp, width=1.0, resolution=64, kernel=k3, projected=False, z_plane=0.05,
                           denoise=True)
    np.testing.assert_allclose(img, ora(temp, k3, 0.05) / ora(np.ones(n), k3, 0.05), rtol=2e-4)
    # denoised grid
    grid = sph.render(pos, mass, rho, sm, temp, target="volume", width=1.0, resolution=16, kernel=k3, denoise=True)
    g = lambda q: oracle.to_3d_grid(16, 16, 16, x, y, z, sm, -0.5, 0.5, -0.5, 0.5, -0.5, 0.5, q, mass, rho, 0.0, np.inf, k3,
                                    [0.0], [0.0], [0.0])
    np.testing.assert_allclose(grid, g(temp) / g(np.ones(n)), rtol=2e-4)
    # healpix: particles on a thick shell, projection and thin shell
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    hp_pos = d * rng.uniform(0.7, 1.3, n)[:, None]
    hx, hy, hz = hp_pos[:, 0], hp_pos[:, 1], hp_pos[:, 2]
    got = sph.render(hp_pos, mass, rho, sm, temp, target="healpix", nside=16, kernel=k3)
    want = oracle.render_spherical_image_core(rho, mass, temp, hx, hy, hz, sm, 16, k2)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-6 * want.max())
    got = sph.render_healpix(hp_pos, mass, rho, sm, temp, nside=16, kernel=k3, shell_distance=1.0)
    want = oracle.render_spherical_image_core(rho, mass, temp, hx, hy, hz, sm, 16, k3, 1.0)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-6 * want.max())
    assert sph.render_healpix(hp_pos, mass, rho, sm, nside=None, kernel=k3).shape == (12 * 64 * 64,)


def test_pipeline_compositions_on_cpu_with_the_oracle_as_native_layer(monkeypatch):
    """The decorator logic itself (which passes are made, with which kernel, quantity and geometry) checked on
    CPU: the native entry points are replaced by the oracle's restatements of the reference functions."""
    from oracle import oracle
    from pynbody_b200.sph import _render, kernels
    monkeypatch.setattr(_render, "render_image", oracle.render_image)
    monkeypatch.setattr(_render, "to_3d_grid", oracle.to_3d_grid)
    monkeypatch.setattr(_render, "render_spherical_image_core", oracle.render_spherical_image_core)
    rng = np.random.default_rng(8)
    n = 3000
    pos = rng.uniform(-0.5, 0.5, (n, 3))
    sm = rng.lognormal(np.log(0.05), 0.3, n)
    mass = np.full(n, 1.0 / n)
    rho = rng.uniform(0.5, 2.0, n)
    temp = rng.uniform(1.0, 2.0, n)
    k3 = kernels.CubicSplineKernel()
    k2 = k3.projection()
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]

    def direct(q, kern, z0=0.0, wrap=(0.0,)):
        return oracle.render_image(24, 24, x, y, z, sm, -0.5, 0.5, -0.5, 0.5, 0.0, z0, q, mass, rho, 0.0, np.inf, -np.inf,
                                   np.inf, 0.0, kern, list(wrap), list(wrap))

    with np.errstate(divide="ignore", invalid="ignore"):
        # plain projection and slice, periodic offsets generated as ImageRenderer does
        img = sph.render_image(pos, mass, rho, sm, temp, width=1.0, resolution=24, kernel=k3, boxsize=1.0)
        assert np.array_equal(img, direct(temp, k2, wrap=(-1.0, 0.0, 1.0)))
        img = sph.render_image(pos, mass, rho, sm, None, width=1.0, resolution=24, kernel=k3, projected=False, z_plane=0.1)
        assert np.array_equal(img, direct(rho, k3, z0=0.1))
        # weighted projection = projected(q w) / projected(w); volume weighting uses w = 1
        img = sph.render_image(pos, mass, rho, sm, temp, width=1.0, resolution=24, kernel=k3, projected=False, weight=rho)
        assert np.array_equal(img, direct(temp * rho, k2) / direct(rho, k2), equal_nan=True)
        img = sph.render(pos, mass, rho, sm, temp, target="image", width=1.0, resolution=24, kernel=k3, projected=False,
                         weight=True)
        assert np.array_equal(img, direct(temp, k2) / direct(np.ones(n), k2), equal_nan=True)
        # denoised slice = slice(q) / slice(1)
        img = sph.render_image(pos, mass, rho, sm, temp, width=1.0, resolution=24, kernel=k3, projected=False, denoise=True)
        assert np.array_equal(img, direct(temp, k3) / direct(np.ones(n), k3), equal_nan=True)
        # volume target: cube of side `width`, denoised
        grid = sph.render(pos, mass, rho, sm, temp, target="volume", width=1.0, resolution=8, kernel=k3, denoise=True)
        g = lambda q: oracle.to_3d_grid(8, 8, 8, x, y, z, sm, -0.5, 0.5, -0.5, 0.5, -0.5, 0.5, q, mass, rho, 0.0, np.inf, k3,
                                        [0.0], [0.0], [0.0])
        assert np.array_equal(grid, g(temp) / g(np.ones(n)), equal_nan=True)
    # healpix: projection by default, thin shell with the 3-D kernel when a shell distance is given
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    hp = d * rng.uniform(0.7, 1.3, n)[:, None]
    got = sph.render(hp, mass, rho, sm, temp, target="healpix", nside=8, kernel=k3)
    assert np.array_equal(got, oracle.render_spherical_image_core(rho, mass, temp, hp[:, 0], hp[:, 1], hp[:, 2], sm, 8, k2))
    got = sph.render_healpix(hp, mass, rho, sm, temp, nside=8, kernel=k3, shell_distance=1.0)
    assert np.array_equal(got, oracle.render_spherical_image_core(rho, mass, temp, hp[:, 0], hp[:, 1], hp[:, 2], sm, 8, k3, 1.0))
