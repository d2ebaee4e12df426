"""Seeded test cases shared by the golden generator, the oracle tests and the GPU parity tests."""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pynbody_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


@dataclass
class SmoothCase:
    gen: str
    n: int
    dtype: type
    boxsize: float | None
    k: int
    kernel_id: int
    seed: int
    leafsize: int = 16
    nn_stride: int = 10
    ball_centre: tuple = (0.9, 0.5, 0.5)
    ball_radius: float = 0.2
    shift: float = 0.0

    def make(self):
        if self.gen == "uniform":
            pos, vel, mass = synth.uniform_box(self.n, self.seed, self.dtype, 1.0)
            rng = np.random.default_rng(self.seed + 100)
            mass = (rng.uniform(0.5, 1.5, self.n) / self.n).astype(self.dtype)
            pos = (pos - self.shift).astype(self.dtype)
        elif self.gen == "blobs":
            pos, vel, mass = synth.gaussian_blobs(self.n, self.seed, self.dtype)
        elif self.gen == "zeldovich":
            pos, vel, mass = synth.zeldovich_box(int(round(self.n ** (1 / 3))), self.seed, self.dtype)
        else:
            raise ValueError(self.gen)
        return pos, vel, mass


SMOOTH_CASES = {
    # C1-like: uniform periodic, k=32, cubic spline, both dtypes
    "uniform_f64": SmoothCase("uniform", 4000, np.float64, 1.0, 32, 0, 11),
    "uniform_f32": SmoothCase("uniform", 4000, np.float32, 1.0, 32, 0, 12),
    # positions in [-0.5, 0.5) as in the reference's particles_in_sphere test
    "uniform_shift_f64": SmoothCase("uniform", 3000, np.float64, 1.0, 16, 1, 13, shift=0.5,
                                    ball_centre=(0.5, 0.0, 0.0), ball_radius=0.3),
    # C2-like: clustered, k=64, WendlandC2
    "zeldovich_f64": SmoothCase("zeldovich", 16 ** 3, np.float64, 1.0, 64, 1, 14),
    "zeldovich_f32": SmoothCase("zeldovich", 16 ** 3, np.float32, 1.0, 64, 1, 15),
    # isolated (non-periodic) clustered set
    "blobs_f64": SmoothCase("blobs", 5000, np.float64, None, 32, 0, 16, ball_centre=(0.0, 0.0, 0.0), ball_radius=0.5),
    "blobs_f32": SmoothCase("blobs", 5000, np.float32, None, 40, 1, 17, ball_centre=(0.0, 0.0, 0.0), ball_radius=0.5),
}


@dataclass
class RenderCase:
    lut: str
    h_power: int
    dtype: type
    nx: int
    ny: int
    nz: int = 0
    grid: bool = False
    z_camera: float = 0.0
    periodic: bool = True
    seed: int = 21
    n: int = 6000
    smooth_lo: float = 0.0
    smooth_hi: float = np.inf
    min_smooth: float = 0.0
    h_scale: float = 0.03
    mixed: bool = False
    extra: dict = field(default_factory=dict)

    def make(self):
        rng = np.random.default_rng(self.seed)
        n = self.n
        pos = rng.uniform(-0.5, 0.5, (n, 3)).astype(self.dtype)
        sm = rng.lognormal(np.log(self.h_scale), 0.9, n).astype(self.dtype)
        mass = (rng.uniform(0.5, 1.5, n) / n).astype(self.dtype)
        rho = rng.uniform(0.5, 2.0, n).astype(self.dtype)
        qty = rng.uniform(0.5, 2.0, n).astype(np.float32 if self.mixed else self.dtype)
        # NaN weights are skipped by render_image but NOT by to_3d_grid (_render.pyx:278 vs :430):
        # give the grid cases a NaN on their most compact particle so it stays local
        qty[int(np.argmin(sm)) if self.grid else 7] = np.nan
        return dict(pos=pos, sm=sm, mass=mass, rho=rho, qty=qty)

    def args(self, a, kern):
        pos = a["pos"]
        x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
        wrap = list(np.linspace(-1.0, 1.0, 3)) if self.periodic else [0.0]
        ar = self.ny / self.nx
        if self.grid:
            return (self.nx, self.ny, self.nz, x, y, z, a["sm"], -0.5, 0.5, -0.5 * ar, 0.5 * ar, -0.5, 0.5,
                    a["qty"], a["mass"], a["rho"], self.smooth_lo, self.smooth_hi, kern, wrap, wrap, wrap)
        return (self.nx, self.ny, x, y, z, a["sm"], -0.5, 0.5, -0.5 * ar, 0.5 * ar, self.z_camera, 0.013,
                a["qty"], a["mass"], a["rho"], self.smooth_lo, self.smooth_hi, -np.inf, np.inf,
                self.min_smooth, kern, wrap, wrap)


RENDER_CASES = {
    "slice_cubic_f64": RenderCase("cubic3d", 3, np.float64, 64, 64),
    "proj_cubic_f64": RenderCase("cubic2d", 2, np.float64, 64, 64),
    "proj_wendland_f32": RenderCase("wendland2d", 2, np.float32, 80, 40, mixed=True),
    "slice_wendland_f32": RenderCase("wendland3d", 3, np.float32, 50, 50, periodic=False),
    "proj_persp_f64": RenderCase("cubic2d", 2, np.float64, 48, 48, z_camera=2.0, periodic=False),
    "proj_range_f64": RenderCase("cubic2d", 2, np.float64, 64, 64, smooth_lo=1.0, smooth_hi=8.0, min_smooth=0.01),
    "proj_big_h_f64": RenderCase("cubic2d", 2, np.float64, 96, 96, h_scale=0.15, n=1500),
    "grid_cubic_f64": RenderCase("cubic3d", 3, np.float64, 16, 16, 16, grid=True),
    "grid_wendland_f32": RenderCase("wendland3d", 3, np.float32, 20, 12, 9, grid=True, periodic=False, mixed=True),
}


@dataclass
class HealpixCase:
    lut: str
    h_power: int
    dtype: type
    nside: int
    shell_radius: float = 0.0
    seed: int = 31
    n: int = 4000
    mixed: bool = False

    def make(self):
        """Particles in a thick shell around the origin (0.6 < r < 1.4) so that angular discs stay well
        below a hemisphere (the reference's fixed-size pixel buffer overflows for near-full-sky discs)."""
        rng = np.random.default_rng(self.seed)
        n = self.n
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        r = rng.uniform(0.6, 1.4, n)
        pos = (d * r[:, None]).astype(self.dtype)
        h = rng.lognormal(np.log(0.05), 0.6, n).astype(self.dtype)
        mass = (rng.uniform(0.5, 1.5, n) / n).astype(self.dtype)
        rho = rng.uniform(0.5, 2.0, n).astype(np.float64 if self.mixed else self.dtype)
        qty = rng.uniform(0.5, 2.0, n).astype(self.dtype)
        qty[11] = np.nan                      # skipped (_render.pyx:128-129)
        return dict(pos=pos, h=h, mass=mass, rho=rho, qty=qty)

    def args(self, a, kern):
        pos = a["pos"]
        return (a["rho"], a["mass"], a["qty"], pos[:, 0], pos[:, 1], pos[:, 2], a["h"], self.nside, kern, self.shell_radius)


HEALPIX_CASES = {
    "proj_cubic_f64": HealpixCase("cubic2d", 2, np.float64, 16),
    "proj_wendland_f32": HealpixCase("wendland2d", 2, np.float32, 32),
    "proj_cubic_f32_mixed": HealpixCase("cubic2d", 2, np.float32, 8, mixed=True),
    "shell_wendland_f64": HealpixCase("wendland3d", 3, np.float64, 16, shell_radius=1.0),
    "shell_cubic_f32": HealpixCase("cubic3d", 3, np.float32, 32, shell_radius=0.9),
}


@dataclass
class SelectCase:
    """filt.geometry_selection.selection: region, parameters, wrap (<= 0: none), dtype."""
    region: str
    params: tuple
    wrap: float
    dtype: type
    n: int = 20000
    seed: int = 51
    scale: float = 1.0          # positions uniform in [-scale/2, scale/2)^3

    def make(self):
        """Uniform positions plus points placed within a few ulps of the region's surface, where only
        the reference's exact operation order gives the reference's answer."""
        rng = np.random.default_rng(self.seed)
        pos = rng.uniform(-0.5 * self.scale, 0.5 * self.scale, (self.n, 3))
        m = self.n // 4
        p = np.array(self.params, dtype=np.float64)
        if self.region == "sphere":
            d = rng.normal(size=(m, 3))
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            r = p[3] * (1.0 + rng.integers(-4, 5, m) * np.finfo(self.dtype).eps)
            pos[:m] = p[:3] + d * r[:, None]
        else:
            lo, hi = p[:3], p[3:]
            pts = rng.uniform(lo, hi, (m, 3))
            ax = rng.integers(0, 3, m)
            side = rng.integers(0, 2, m)
            face = np.where(side == 0, lo[ax], hi[ax])
            pts[np.arange(m), ax] = face * (1.0 + rng.integers(-4, 5, m) * np.finfo(self.dtype).eps)
            pos[:m] = pts
        if self.wrap > 0:       # fold into the box, some points exactly on +-wrap/2
            pos = (pos + 0.5 * self.wrap) % self.wrap - 0.5 * self.wrap
            pos[m:m + 8, 0] = 0.5 * self.wrap
            pos[m + 8:m + 16, 1] = -0.5 * self.wrap
        return np.ascontiguousarray(pos.astype(self.dtype))

    def args(self, pos):
        return (pos, self.region, self.params, self.dtype(self.wrap))


SELECT_CASES = {
    "sphere_f64": SelectCase("sphere", (0.1, -0.05, 0.02, 0.3), -1.0, np.float64),
    "sphere_f32": SelectCase("sphere", (0.1, -0.05, 0.02, 0.3), -1.0, np.float32),
    "sphere_wrap_f64": SelectCase("sphere", (0.45, -0.48, 0.4, 0.25), 1.0, np.float64),
    "sphere_wrap_f32": SelectCase("sphere", (-0.49, 0.3, 0.5, 0.2), 1.0, np.float32),
    "cube_f64": SelectCase("cube", (-0.2, -0.1, -0.3, 0.25, 0.3, 0.1), -1.0, np.float64),
    "cube_f32": SelectCase("cube", (-0.2, -0.1, -0.3, 0.25, 0.3, 0.1), 0.0, np.float32),
    "cube_wrap_f64": SelectCase("cube", (0.3, 0.35, -0.1, 0.7, 0.6, 0.2), 1.0, np.float64),
    "cube_wrap_f32_big": SelectCase("cube", (20.0, -10.0, 30.0, 60.0, 25.0, 70.0), 100.0, np.float32, scale=100.0),
}


def load_luts():
    return dict(np.load(os.path.join(GOLDEN, "luts.npz")))


def load_golden(kind, name):
    return dict(np.load(os.path.join(GOLDEN, f"{kind}_{name}.npz")))
