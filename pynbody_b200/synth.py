"""Synthetic particle sets for the workloads BASELINE.json names.

All generators are seeded and return plain numpy arrays (``pos`` N x 3, ``vel``
N x 3, ``mass`` N) in the requested dtype, with positions wrapped into
``[0, boxsize)``.  They are used by the tests, the golden-vector generator and
``bench.py``; nothing here touches the GPU.
"""
from __future__ import annotations

import math

import numpy as np


def uniform_box(n, seed=1, dtype=np.float64, boxsize=1.0):
    """C1: uniform-random periodic box, equal masses (mirrors tests/performance_kdtree.py:32-52)."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0.0, boxsize, (n, 3)).astype(dtype)
    pos[pos >= boxsize] = 0.0  # float32 rounding can land exactly on the box edge
    vel = rng.normal(0.0, 1.0, (n, 3)).astype(dtype)
    mass = np.full(n, 1.0 / n, dtype=dtype)
    return pos, vel, mass


def zeldovich_box(n_side, seed=2, dtype=np.float64, boxsize=1.0, sigma_cells=2.0, jitter=0.0):
    """C2/C5: Zel'dovich-like clustered box.

    Regular ``n_side^3`` lattice displaced by ``psi = grad(phi)`` where ``phi`` is a
    Gaussian random potential whose displacement power spectrum is ``P(k) ~ k^-2``,
    cut at ``k_Nyq / 4``; the amplitude is set so that each component of ``psi``
    has an r.m.s. of ``sigma_cells`` lattice spacings (shell-crossed, strongly
    clustered).  Velocities follow the displacement plus thermal noise.
    """
    rng = np.random.default_rng(seed)
    n = n_side
    k1 = np.fft.fftfreq(n, d=1.0 / n)
    kz1 = np.fft.rfftfreq(n, d=1.0 / n)
    kx, ky, kz = np.meshgrid(k1, k1, kz1, indexing="ij")
    k2 = kx * kx + ky * ky + kz * kz
    k2[0, 0, 0] = 1.0
    kcut = n / 8.0  # k_Nyq / 4
    # displacement spectrum ~ k^-2  =>  potential amplitude ~ k^-2 (psi_k = i k phi_k)
    amp = np.where(k2 <= kcut * kcut, k2 ** -1.0, 0.0)
    amp[0, 0, 0] = 0.0
    phase = rng.normal(size=k2.shape) + 1j * rng.normal(size=k2.shape)
    phik = amp * phase
    psi = []
    for kk in (kx, ky, kz):
        comp = np.fft.irfftn(1j * kk * phik, s=(n, n, n), axes=(0, 1, 2))
        psi.append(comp)
    psi = np.stack([p.ravel() for p in psi], axis=1)
    psi *= sigma_cells / psi.std()
    g = (np.arange(n) + 0.5)
    q = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    if jitter:
        q = q + rng.uniform(-jitter, jitter, q.shape)
    pos = (q + psi) * (boxsize / n)
    pos = np.mod(pos, boxsize)
    vel = psi * 0.5 + rng.normal(0.0, 0.1, psi.shape)
    pos = pos.astype(dtype)
    pos[pos >= boxsize] = 0.0
    mass = np.full(len(pos), 1.0 / len(pos), dtype=dtype)
    return pos, vel.astype(dtype), mass


def clumpy_box(n, seed=4, dtype=np.float32, boxsize=1.0, clump_fraction=0.2, n_clumps=None):
    """C4: uniform background plus ``clump_fraction`` of the particles in Plummer-like clumps."""
    rng = np.random.default_rng(seed)
    n_cl = int(n * clump_fraction)
    n_bg = n - n_cl
    if n_clumps is None:
        n_clumps = max(1, min(10_000, n_cl // 200))
    bg = rng.uniform(0.0, boxsize, (n_bg, 3))
    centres = rng.uniform(0.0, boxsize, (n_clumps, 3))
    scale = boxsize * 0.01 * rng.lognormal(0.0, 0.5, n_clumps)
    which = rng.integers(0, n_clumps, n_cl)
    u = rng.uniform(1e-6, 0.999, n_cl)
    r = scale[which] / np.sqrt(u ** (-2.0 / 3.0) - 1.0)  # Plummer radial CDF inversion
    d = rng.normal(size=(n_cl, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    cl = centres[which] + d * r[:, None]
    pos = np.mod(np.concatenate([bg, cl]), boxsize)
    perm = rng.permutation(n)
    pos = pos[perm].astype(dtype)
    pos[pos >= boxsize] = 0.0
    vel = rng.normal(0.0, 1.0, (n, 3)).astype(dtype)
    mass = np.full(n, 1.0 / n, dtype=dtype)
    return pos, vel, mass


def gaussian_blobs(n, seed=7, dtype=np.float64, n_blobs=5):
    """Small non-periodic clustered set used by parity tests (isolated-halo-like)."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-1.0, 1.0, (n_blobs, 3))
    widths = rng.uniform(0.02, 0.3, n_blobs)
    which = rng.integers(0, n_blobs, n)
    pos = centres[which] + rng.normal(size=(n, 3)) * widths[which][:, None]
    vel = rng.normal(size=(n, 3)) + centres[which] * 3.0
    mass = rng.uniform(0.5, 1.5, n) / n
    return pos.astype(dtype), vel.astype(dtype), mass.astype(dtype)


def zeldovich_box_device(n_side, seed=2, dtype=np.float64, device="cuda", rank=0, world=1, boxsize=1.0,
                         sigma_cells=2.0):
    """The ``zeldovich_box`` recipe evaluated with torch on ``device`` (C4/C5: the global field is too
    large to generate with numpy on every rank).  Every rank evaluates the same global field (same
    seed) and returns its Lagrangian slab ``rank`` of ``world`` -- a share of the particles that is
    *not* a spatial partition, as a file chunk would be.  Returns torch tensors (pos, vel, mass)."""
    import torch
    tdt = torch.float64
    n = n_side
    g = torch.Generator(device=device).manual_seed(int(seed))
    k1 = torch.fft.fftfreq(n, d=1.0 / n, device=device, dtype=tdt)
    kz1 = torch.fft.rfftfreq(n, d=1.0 / n, device=device, dtype=tdt)
    k2 = (k1[:, None, None] ** 2 + k1[None, :, None] ** 2 + kz1[None, None, :] ** 2)
    k2[0, 0, 0] = 1.0
    kcut = n / 8.0
    amp = torch.where(k2 <= kcut * kcut, 1.0 / k2, torch.zeros_like(k2))
    amp[0, 0, 0] = 0.0
    del k2
    phik = torch.complex(torch.randn(amp.shape, generator=g, device=device, dtype=tdt),
                         torch.randn(amp.shape, generator=g, device=device, dtype=tdt)) * amp
    del amp
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    comps, var = [], 0.0
    for axis, kk in enumerate((k1[:, None, None], k1[None, :, None], kz1[None, None, :])):
        comp = torch.fft.irfftn(1j * kk * phik, s=(n, n, n), dim=(0, 1, 2))
        var += float(comp.var())
        comps.append(comp[lo:hi].reshape(-1).clone())
        del comp
    del phik
    psi = torch.stack(comps, dim=1)
    psi *= sigma_cells / math.sqrt(var / 3.0)
    gl = torch.arange(n, device=device, dtype=tdt) + 0.5
    q = torch.stack(torch.meshgrid(gl[lo:hi], gl, gl, indexing="ij"), dim=-1).reshape(-1, 3)
    pos = torch.remainder((q + psi) * (boxsize / n), boxsize)
    vel = psi * 0.5 + 0.1 * torch.randn(psi.shape, generator=g, device=device, dtype=tdt)
    out_dt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    pos = pos.to(out_dt)
    pos[pos >= boxsize] = 0.0
    mass = torch.full((pos.shape[0],), 1.0 / n ** 3, dtype=out_dt, device=device)
    return pos.contiguous(), vel.to(out_dt).contiguous(), mass
