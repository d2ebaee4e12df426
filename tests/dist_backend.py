"""TEST INFRASTRUCTURE: a CPU stand-in for the per-rank device tree (pynbody_b200.distributed.GpuLocalTree).

Lets the multi-rank host logic (domain decomposition, halo selection and exchange, result routing) run
under gloo without a GPU.  Exact brute-force-equivalent answers via scipy's periodic cKDTree; never
used by the product path.
"""
import numpy as np
import torch
from scipy.spatial import cKDTree


def _w(kernel_id, k, q2):
    q2 = np.asarray(q2, dtype=np.float64)
    if kernel_id == 0:
        q = np.sqrt(q2)
        t = 2.0 - q
        return np.where(t < 0, 0.0, np.where(q2 < 1.0, 1.0 - 0.75 * t * q2, 0.25 * t ** 3))
    au = np.sqrt(q2 * 0.25)
    w = (21 / 16.0) * (1 - au) ** 4 * (1 + 4 * au)
    w = np.where(q2 > 4.0, 0.0, w)
    return np.where(q2 == 0.0, (21 / 16.0) * (1 - 0.0294 * (k * 0.01) ** -0.977), w)


class CpuLocalTree:
    def __init__(self, pos, mass, boxsize):
        self.pos = pos.numpy().astype(np.float64)
        self.mass = mass.numpy().astype(np.float64)
        self.n = len(self.pos)
        self.boxsize = boxsize
        self.dtype = pos.dtype
        self.tree = cKDTree(self.pos, boxsize=boxsize) if self.n else None

    def _sel(self, mask):
        return np.arange(self.n) if mask is None else np.nonzero(mask.numpy())[0]

    def knn_smooth(self, k, mask=None, kernel_id=0):
        sel = self._sel(mask)
        sm = np.zeros(self.n)
        if len(sel):
            d, _ = self.tree.query(self.pos[sel], k)
            d = d.reshape(len(sel), -1)
            sm[sel] = 0.5 * d[:, -1]
        return torch.from_numpy(sm).to(self.dtype)

    def _neighbours(self, i, h):
        idx = np.array(self.tree.query_ball_point(self.pos[i], 2 * h * (1 + 1e-12)), dtype=np.int64)
        d = self.pos[idx] - self.pos[i]
        if self.boxsize:
            d -= self.boxsize * np.round(d / self.boxsize)
        return idx, (d * d).sum(axis=1)

    def density(self, sm, k, kernel_id, mask=None):
        sm = sm.numpy().astype(np.float64)
        rho = np.zeros(self.n)
        for i in self._sel(mask):
            idx, d2 = self._neighbours(i, sm[i])
            order = np.argsort(idx)
            idx, d2 = idx[order], d2[order]
            rho[i] = np.sum(self.mass[idx] * _w(kernel_id, k, d2 / sm[i] ** 2)) / (np.pi * sm[i] ** 3)
        return torch.from_numpy(rho).to(self.dtype)

    def reduce(self, op, sm, rho, qty, k, kernel_id, mask=None):
        assert op == "mean"
        sm = sm.numpy().astype(np.float64)
        rho = rho.numpy().astype(np.float64)
        q = qty.numpy().astype(np.float64)
        out = np.zeros_like(q)
        for i in self._sel(mask):
            idx, d2 = self._neighbours(i, sm[i])
            order = np.argsort(idx)
            idx, d2 = idx[order], d2[order]
            w = self.mass[idx] / rho[idx] * _w(kernel_id, k, d2 / sm[i] ** 2) / (np.pi * sm[i] ** 3)
            out[i] = (w[:, None] * q[idx]).sum(axis=0) if q.ndim == 2 else np.sum(w * q[idx])
        return torch.from_numpy(out).to(qty.dtype)

    def mark_in_balls(self, centres, radii):
        flags = np.zeros(self.n, dtype=bool)
        c = centres.numpy().astype(np.float64)
        r = radii.numpy().astype(np.float64)
        if self.boxsize:
            c = np.mod(c, self.boxsize)
        for ci, ri in zip(c, r):
            flags[self.tree.query_ball_point(ci, min(ri, 1e9))] = True
        return torch.from_numpy(flags)


class CpuBackend:
    def make_tree(self, pos, mass, boxsize):
        return CpuLocalTree(pos, mass, boxsize)
