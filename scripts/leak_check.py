"""Stability probe: repeated tree build / smooth / rho / gather / render / selection cycles with varying particle
counts must not grow device memory without bound (the library caches work buffers by size class) and
pnb_release_cached_memory must hand everything back.  Usage: python scripts/leak_check.py [cycles]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from pynbody_b200 import _lib, synth
from pynbody_b200.filt import geometry_selection
from pynbody_b200.kdtree import KDTree
from pynbody_b200.sph import _render, kernels


def used_mb():
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 2 ** 20


def main():
    cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(0)
    k2 = kernels.create_kernel("CubicSplineKernel").projection()
    torch.zeros(1, device="cuda")
    base = used_mb()
    peak = 0.0
    for c in range(cycles):
        n = int(rng.integers(200_000, 3_000_000))
        dtype = np.float32 if c % 2 else np.float64
        pos, vel, mass = synth.uniform_box(n, c, dtype)
        kd = KDTree(pos, mass, leafsize=16, boxsize=1.0)
        sm = np.empty(n, dtype)
        rho = np.empty(n, dtype)
        kd.set_array_ref("smooth", sm); kd.populate("hsm", 32)
        kd.set_array_ref("mass", mass); kd.set_array_ref("rho", rho); kd.populate("rho", 32)
        kd.sph_mean(vel, 32)
        _render.render_image(256, 256, pos[:, 0], pos[:, 1], pos[:, 2], sm, 0.0, 1.0, 0.0, 1.0, 0.0, 0.5, rho, mass, rho,
                             0.0, np.inf, -np.inf, np.inf, 0.0, k2)
        geometry_selection.selection(pos, "sphere", (0.5, 0.5, 0.5, 0.2), dtype(1.0))
        del kd
        u = used_mb() - base
        peak = max(peak, u)
        if c % 5 == 4:
            print(f"cycle {c + 1:3d}: n = {n:8d}  device memory in use above baseline {u:8.1f} MiB (peak {peak:.1f})", flush=True)
    _lib.lib().pnb_release_cached_memory()
    left = used_mb() - base
    print(f"after pnb_release_cached_memory: {left:.1f} MiB above baseline")
    assert left < 64.0, "device memory not returned"
    print("LEAK CHECK OK")


if __name__ == "__main__":
    main()
