"""Developer probe: the C3 image (2048^2 projected, 3x3 periodic images) of the C2 box, device-resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pynbody_b200 import synth
from pynbody_b200.kdtree import KDTree, kdmain
from pynbody_b200.sph import _render, kernels
side = int(sys.argv[1]) if len(sys.argv) > 1 else 256
res = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pos, vel, mass = synth.zeldovich_box_device(side, 2, np.float64)
n = pos.shape[0]
kd = KDTree(pos, mass, leafsize=16, boxsize=1.0); kd.set_kernel("WendlandC2Kernel")
sm = torch.empty(n, dtype=pos.dtype, device="cuda"); rho = torch.empty_like(sm)
kd.set_array_ref("smooth", sm); kd.populate("hsm", 64)
kd.set_array_ref("mass", mass); kd.set_array_ref("rho", rho); kd.populate("rho", 64)
k2 = kernels.WendlandC2Kernel().projection(); k3 = kernels.WendlandC2Kernel()
wrap = [-1.0, 0.0, 1.0]
for r in range(reps):
    img = _render.render_image(res, res, pos[:, 0], pos[:, 1], pos[:, 2], sm, 0.0, 1.0, 0.0, 1.0, 0.0, 0.5, rho, mass, rho,
                               0.0, float("inf"), float("-inf"), float("inf"), 0.0, k2, wrap, wrap)
    print("image ms (call, tile kernel):", _render.last_render_ms(), flush=True)
if len(sys.argv) > 4:
    g = _render.to_3d_grid(256, 256, 256, pos[:, 0], pos[:, 1], pos[:, 2], sm, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, rho, mass, rho,
                           0.0, float("inf"), k3, wrap, wrap, wrap)
    print("grid ms:", _render.last_render_ms(), flush=True)
    c = pos - 0.5
    r2 = (c * c).sum(dim=1)
    shell = torch.nonzero((r2 > 0.09) & (r2 < 0.25)).squeeze(1)
    hp = _render.render_spherical_image_core(rho[shell], mass[shell], rho[shell], c[shell][:, 0].contiguous(), c[shell][:, 1].contiguous(),
                                             c[shell][:, 2].contiguous(), sm[shell], 256, k2)
    print("healpix nside 256,", int(shell.numel()), "particles, ms:", _render.last_render_ms(), flush=True)
