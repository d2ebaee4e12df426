"""Developer probe (not the contract bench): the sim['rho'] step on the C2 box with per-stage device times,
short enough to put under ncu.   python scripts/r2_probe.py [n_side] [f64|f32] [reps] [ops]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pynbody_b200 import synth
from pynbody_b200.kdtree import KDTree, kdmain

side = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dtype = np.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else np.float64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ops = len(sys.argv) > 4 and sys.argv[4] in ("ops", "walk")
walk = len(sys.argv) > 4 and sys.argv[4] == "walk"
k = 64
pos, vel, mass = synth.zeldovich_box_device(side, 2, dtype)
n = pos.shape[0]
sm = torch.empty(n, dtype=pos.dtype, device="cuda")
rho = torch.empty_like(sm)
for rep in range(reps):
    kd = KDTree(pos, mass, leafsize=16, boxsize=1.0)
    kd.set_kernel("WendlandC2Kernel")
    b = kdmain.last_device_ms(0)
    kd.set_array_ref("smooth", sm)
    kd.populate("hsm", k)
    h, hk = kdmain.last_device_ms(1), kdmain.last_kernel_ms(0)
    kd.set_array_ref("mass", mass); kd.set_array_ref("rho", rho)
    kd.populate("rho", k)
    r = kdmain.last_device_ms(1)
    line = f"n={n} {np.dtype(dtype).name} build {b:.2f} | hsm {h:.2f} (kernel {hk:.2f}) | rho {r:.2f}"
    if ops:
        for name, fn in (("mean", kd.sph_mean), ("disp", kd.sph_dispersion), ("div", kd.sph_divergence), ("curl", kd.sph_curl)):
            fn(vel, k)
            line += f" | {name} {kdmain.last_device_ms(1):.2f} (kernel {kdmain.last_kernel_ms(1):.2f})"
    if walk:      # foreign smooth array: the tree-walking fixed-radius gather (gather.cu)
        sm2 = sm.clone(); sm2[0] *= 1.0000001
        kd.set_array_ref("smooth", sm2)
        kd.populate("rho", k)
        line += f" | rho by tree walk {kdmain.last_device_ms(1):.2f} (kernel {kdmain.last_kernel_ms(1):.2f})"
    print(line, flush=True)
    del kd
