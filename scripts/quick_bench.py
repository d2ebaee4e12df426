"""Developer timing probe (not the contract bench): device-resident build / kNN / gather times."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pynbody_b200 import synth
from pynbody_b200.kdtree import KDTree, kdmain


def run(gen, n, dtype, k, kernel, reps=2):
    t0 = time.time()
    if gen == "uniform":
        pos, vel, mass = synth.uniform_box(n, 1, dtype)
    else:
        pos, vel, mass = synth.zeldovich_box(int(round(n ** (1 / 3))), 2, dtype)
    tg = time.time() - t0
    tp, tm, tv = (torch.from_numpy(a).cuda() for a in (pos, mass, vel))
    tdt = tp.dtype
    for rep in range(reps):
        torch.cuda.synchronize(); t0 = time.time()
        kd = KDTree(tp, tm, leafsize=16, boxsize=1.0)
        kd.set_kernel(kernel)
        torch.cuda.synchronize(); t_build = time.time() - t0; d_build = kdmain.last_device_ms(0)
        sm = torch.empty(len(pos), dtype=tdt, device="cuda")
        kd.set_array_ref("smooth", sm)
        t0 = time.time(); kd.populate("hsm", k); torch.cuda.synchronize(); t_knn = time.time() - t0; d_knn = kdmain.last_device_ms(1)
        rho = torch.empty_like(sm)
        kd.set_array_ref("mass", tm); kd.set_array_ref("rho", rho)
        t0 = time.time(); kd.populate("rho", k); torch.cuda.synchronize(); t_rho = time.time() - t0; d_rho = kdmain.last_device_ms(1)
        # force the general gather path
        sm2 = sm.clone(); sm2[0] *= 1.0000001
        kd.set_array_ref("smooth", sm2)
        t0 = time.time(); kd.populate("rho", k); torch.cuda.synchronize(); t_rho2 = time.time() - t0; d_rho2 = kdmain.last_device_ms(1)
        kd.set_array_ref("smooth", sm)
        t0 = time.time(); vm = kd.sph_mean(tv, k); torch.cuda.synchronize(); t_mean = time.time() - t0; d_mean = kdmain.last_device_ms(1)
        t0 = time.time(); vd = kd.sph_dispersion(tv, k); torch.cuda.synchronize(); t_disp = time.time() - t0; d_disp = kdmain.last_device_ms(1)
        print(f"{gen} n={n} {np.dtype(dtype).name} k={k} {kernel}: gen {tg:.1f}s | build {t_build*1e3:.1f} ms (dev {d_build:.1f}) | "
              f"knn+rho {t_knn*1e3:.1f} (dev {d_knn:.1f}) | rho(cached) {t_rho*1e3:.1f} | rho(gather) {t_rho2*1e3:.1f} (dev {d_rho2:.1f}) | "
              f"v_mean {t_mean*1e3:.1f} | v_disp {t_disp*1e3:.1f} | "
              f"sim[rho] {n/(t_build+t_knn+t_rho)/1e6:.1f} Mpart/s", flush=True)
        del kd


if __name__ == "__main__":
    big = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    run("uniform", 1_000_000, np.float64, 32, "CubicSplineKernel")
    run("uniform", 1_000_000, np.float32, 32, "CubicSplineKernel")
    run("zeldovich", big ** 3, np.float64, 64, "WendlandC2Kernel")
    run("zeldovich", big ** 3, np.float32, 64, "WendlandC2Kernel")
