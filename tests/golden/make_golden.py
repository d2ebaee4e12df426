"""Regenerates the committed golden vectors in tests/golden/ from the REFERENCE ITSELF.

Run in the build container (needs /root/reference for the kernel tables and the
compiled reference modules in oracle/_ref, see oracle/Makefile):

    make -C oracle ref && python tests/golden/make_golden.py

Inputs are regenerated from seeds by tests/cases.py, so only outputs are stored.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref  # noqa: E402
import cases  # noqa: E402


def reference_kernels():
    spec = importlib.util.spec_from_file_location("ref_kernels", "/root/reference/pynbody/sph/kernels.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    rk = reference_kernels()
    luts = {
        "cubic3d": rk.CubicSplineKernel().get_samples(),
        "cubic2d": rk.CubicSplineKernel().projection().get_samples(),
        "wendland3d": rk.WendlandC2Kernel().get_samples(),
        "wendland2d": rk.WendlandC2Kernel().projection().get_samples(),
    }
    np.savez(os.path.join(HERE, "luts.npz"), **luts)

    for name, case in cases.SMOOTH_CASES.items():
        pos, vel, mass = case.make()
        kd = ref.RefKDTree(pos, mass, case.leafsize, case.boxsize, 4, case.kernel_id)
        out = {"smooth": kd.smooth(case.k), "rho": None}
        out["rho"] = kd.rho(case.k)
        for op in ("mean", "disp", "div", "curl"):
            out["v_" + op] = kd.qty_op(op, vel, case.k)
        out["vx_mean"] = kd.qty_op("mean", np.ascontiguousarray(vel[:, 0]), case.k)
        out["vx_disp"] = kd.qty_op("disp", np.ascontiguousarray(vel[:, 0]), case.k)
        other = np.float32 if vel.dtype == np.float64 else np.float64
        out["v_mean_othertype"] = kd.qty_op("mean", vel.astype(other), case.k)
        nn = kd.all_nn(case.k)
        sel = sorted(nn, key=lambda r: r[0])[::case.nn_stride]
        out["nn_i"] = np.array([r[0] for r in sel], np.int64)
        out["nn_idx"] = np.array([sorted(r[2]) for r in sel], np.int32)
        out["nn_d2max"] = np.array([max(r[3]) for r in sel], pos.dtype)
        out["ball"] = np.sort(kd.particles_in_sphere(case.ball_centre, case.ball_radius)).astype(np.int32)
        np.savez_compressed(os.path.join(HERE, f"smooth_{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})

    for name, case in cases.RENDER_CASES.items():
        a = case.make()
        kern = ref.LutKernel(luts[case.lut], case.h_power)
        if case.grid:
            img = ref.to_3d_grid(*case.args(a, kern))
        else:
            img = ref.render_image(*case.args(a, kern))
        np.savez_compressed(os.path.join(HERE, f"render_{name}.npz"), image=img)
        print(name, img.shape, float(np.nansum(img)))

    for name, case in cases.HEALPIX_CASES.items():
        a = case.make()
        kern = ref.LutKernel(luts[case.lut], case.h_power)
        img = ref.render_spherical_image_core(*case.args(a, kern))
        np.savez_compressed(os.path.join(HERE, f"healpix_{name}.npz"), image=img)
        print(name, img.shape, float(np.nansum(img)))

    for name, case in cases.SELECT_CASES.items():
        flags = ref.selection(*case.args(case.make()))
        np.savez_compressed(os.path.join(HERE, f"select_{name}.npz"), flags=np.packbits(flags), count=int(flags.sum()))
        print(name, flags.shape, int(flags.sum()))


if __name__ == "__main__":
    main()
