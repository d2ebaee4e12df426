"""Multi-GPU parity check (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dist_check.py [--n-side 64] [--dtype f64]

Every rank holds a Lagrangian slab of the same global Zel'dovich box.  The distributed path
(pynbody_b200.distributed.DistributedKDTree: domain decomposition, NCCL all-to-all, need-based halo
exchange, masked re-search) must reproduce what ONE GPU computes for the whole box: smoothing lengths
bit-exact, densities / SPH means / divergence to rounding; the reduced image must equal the
single-GPU image.  Rank 0 also holds the whole box to make that comparison.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-side", type=int, default=64)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--k", type=int, default=64)
    args = ap.parse_args()
    from pynbody_b200 import synth
    from pynbody_b200.distributed import DistributedKDTree, GpuLocalTree, reduce_image
    from pynbody_b200.sph import _render, kernels

    rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype = np.float64 if args.dtype == "f64" else np.float32
    k, kern_id = args.k, 1
    pos, vel, mass = synth.zeldovich_box_device(args.n_side, 11, dtype, "cuda", rank, world)

    d = DistributedKDTree(pos, mass, boxsize=1.0)
    sm, rho = d.smooth_rho(k, kern_id)
    vmean = d.sph_op("mean", vel, k, kern_id)
    vdiv = d.sph_op("div", vel, k, kern_id)
    opos, omass, osm, orho = d.owned()
    k2 = kernels.WendlandC2Kernel().projection()
    wrap = [-1.0, 0.0, 1.0]
    img = _render.render_image(256, 256, opos[:, 0], opos[:, 1], opos[:, 2], osm, 0.0, 1.0, 0.0, 1.0, 0.0, 0.5,
                               orho, omass, orho, 0.0, float("inf"), float("-inf"), float("inf"), 0.0, k2, wrap, wrap)
    reduce_image(img)
    stats = torch.tensor([d.n_own, d.stats.get("boundary_queries", 0), d.stats.get("halo", 0)], device="cuda")
    allstats = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(allstats, stats)

    # gather the distributed answers on rank 0 and compare with one GPU doing the whole box
    def gather(t):
        parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t.contiguous(), parts, dst=0)
        return torch.cat(parts) if rank == 0 else None
    g = {name: gather(t) for name, t in (("pos", pos), ("vel", vel), ("mass", mass), ("sm", sm), ("rho", rho),
                                           ("vmean", vmean), ("vdiv", vdiv))}
    ok = True
    if rank == 0:
        t = GpuLocalTree(g["pos"], g["mass"], 1.0)
        sm1 = t.knn_smooth(k, None, kern_id)
        rho1 = t.density(sm1, k, kern_id)
        vm1 = t.reduce("mean", sm1, rho1, g["vel"], k, kern_id)
        vd1 = t.reduce("div", sm1, rho1, g["vel"], k, kern_id)
        img1 = _render.render_image(256, 256, g["pos"][:, 0], g["pos"][:, 1], g["pos"][:, 2], sm1, 0.0, 1.0, 0.0, 1.0, 0.0, 0.5,
                                    rho1, g["mass"], rho1, 0.0, float("inf"), float("-inf"), float("inf"), 0.0, k2, wrap, wrap)
        tol = 1e-5 if dtype == np.float64 else 1e-4
        nbad = int((g["sm"] != sm1).sum())
        def relerr(a, b):
            return float(((a - b).abs() / (b.abs() + 1e-3 * b.abs().max())).max())
        errs = {"rho": relerr(g["rho"], rho1), "v_mean": relerr(g["vmean"], vm1), "v_div": relerr(g["vdiv"], vd1),
                "image": relerr(img, img1)}
        ok = nbad == 0 and all(e < tol for e in errs.values())
        print(f"[dist_check] world={world} n={g['pos'].shape[0]} {args.dtype} k={k}: smooth mismatches={nbad} "
              f"errs={ {k_: f'{v:.2e}' for k_, v in errs.items()} } "
              f"per-rank (owned, boundary queries, halo)={[s.tolist() for s in allstats]} -> {'OK' if ok else 'FAIL'}",
              flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
