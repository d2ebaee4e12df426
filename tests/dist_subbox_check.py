"""Parity of the DISTRIBUTED path at full size against the compiled reference, on a sub-box (launch with
torchrun, one rank per GPU -- this is test infrastructure, which is why it lives under tests/ and may use oracle/):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        tests/dist_subbox_check.py [--n-side 256] [--dtype f64] [--k 64] [--half 0.04]

Every rank holds a Lagrangian slab of ONE global Zel'dovich box of world * n_side^3 particles (134 M at 8 x 256^3, 10^9 at
8 x 500^3).  After DistributedKDTree.smooth_rho the particles inside a cube of half-width ``half + margin`` around a point
where several domains of the decomposition meet are gathered on rank 0 together with the distributed answers; the reference
(oracle/_ref: kd.cpp / smooth.h, non-periodic) is run on that sample alone and must give

* bit-identical smoothing lengths for every particle of the inner cube whose ball (2 h) stays inside the sample -- its
  neighbours are then the same particles in both computations -- and
* densities within 1e-5 (f64) / 1e-4 (f32).

The centre is the corner of rank 0's domain nearest to the middle of the box, so the inner cube holds owned-interior queries,
boundary queries re-searched on the halo tree and halo particles of several ranks.
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
    ap.add_argument("--n-side", type=int, default=256)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--half", type=float, default=0.04)
    args = ap.parse_args()
    from pynbody_b200 import synth
    from pynbody_b200.distributed import DistributedKDTree

    rank, world, local = (int(os.environ.get(v, d)) for v, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype = np.float64 if args.dtype == "f64" else np.float32
    k, kern_id = args.k, 1
    n_side_global = int(round(args.n_side * world ** (1.0 / 3.0)))
    pos, vel, mass = synth.zeldovich_box_device(n_side_global, 11, dtype, "cuda", rank, world)
    # unequal masses: the density then comes from the neighbour lists / the halo tree, not from the fused epilogue alone
    mass = mass * (1.0 + 0.5 * torch.sin(pos[:, 0] * 40.0) ** 2)

    d = DistributedKDTree(pos, mass, boxsize=1.0)
    sm, rho = d.smooth_rho(k, kern_id)

    # a point where domains meet: the corner of rank 0's domain nearest to the middle of the box
    lo0, hi0 = (t.to(torch.float64) for t in d.domains[0])
    mid = torch.full((3,), 0.5, dtype=torch.float64)
    centre = torch.where((lo0 - mid).abs() < (hi0 - mid).abs(), lo0, hi0).clamp(0.1, 0.9)
    cbuf = centre.cuda()
    dist.broadcast(cbuf, src=0)
    centre = cbuf.to(pos.dtype)
    margin = float((2 * sm).max())                     # largest kernel radius anywhere
    mt = torch.tensor([margin], device="cuda", dtype=torch.float64)
    dist.all_reduce(mt, op=dist.ReduceOp.MAX)
    margin = min(float(mt.item()), 0.05)
    outer = args.half + margin
    sel = ((pos - centre).abs() <= outer).all(dim=1)
    mine = [pos[sel], mass[sel], sm[sel], rho[sel]]
    counts = torch.tensor([int(sel.sum())], device="cuda")
    allc = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts)
    allc = [int(c.item()) for c in allc]

    def gather(t):
        if rank == 0:
            parts = [torch.empty((c,) + tuple(t.shape[1:]), dtype=t.dtype, device="cuda") for c in allc]
            parts[0] = t.contiguous()
            for r in range(1, world):
                if allc[r]:
                    dist.recv(parts[r], src=r)
            return torch.cat(parts).cpu().numpy()
        if t.shape[0]:
            dist.send(t.contiguous(), dst=0)
        return None

    g = [gather(t) for t in mine]
    stats = torch.tensor([d.n_own, d.stats.get("boundary_queries", 0), d.stats.get("halo", 0)], device="cuda")
    allstats = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(allstats, stats)
    ok = True
    if rank == 0:
        from oracle import ref
        spos, smass, ssm, srho = (np.ascontiguousarray(a) for a in g)
        t = ref.RefKDTree(spos, smass, 16, None, None, kern_id)       # non-periodic: the sample does not wrap
        rsm = t.smooth(k)
        rrho = t.rho(k)
        c = centre.cpu().numpy()
        dist_in = outer - np.abs(spos - c).max(axis=1)                # distance to the sample's faces
        inner = (np.abs(spos - c) <= args.half).all(axis=1) & (2.0 * rsm < dist_in) & (2.0 * ssm < dist_in)
        nbad = int((ssm[inner] != rsm[inner]).sum())
        rel = float(np.max(np.abs(srho[inner] - rrho[inner]) / rrho[inner])) if inner.any() else float("nan")
        tol = 1e-5 if dtype == np.float64 else 1e-4
        ok = nbad == 0 and inner.sum() > 1000 and rel < tol
        print(f"[dist_subbox_check] world={world} global n={n_side_global ** 3} {args.dtype} k={k}: sample {len(spos)} particles "
              f"around {np.round(c, 4).tolist()} (half {args.half}, margin {margin:.4f}), {int(inner.sum())} compared with the "
              f"compiled reference: smooth mismatches={nbad}, rho max rel err={rel:.2e}; per-rank (owned, boundary queries, halo)="
              f"{[s.tolist() for s in allstats]} -> {'OK' if ok else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
