"""Multi-rank host logic on CPU: world_size 2 and 4 over gloo (no GPU).

The per-rank device tree is replaced by tests/dist_backend.CpuBackend (exact periodic kNN via scipy), so
what is exercised is exactly the product's host code in pynbody_b200/distributed.py: the coordinate
bisection, particle routing (all_to_all), the need-based halo selection and both halo exchanges, the
masked re-search of boundary queries and the routing of results back to their origin.  The bar: the
distributed answers equal the single-rank answers on the same particles.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pynbody_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, gen, n, k, periodic, outdir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from dist_backend import CpuBackend
    from pynbody_b200.distributed import DistributedKDTree
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if gen == "three_clumps":
            # fixed split planes in the voids between the clumps (x = 0.3 and x = 0.7), in the form decompose() returns
            import pynbody_b200.distributed as D
            inf = float("inf")
            doms = D.Domains([(torch.tensor([lo, -inf, -inf], dtype=torch.float64), torch.tensor([hi, inf, inf], dtype=torch.float64))
                              for lo, hi in ((-inf, 0.3), (0.3, 0.7), (0.7, inf))])
            doms.splits = [(0, 0.3, -1, 1), (0, 0.7, -2, -3)]
            D.decompose = lambda *a, **kw: (doms, False)
        pos, vel, mass = _particles(gen, n)
        # every rank starts with an arbitrary (strided) share of the particles, as a file chunk would be
        mine = slice(rank, None, world)
        tp, tm, tv = (torch.from_numpy(np.ascontiguousarray(a[mine])) for a in (pos, mass, vel))
        d = DistributedKDTree(tp, tm, boxsize=1.0 if periodic else None, backend=CpuBackend())
        sm, rho = d.smooth_rho(k, 1)
        vm = d.sph_op("mean", tv, k, 1)
        np.savez(os.path.join(outdir, f"r{rank}.npz"), sm=sm.numpy(), rho=rho.numpy(), vm=vm.numpy(),
                 owned=d.n_own, halo=getattr(d, "n_halo", 0), flagged=d.stats.get("boundary_queries", 0))
    finally:
        dist.destroy_process_group()


def _particles(gen, n):
    if gen == "uniform":
        pos, vel, mass = synth.uniform_box(n, 5, np.float64)
        rng = np.random.default_rng(6)
        mass = rng.uniform(0.5, 1.5, n) / n
    elif gen == "three_clumps":
        # three well separated clumps along x in an open box; the worker pins the split planes into the voids.
        # The outer clumps are tight (no ball reaches a face: zero boundary queries on ranks 0 and 2); the middle
        # one has a few stragglers next to the x = 0.3 plane whose balls cross it
        rng = np.random.default_rng(8)
        pos = np.concatenate([rng.normal(0.0, 0.004, (n // 3, 3)) + np.array([c, 0.5, 0.5]) for c in (0.1, 0.5, 0.9)])
        pos[n // 3: n // 3 + 5] = np.array([0.3001, 0.5, 0.5]) + rng.normal(0.0, 1e-5, (5, 3))
        vel = rng.normal(0.0, 1.0, (n, 3))
        mass = rng.uniform(0.5, 1.5, n) / n
    else:
        pos, vel, mass = synth.clumpy_box(n, 7, np.float64, n_clumps=6)
    return pos, vel, mass


def _single(gen, n, k, periodic):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from dist_backend import CpuLocalTree
    pos, vel, mass = _particles(gen, n)
    t = CpuLocalTree(torch.from_numpy(pos), torch.from_numpy(mass), 1.0 if periodic else None)
    sm = t.knn_smooth(k, None, 1)
    rho = t.density(sm, k, 1)
    vm = t.reduce("mean", sm, rho, torch.from_numpy(vel), k, 1)
    return sm.numpy(), rho.numpy(), vm.numpy()


@pytest.mark.parametrize("world,gen,periodic", [(2, "uniform", True), (4, "clumpy", True), (2, "clumpy", False)])
def test_distributed_equals_single_rank(world, gen, periodic, tmp_path):
    n, k = 3000, 16
    mp.spawn(_worker, args=(world, _free_port(), gen, n, k, periodic, str(tmp_path)), nprocs=world, join=True)
    sm1, rho1, vm1 = _single(gen, n, k, periodic)
    sm = np.empty(n); rho = np.empty(n); vm = np.empty((n, 3))
    owned = halo = flagged = 0
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        sm[r::world], rho[r::world], vm[r::world] = z["sm"], z["rho"], z["vm"]
        owned += int(z["owned"]); halo += int(z["halo"]); flagged += int(z["flagged"])
    assert owned == n                                   # every particle has exactly one owner
    assert 0 < flagged < n and 0 < halo < 2 * n         # a real halo was exchanged, and it is need-based
    assert np.array_equal(sm, sm1), "smoothing lengths must not depend on the decomposition"
    np.testing.assert_allclose(rho, rho1, rtol=1e-12)
    np.testing.assert_allclose(vm, vm1, rtol=1e-9, atol=1e-12)


def test_rank_without_boundary_queries_still_joins_every_collective(tmp_path):
    """Three ranks, one of which has NO boundary query (an isolated clump): the halo-value exchanges of sph_op
    are collectives and every rank must enter them (a rank that skipped them hung the others or mis-paired the
    all-to-all with the next one).  mp.spawn would time out / fail on a mismatch; results must equal one rank's."""
    world, n, k = 3, 3000, 16
    mp.spawn(_worker, args=(world, _free_port(), "three_clumps", n, k, False, str(tmp_path)), nprocs=world, join=True)
    sm1, rho1, vm1 = _single("three_clumps", n, k, False)
    sm = np.empty(n); rho = np.empty(n); vm = np.empty((n, 3))
    flagged = []
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        sm[r::world], rho[r::world], vm[r::world] = z["sm"], z["rho"], z["vm"]
        flagged.append(int(z["flagged"]))
    assert min(flagged) == 0 and max(flagged) > 0, flagged
    assert np.array_equal(sm, sm1)
    np.testing.assert_allclose(rho, rho1, rtol=1e-12)
    np.testing.assert_allclose(vm, vm1, rtol=1e-9, atol=1e-12)


def test_decompose_is_balanced_and_tiles_the_box():
    from pynbody_b200.distributed import decompose, owner_of
    pos, _, _ = synth.clumpy_box(20000, 9, np.float64)
    tp = torch.from_numpy(pos)
    # single process: world = 1 -> one domain covering the periodic box
    doms, periodic = decompose(tp, 1.0)
    assert periodic and len(doms) == 1
    lo, hi = doms[0]
    assert torch.allclose(hi - lo, torch.ones(3, dtype=torch.float64))
    assert int(owner_of(tp, doms).sum()) == 0
    # bisection of a sample into 8 boxes: disjoint, covering, balanced
    from pynbody_b200.distributed import _bisect
    out = [None] * 8
    cols = tp.t().contiguous()                           # [3][m]
    zero, one = torch.zeros(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64)
    _bisect(cols, None, zero, one, list(range(8)), out)
    owner = owner_of(tp, out)
    counts = torch.bincount(owner, minlength=8)
    assert counts.sum() == len(pos) and counts.min() > 0.9 * len(pos) / 8 and counts.max() < 1.1 * len(pos) / 8
    vol = sum(float(torch.prod(h - l)) for l, h in out)
    assert abs(vol - 1.0) < 1e-12
    # cost feedback: particles in the lower half of the box are declared twice as expensive -> the domains
    # there shrink so that every rank gets the same estimated work
    w = np.where(pos[:, 2] < 0.5, 2.0, 1.0)
    out2 = [None] * 8
    _bisect(cols, torch.from_numpy(w), zero, one, list(range(8)), out2)
    owner2 = owner_of(tp, out2)
    work = np.array([w[(owner2 == r).numpy()].sum() for r in range(8)])
    assert work.min() > 0.9 * work.mean() and work.max() < 1.1 * work.mean()
    assert abs(sum(float(torch.prod(h - l)) for l, h in out2) - 1.0) < 1e-12


@pytest.mark.parametrize("world", [2, 3, 5, 8])
def test_owner_by_bisection_walk_equals_owner_by_box_test(world):
    """owner_of walks the recorded bisection when it is available; the result must be the one of the plain
    box test against every rank's cuboid (particles on a split plane go to the upper side in both)."""
    from pynbody_b200.distributed import Domains, _bisect, owner_of
    pos, _, _ = synth.clumpy_box(30000, 13, np.float64)
    tp = torch.from_numpy(pos)
    doms = Domains([None] * world)
    doms.splits = []
    zero, one = torch.zeros(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64)
    _bisect(tp.t().contiguous(), None, zero, one, list(range(world)), doms, doms.splits)
    assert len(doms.splits) == world - 1
    plain = list(doms)                                   # a plain list has no bisection attached
    by_walk, by_box = owner_of(tp, doms), owner_of(tp, plain)
    assert torch.equal(by_walk, by_box)
    # points exactly on a split plane and outside the decomposed box
    axis, plane = doms.splits[0][0], doms.splits[0][1]
    edge = tp[:64].clone()
    edge[:32, axis] = plane
    edge[32:, 0] = 1.5
    assert torch.equal(owner_of(edge, doms), owner_of(edge, plain))
    assert int(torch.bincount(by_walk, minlength=world).min()) > 0
