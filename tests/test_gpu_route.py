"""Multi-GPU routing kernels (csrc/route.cu) on one GPU: pnb_route_plan + pnb_route_pack against the eager
formulation they replace (owner by the recorded bisection, stable sort by owner, row gather)."""
import numpy as np
import pytest
import torch

from pynbody_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3, 8, 13])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 1000, 250_007])
def test_route_equals_stable_sort_by_owner(world, dtype, n):
    from pynbody_b200.distributed import Domains, GpuBackend, _bisect, owner_of
    pos, _, mass = synth.clumpy_box(max(n, 8), 5 + world, dtype)
    sample = torch.from_numpy(pos.astype(np.float64))
    doms = Domains([None] * world)
    doms.splits = []
    _bisect(sample.t().contiguous(), None, torch.zeros(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64),
            list(range(world)), doms, doms.splits)
    pos, mass = pos[:n], (mass[:n] * np.arange(1, n + 1)).astype(dtype)
    tp, tm = torch.from_numpy(pos).cuda(), torch.from_numpy(mass).cuda()
    if n > 2:                                              # points exactly on a split plane go to the upper side
        axis, plane = doms.splits[0][0], doms.splits[0][1]
        tp[1, axis] = plane
    packed, src, counts = GpuBackend().route(tp, tm, doms, world)
    owner = owner_of(tp, doms)
    order = torch.sort(owner.to(torch.uint8), stable=True).indices
    assert counts == torch.bincount(owner, minlength=world).tolist() and sum(counts) == n
    assert torch.equal(src, order)
    assert torch.equal(packed[:, :3], tp[order]) and torch.equal(packed[:, 3], tm[order])
    # strided input (a column block of a wider array) routes identically
    wide = torch.zeros((n, 5), dtype=tp.dtype, device="cuda")
    wide[:, 1:4] = tp
    packed2, src2, counts2 = GpuBackend().route(wide[:, 1:4], tm, doms, world)
    assert counts2 == counts and torch.equal(src2, src) and torch.equal(packed2, packed)
