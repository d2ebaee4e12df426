"""The multi-rank path with the REAL device kernels, on one GPU: world sizes 2, 3 and 4 over gloo (host tensors, or device
tensors staged through the host for the collectives; every rank's tree / masked searches / ball marking / reductions /
particle routing through the C ABI on cuda:0), compared with the ORACLE on the whole particle set -- not with another
run of the library.

(With NCCL each rank needs its own GPU: scripts/dist_check.py is the same gate for 2 / 4 / 8 GPUs and is run
by hand with gpurun --gpus N; its last outputs are kept under profiles/.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pynbody_b200 import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _particles(n, dtype):
    pos, vel, mass = synth.clumpy_box(n, 19, dtype)
    mass = (mass * np.random.default_rng(2).uniform(0.5, 1.5, n)).astype(dtype)
    return pos, vel, mass


def _worker(rank, world, port, n, k, dtype_name, outdir, on_device):
    from pynbody_b200.distributed import DistributedKDTree, GpuBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pos, vel, mass = _particles(n, np.dtype(dtype_name).type)
        mine = slice(rank, None, world)
        tp, tm, tv = (torch.from_numpy(np.ascontiguousarray(a[mine])) for a in (pos, mass, vel))
        if on_device:          # device tensors: also the library's routing kernels (pnb_route_plan / pnb_route_pack) run
            tp, tm, tv = tp.cuda(), tm.cuda(), tv.cuda()
        d = DistributedKDTree(tp, tm, boxsize=1.0, backend=GpuBackend())
        sm, rho = d.smooth_rho(k, 1)
        out = {"sm": sm.cpu().numpy(), "rho": rho.cpu().numpy()}
        for op in ("mean", "disp", "div", "curl"):
            out[op] = d.sph_op(op, tv, k, 1).cpu().numpy()
        np.savez(os.path.join(outdir, f"r{rank}.npz"), flagged=d.stats.get("boundary_queries", 0),
                 halo=getattr(d, "n_halo", 0), **out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dtype,two_phase,on_device", [(2, np.float64, False, False), (3, np.float32, False, True),
                                                             (2, np.float64, True, False), (4, np.float64, False, True),
                                                             (2, np.float32, True, True)])
def test_distributed_device_path_equals_the_oracle(world, dtype, two_phase, on_device, tmp_path, monkeypatch):
    """two_phase: the opt-in variant that searches near-face particles first and the interior on a helper thread
    (its own CUDA stream) beside the halo pipeline -- multi-pass query masks, pnb_mark_near_faces, per-thread streams."""
    from oracle import oracle
    n, k = 60_000, 32
    if two_phase:
        monkeypatch.setenv("PNB_DIST_TWO_PHASE", "1")
    else:
        monkeypatch.delenv("PNB_DIST_TWO_PHASE", raising=False)
    mp.spawn(_worker, args=(world, _free_port(), n, k, np.dtype(dtype).name, str(tmp_path), on_device), nprocs=world, join=True)
    pos, vel, mass = _particles(n, dtype)
    o = oracle.OracleKDTree(pos, mass, 16, 1.0, None, 1)
    want = {"sm": o.smooth(k)}
    want["rho"] = o.rho(k)
    for op in ("mean", "disp", "div", "curl"):
        want[op] = o.qty_op(op, vel, k)
    got = {name: np.empty_like(v) for name, v in want.items()}
    flagged = halo = 0
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        for name in got:
            got[name][r::world] = z[name]
        flagged += int(z["flagged"]); halo += int(z["halo"])
    assert 0 < flagged < n and halo > 0
    assert np.array_equal(got["sm"], want["sm"]), "smoothing lengths must not depend on the decomposition"
    tol = 1e-5 if dtype == np.float64 else 1e-4
    np.testing.assert_allclose(got["rho"], want["rho"], rtol=tol)
    for op in ("mean", "disp", "div", "curl"):
        scale = np.abs(want[op]).max()
        np.testing.assert_allclose(got[op], want[op], rtol=tol, atol=tol * scale, err_msg=op)
