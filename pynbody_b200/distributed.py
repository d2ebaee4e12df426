"""Multi-GPU host layer: spatial domain decomposition, halo exchange, image reduction.

The reference is single-process; this module is what lets the same path run on
the 8 B200s of one box, one process per GPU, with ``torch.distributed`` (NCCL over NVLink) for the
exchange steps.  Per rank the work is the single-GPU path of ``libpnb_b200.so``; the only collectives
are where the algorithm has a real exchange:

  1. ``decompose``     recursive coordinate bisection of the (periodic) box into one cuboid domain per
                       rank, split planes from an all-gathered coordinate sample (equal particle counts);
  2. ``_alltoallv``    particles move to the rank that owns their domain (all_to_all_single);
  3. smoothing pass    every rank searches its OWN particles first.  That already is the exact answer
                       for every particle whose k-neighbour ball lies inside the domain, and an upper
                       bound r_ub on the ball radius for the rest ("boundary queries");
  4. halo exchange     boundary queries (centre, r_ub) are sent to the ranks whose domain their ball
                       touches; those ranks mark exactly the particles inside any received ball
                       (``pnb_mark_in_balls``) and send them back -- a need-based halo, not a fixed-width
                       shell, so void particles with large h do not inflate it;
  5. re-search         boundary queries only (``pnb_set_query_mask``) on owned + halo particles;
  6. later passes      (mean / dispersion / div / curl need rho and the quantity of halo particles):
                       the same send lists carry those per-particle values (halo exchange #2);
  7. results return to the rank and position they came from (inverse all_to_all);
  8. images: every rank renders its own particles into a full frame, ``reduce_image`` sums the frames.

All tensor plumbing is torch (CPU tensors + gloo in the tests, CUDA tensors + NCCL in production); the
local compute goes through a backend object -- ``GpuBackend`` (the C ABI) unless a test injects another.
"""
from __future__ import annotations

import math
import os
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

OPS = {"mean": "qty_mean", "disp": "qty_disp", "div": "qty_div", "curl": "qty_curl"}


# ------------------------------------------------------------------------------------------------
# local compute backend: the single-GPU C-ABI path
# ------------------------------------------------------------------------------------------------
class GpuLocalTree:
    """One rank's device tree over (owned [+ halo]) particles; thin wrapper over the C ABI."""

    def __init__(self, pos, mass, boxsize):
        from .kdtree import KDTree
        self.pos, self.mass = pos.contiguous(), mass.contiguous()
        self.n = int(pos.shape[0])
        self.kd = KDTree(self.pos, self.mass, leafsize=32, boxsize=boxsize)
        self.boxsize = boxsize

    def _mask(self, mask, invert=False):
        """Restrict the next pass to a subset.  Results of earlier passes stay valid (pnb_set_query_mask2 with
        keep_results): this layer only ever asks a tree about particles it has searched on that tree."""
        from . import _lib
        h = self.kd.kdtree.require_built()
        if mask is None:
            _lib.check(_lib.lib().pnb_set_query_mask2(h, None, _lib.DEVICE, 0, 1))
        else:
            m = mask.to(torch.uint8).contiguous()
            _lib.check(_lib.lib().pnb_set_query_mask2(h, m.data_ptr(), _lib.DEVICE if m.is_cuda else _lib.HOST,
                                                      1 if invert else 0, 1))

    def mark_near_faces(self, k, lo, hi):
        """bool[n]: particles whose k-neighbour ball might reach a face of the cuboid [lo, hi) (pnb_mark_near_faces)."""
        from . import _lib
        flags = torch.zeros(self.n, dtype=torch.uint8, device=self.pos.device)
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        _lib.check(_lib.lib().pnb_mark_near_faces(self.kd.kdtree.require_built(), int(k), lo.ctypes.data, hi.ctypes.data,
                                                  flags.data_ptr(), _lib.DEVICE if flags.is_cuda else _lib.HOST))
        return flags.bool()

    def knn_smooth(self, k, mask=None, kernel_id=0, invert=False):
        self.kd._kernel_id = int(kernel_id)
        self._mask(mask, invert)
        sm = torch.empty(self.n, dtype=self.pos.dtype, device=self.pos.device)
        self.kd.set_array_ref("smooth", sm)
        self.kd.populate("hsm", k)
        return sm

    def last_knn_ms(self):
        from .kdtree import kdmain
        return kdmain.last_kernel_ms(0)

    def density(self, sm, k, kernel_id, mask=None):
        self.kd._kernel_id = int(kernel_id)
        self._mask(mask)
        rho = torch.empty(self.n, dtype=self.pos.dtype, device=self.pos.device)
        self.kd.set_array_ref("smooth", sm)
        self.kd.set_array_ref("mass", self.mass)
        self.kd.set_array_ref("rho", rho)
        self.kd.populate("rho", k)
        return rho

    def reduce(self, op, sm, rho, qty, k, kernel_id, mask=None):
        self.kd._kernel_id = int(kernel_id)
        self._mask(mask)
        out_shape = tuple(qty.shape) if op in ("mean", "curl") else (self.n,)
        out = torch.empty(out_shape, dtype=qty.dtype, device=qty.device)
        self.kd.set_array_ref("smooth", sm)
        self.kd.set_array_ref("mass", self.mass)
        self.kd.set_array_ref("rho", rho)
        self.kd.set_array_ref("qty", qty.contiguous())
        self.kd.set_array_ref("qty_sm", out)
        self.kd.populate(OPS[op], k)
        return out

    def mark_in_balls(self, centres, radii):
        from . import _lib
        flags = torch.zeros(self.n, dtype=torch.uint8, device=self.pos.device)
        nq = int(centres.shape[0])
        if nq == 0:
            return flags.bool()
        c = centres.to(self.pos.dtype).contiguous()
        r = radii.to(self.pos.dtype).contiguous()
        es = c.element_size()
        period = -1.0 if self.boxsize is None else float(self.boxsize)
        _lib.check(_lib.lib().pnb_mark_in_balls(self.kd.kdtree.require_built(), c.data_ptr(), 3 * es, es, r.data_ptr(), es,
                                                nq, period, flags.data_ptr(), _lib.DEVICE if c.is_cuda else _lib.HOST))
        return flags.bool()


class GpuBackend:
    native_routing = True      # owner assignment and send-buffer packing by the library's kernels (csrc/route.cu)
    two_phase = True           # near-face queries first; the interior search then overlaps the halo pipeline

    def make_tree(self, pos, mass, boxsize):
        return GpuLocalTree(pos, mass, boxsize)

    def route(self, pos, mass, domains, world):
        """(packed [n][4] rows {x,y,z,m} destination-major and stable, original index of every row, rows per
        destination) -- pnb_route_plan + pnb_route_pack."""
        from . import _lib
        L = _lib.lib()
        n = int(pos.shape[0])
        es = pos.element_size()
        code = _lib.dtype_code(pos)
        splits = domains.splits or []
        axis = np.array([sp[0] for sp in splits], np.int32)
        plane = np.array([sp[1] for sp in splits], np.float64)
        child = np.array([[sp[2], sp[3]] for sp in splits], np.int32).reshape(-1)
        owner = torch.empty(n, dtype=torch.uint8, device=pos.device)
        counts = np.zeros(world, np.int64)
        _lib.check(L.pnb_route_plan(pos.data_ptr(), pos.stride(0) * es, pos.stride(1) * es, n, code,
                                    axis.ctypes.data, plane.ctypes.data, child.ctypes.data, len(splits), world,
                                    owner.data_ptr(), counts.ctypes.data))
        packed = torch.empty((n, 4), dtype=pos.dtype, device=pos.device)
        src = torch.empty(n, dtype=torch.int64, device=pos.device)
        m = mass.to(pos.dtype)
        _lib.check(L.pnb_route_pack(pos.data_ptr(), pos.stride(0) * es, pos.stride(1) * es, m.data_ptr(), m.stride(0) * es,
                                    n, code, owner.data_ptr(), world, packed.data_ptr(), src.data_ptr()))
        return packed, src, [int(c) for c in counts]


# ------------------------------------------------------------------------------------------------
# domain decomposition
# ------------------------------------------------------------------------------------------------
class Domains(list):
    """One (lo[3], hi[3]) cuboid per rank, plus the bisection that produced them: ``splits`` is a list of
    (axis, plane, left, right) with children >= 0 pointing at another split and -(rank + 1) at a leaf."""
    splits = None


def _bisect(sample, weight, lo, hi, ranks, out, splits=None):
    """Recursive coordinate bisection of box [lo,hi) among `ranks`: split planes at the (weighted) quantiles
    of the coordinate sample.  ``sample`` is a [3][m] tensor (one contiguous row per axis) that stays on the
    device it was gathered on -- only the chosen axis and plane of each node come back to the host;
    ``weight`` = estimated search cost per sample point, or None.  lo / hi are host tensors."""
    if len(ranks) == 1:
        out[ranks[0]] = (lo.clone(), hi.clone())
        return -(ranks[0] + 1)
    me = None
    if splits is not None:
        me = len(splits)
        splits.append(None)
    nleft = len(ranks) // 2
    m = int(sample.shape[1])
    if m > 1:
        ext = (sample.max(dim=1).values - sample.min(dim=1).values).tolist()
        axis = max(range(3), key=lambda a: ext[a])
        col = sample[axis]
        if weight is None:                                  # equal weights: selection, not sorting
            kth = min(m - 1, max(1, int(round(m * nleft / len(ranks)))))
            q = torch.kthvalue(col, kth + 1).values
        else:
            order = torch.argsort(col)
            cw = torch.cumsum(weight[order], 0)
            i = int(torch.searchsorted(cw, cw[-1] * nleft / len(ranks)))
            q = col[order[min(max(i, 1), m - 1)]]
        left = col < q
        q = q.to(lo.device, lo.dtype)
    else:
        ext = (hi - lo).tolist()
        axis = max(range(3), key=lambda a: ext[a])
        mid = 0.5 * (lo[axis] + hi[axis])
        q = mid if bool(torch.isfinite(mid)) else torch.zeros((), dtype=lo.dtype)
        left = sample[axis] < q.to(sample.device, sample.dtype)
    hi_l, lo_r = hi.clone(), lo.clone()
    hi_l[axis] = q
    lo_r[axis] = q
    cl = _bisect(sample[:, left], None if weight is None else weight[left], lo, hi_l, ranks[:nleft], out, splits)
    cr = _bisect(sample[:, ~left], None if weight is None else weight[~left], lo_r, hi, ranks[nleft:], out, splits)
    if splits is not None:
        splits[me] = (axis, float(q), cl, cr)
    return me


# Optional cost feedback (PNB_DIST_COST_FEEDBACK=1): after a smoothing pass the ranks share how long their
# search took per particle and the next decomposition weights its coordinate sample with that
# piecewise-constant cost field, as N-body codes do with the previous step's measured costs.  Off by
# default: on the clustered boxes measured here the per-particle search cost of equal-count domains
# agrees to ~2 % (4.28-4.39 ms per million particles at 4 GPUs), so there is nothing to gain, and split
# planes that move from call to call defeat the re-use of work buffers.  Purely a performance hint: any
# set of split planes gives the same results.
_COST_FIELD = {}


def record_costs(world, domains, cost_per_particle):
    _COST_FIELD[world] = ([(l.numpy().copy(), h.numpy().copy()) for l, h in domains], np.asarray(cost_per_particle, float))


def _sample_weights(world, samp):
    """samp: [3][m] tensor; returns a weight per sample point or None."""
    if world not in _COST_FIELD or not samp.shape[1] or not os.environ.get("PNB_DIST_COST_FEEDBACK"):
        return None
    w = torch.ones(samp.shape[1], dtype=torch.float64, device=samp.device)
    doms, cost = _COST_FIELD[world]
    if np.all(np.isfinite(cost)) and cost.min() > 0:
        cost = np.clip(cost / cost.mean(), 0.5, 2.0)
        for (lo, hi), c in zip(doms, cost):
            lo_t = torch.as_tensor(lo, device=samp.device, dtype=samp.dtype)[:, None]
            hi_t = torch.as_tensor(hi, device=samp.device, dtype=samp.dtype)[:, None]
            inside = ((samp >= lo_t) & (samp < hi_t)).all(dim=0)
            w[inside] = float(c)
    return w


def decompose(pos, boxsize, group=None, sample_per_rank=8192, seed=0):
    """One cuboid domain (lo[3], hi[3]) per rank; identical on every rank.

    Periodic: domains tile [gmin, gmin + boxsize)^3.  Open: outer faces are +-inf.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev, dt = pos.device, pos.dtype
    n = pos.shape[0]
    g = torch.Generator(device="cpu").manual_seed(seed + 7919 * (dist.get_rank(group) if dist.is_initialized() else 0))
    m = sample_per_rank
    samp = torch.full((m, 3), float("nan"), dtype=dt, device=dev)
    if n > 0:
        idx = torch.randint(0, n, (min(m, n),), generator=g).to(dev)
        samp[: idx.numel()] = pos[idx]
    if n > 0:
        mn, mx = pos.min(dim=0).values, pos.max(dim=0).values
    else:
        mn = torch.full((3,), float("inf"), dtype=dt, device=dev)
        mx = -mn
    if world > 1:
        staged = dev.type == "cuda" and _staged(group)             # (gloo: device tensors cross through host memory)
        if staged:
            samp, mn, mx = samp.cpu(), mn.cpu(), mx.cpu()
        allsamp = [torch.empty_like(samp) for _ in range(world)]
        dist.all_gather(allsamp, samp, group=group)
        samp = torch.cat(allsamp)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        if staged:
            samp = samp.to(dev)
    samp = samp[~torch.isnan(samp[:, 0])].t().contiguous()          # [3][m]: one contiguous row per axis
    mn, mx = mn.cpu(), mx.cpu()
    periodic = boxsize is not None and boxsize > 0 and bool(((mx - mn) <= boxsize).all())
    if periodic:
        lo, hi = mn.clone(), mn + torch.tensor(boxsize, dtype=dt)
    else:
        lo = torch.full((3,), float("-inf"), dtype=dt)
        hi = torch.full((3,), float("inf"), dtype=dt)
    out = Domains([None] * world)
    out.splits = []
    _bisect(samp, _sample_weights(world, samp), lo, hi, list(range(world)), out, out.splits)
    return out, periodic


def owner_of(pos, domains):
    """Rank owning each particle.  Faces on the outside of the decomposed box are treated as open, so
    a particle sitting exactly on (or rounded just past) the outer edge still has exactly one owner."""
    splits = getattr(domains, "splits", None)
    if splits:
        # walk the bisection: one coordinate comparison per level instead of a box test per rank
        n = pos.shape[0]
        dev = pos.device
        axis = torch.tensor([sp[0] for sp in splits], dtype=torch.int64, device=dev)
        plane = torch.tensor([sp[1] for sp in splits], dtype=pos.dtype, device=dev)
        child = torch.tensor([[sp[2], sp[3]] for sp in splits], dtype=torch.int64, device=dev)
        node = torch.zeros(n, dtype=torch.int64, device=dev)             # >= 0: split index; < 0: -(rank + 1)
        depth = max(1, int(math.ceil(math.log2(max(len(domains), 2)))) + 1)
        for _ in range(depth):
            at = node.clamp(min=0)
            coord = pos.gather(1, axis[at][:, None]).squeeze(1)
            nxt = child[at, (coord >= plane[at]).to(torch.int64)]
            node = torch.where(node >= 0, nxt, node)
        return -node - 1
    owner = torch.zeros(pos.shape[0], dtype=torch.int64, device=pos.device)
    root_lo = torch.stack([d[0] for d in domains]).min(dim=0).values
    root_hi = torch.stack([d[1] for d in domains]).max(dim=0).values
    for r, (lo, hi) in enumerate(domains):
        lo = torch.where(lo <= root_lo, torch.full_like(lo, float("-inf")), lo).to(pos.device)
        hi = torch.where(hi >= root_hi, torch.full_like(hi, float("inf")), hi).to(pos.device)
        inside = ((pos >= lo) & (pos < hi)).all(dim=1)
        owner = torch.where(inside, torch.full_like(owner, r), owner)
    return owner


def _box_gap2(x, lo, hi, period):
    """Squared distance from points x[n,3] to the cuboid [lo,hi), nearest periodic image per axis."""
    def gap(v):
        return torch.clamp(torch.maximum(lo - v, v - hi), min=0)
    g = gap(x)
    if period is not None:
        g = torch.minimum(g, torch.minimum(gap(x + period), gap(x - period)))
    return (g * g).sum(dim=-1)


# ------------------------------------------------------------------------------------------------
# collectives
# ------------------------------------------------------------------------------------------------
def _staged(group):
    """True when device tensors have to cross the collective through host memory: the gloo backend (CPU tests, and the
    single-GPU multi-rank test of the device kernels) has no all-to-all for CUDA tensors.  NCCL moves them directly."""
    return dist.get_backend(group) == "gloo"


def _alltoallv(per_dest, group=None):
    """Variable-size all-to-all: per_dest[r] goes to rank r; returns the list received from each rank."""
    world = dist.get_world_size(group)
    dev = per_dest[0].device
    if dev.type == "cuda" and _staged(group):
        return [t.to(dev) for t in _alltoallv([t.cpu() for t in per_dest], group)]
    counts = torch.tensor([t.shape[0] for t in per_dest], dtype=torch.int64, device=dev)
    rcounts = torch.empty_like(counts)
    dist.all_to_all_single(rcounts, counts, group=group)
    rc, sc = rcounts.tolist(), counts.tolist()
    send = torch.cat(per_dest).contiguous()
    recv = torch.empty((sum(rc),) + tuple(send.shape[1:]), dtype=send.dtype, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc, group=group)
    return list(recv.split(rc))


def _alltoall_rows(send, counts, group=None, rcounts=None):
    """The same for rows that already lie destination-major in one tensor: (received rows, rows per source).
    ``rcounts``: the rows every source will send, when an earlier exchange over the same routes has told us
    (then no count exchange and no host synchronisation precede the data)."""
    dev = send.device
    if dev.type == "cuda" and _staged(group):
        recv, rc = _alltoall_rows(send.cpu(), counts, group, rcounts)
        return recv.to(dev), rc
    if rcounts is not None:
        rc = list(rcounts)
    else:
        sc = torch.tensor(counts, dtype=torch.int64, device=dev)
        rct = torch.empty_like(sc)
        dist.all_to_all_single(rct, sc, group=group)
        rc = rct.tolist()
    recv = torch.empty((sum(rc),) + tuple(send.shape[1:]), dtype=send.dtype, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=list(counts), group=group)
    return recv, rc


def reduce_image(img, dst=0, group=None):
    """Sum the per-rank partial frames (each rank rendered its own particles) onto rank `dst`."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(img, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return img


# ------------------------------------------------------------------------------------------------
class DistributedKDTree:
    """Smoothing lengths, densities and SPH reductions of a particle set spread over ranks.

    ``pos`` / ``mass`` are this rank's share of the particles in ANY order (e.g. a file chunk); results
    come back aligned with them.  Values equal the single-process results: neighbour searches are exact
    and the halo is complete by construction.
    """

    def _tick(self, name):
        """Phase timing (PNB_DIST_PROFILE=1): synchronised wall-clock per phase in self.timings (ms)."""
        if not self._profile:
            return
        if self.dev.type == "cuda":
            torch.cuda.current_stream().synchronize()     # (not the device: the interior search may be running beside us)
        now = time.perf_counter()
        self.timings[name] = self.timings.get(name, 0.0) + (now - self._t_last) * 1e3
        self._t_last = now

    def __init__(self, pos, mass, boxsize=None, group=None, backend=None):
        self._profile = bool(os.environ.get("PNB_DIST_PROFILE"))
        self.timings = {}
        self.dev = pos.device
        if self._profile and self.dev.type == "cuda":
            torch.cuda.synchronize()
        self._t_last = time.perf_counter()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.backend = backend or GpuBackend()
        self.n_in = int(pos.shape[0])
        self.dev, self.dt = pos.device, pos.dtype
        self.domains, periodic = decompose(pos, boxsize, group)
        self._tick("decompose")
        self.boxsize = float(boxsize) if periodic else None
        self.requested_boxsize = boxsize
        if self.world > 1 and pos.is_cuda and getattr(self.backend, "native_routing", False) and self.world <= 256 \
                and getattr(self.domains, "splits", None) is not None:
            # positions and masses travel together as [n][4] rows, packed destination-major by the library
            packed, src_idx, counts = self.backend.route(pos, mass, self.domains, self.world)
            self._ret_counts = counts                               # what I sent to each rank (returns come back so)
            self._ret_index = src_idx                               # original local index of each sent particle
            recv, self._own_counts = _alltoall_rows(packed, counts, group)   # owned particles grouped by source rank
            self.pos, self.mass = recv[:, :3].contiguous(), recv[:, 3].contiguous()
        elif self.world > 1:
            owner = owner_of(pos, self.domains)
            # one-byte keys: the stable radix sort needs a single pass instead of eight
            key = owner.to(torch.uint8) if self.world <= 255 else owner
            order = torch.sort(key, stable=True).indices
            counts = torch.bincount(owner, minlength=self.world).tolist()
            src_idx = order                                         # original local index of each sent particle
            self._ret_counts = counts                               # what I sent to each rank (returns come back so)
            self._ret_index = src_idx
            # positions and masses travel together: one exchange of [n][4] rows
            packed = torch.cat([pos[order], mass[order, None].to(pos.dtype)], dim=1)
            recv = _alltoallv(list(packed.split(counts)), group)
            self._own_counts = [int(t.shape[0]) for t in recv]      # owned particles grouped by source rank
            recv = torch.cat(recv)
            self.pos, self.mass = recv[:, :3].contiguous(), recv[:, 3].contiguous()
        else:
            self.pos, self.mass = pos, mass
        self._tick("route_particles")
        self.n_own = int(self.pos.shape[0])
        self.tree1 = self.backend.make_tree(self.pos, self.mass, self.boxsize)
        self._tick("build_owned_tree")
        self.tree2 = None
        self._halo_send = None          # per destination rank: indices (into owned) of the particles it needs
        self._k = None
        self.stats = {}

    # -- helpers ------------------------------------------------------------------------------------
    def _to_origin(self, owned_values):
        """Owned-order values -> the rank/order the particles came from."""
        if self.world == 1:
            return owned_values
        # the routes of the particle exchange, backwards: every count is known
        flat, _ = _alltoall_rows(owned_values.contiguous(), self._own_counts, self.group, self._ret_counts)
        out = torch.empty((self.n_in,) + tuple(flat.shape[1:]), dtype=flat.dtype, device=flat.device)
        out[self._ret_index] = flat
        return out

    def _from_origin(self, values):
        """Per-particle values given in the input order -> owned order (same routing as the positions)."""
        if self.world == 1:
            return values
        v = values[self._ret_index]
        return _alltoall_rows(v.contiguous(), self._ret_counts, self.group, self._own_counts)[0]

    def _faces(self):
        """(lo[3], hi[3]) of this rank's domain as numpy arrays, +-inf where there is no face (open outer faces, or a
        periodic domain that spans the whole period along an axis)."""
        lo, hi = (t.clone().to(torch.float64) for t in self.domains[self.rank])
        if self.boxsize is not None:
            full = (hi - lo) >= self.boxsize * (1 - 1e-12)
            lo[full] = float("-inf")
            hi[full] = float("inf")
        return lo.numpy(), hi.numpy()

    def _boundary_distance(self):
        lo, hi = (t.to(self.dev) for t in self.domains[self.rank])
        d = torch.minimum(self.pos - lo, hi - self.pos)             # +inf along open / unsplit outer faces
        if self.boxsize is not None:
            full = (hi - lo) >= self.boxsize * (1 - 1e-12)          # domain spans the whole period: no face
            d = torch.where(full, torch.full_like(d, float("inf")), d)
        return d.min(dim=1).values

    # -- smoothing lengths + halo construction ---------------------------------------------------------
    def smooth(self, k, kernel_id=0):
        """Smoothing lengths (and, as a by-product, the halo for all later passes with this k)."""
        self._tick("idle")
        worker = None
        # (opt-in, PNB_DIST_TWO_PHASE=1: measured SLOWER at 2 GPUs x 16.7 M particles -- 97 vs 89 ms per step,
        # profiles/r2_scaling.md.  The interior search is a persistent kernel that fills every SM; the pipeline's NCCL
        # and sort kernels do not find room beside it and wait for it, and confining it to a share of the SMs
        # slows the pipeline's own searches by more than the overlap returns.)
        two_phase = (self.world > 1 and self.n_own >= k and getattr(self.backend, "two_phase", False)
                     and bool(os.environ.get("PNB_DIST_TWO_PHASE")))
        if two_phase:
            # Phase A: only the particles whose ball MIGHT reach a face of the domain (a guaranteed bound from the
            # tree, pnb_mark_near_faces).  Phase B -- everybody else, i.e. most of the domain, none of whom can need
            # a halo -- then runs on a helper thread (its own CUDA stream) while this thread goes through the halo
            # pipeline: boundary-query exchange, halo exchange, second tree, re-search.
            near = self.tree1.mark_near_faces(k, *self._faces())
            sm = self.tree1.knn_smooth(k, near, kernel_id)
            self._tick("knn_near_faces")
            from .kdtree import kdmain
            kdmain.set_option("knn_blocks_per_sm", 8)            # leave room on every SM for the pipeline's kernels
            box = {}

            def interior():
                try:
                    if self.dev.type == "cuda":
                        torch.cuda.set_device(self.dev)             # the current device is per thread
                    box["sm"] = self.tree1.knn_smooth(k, near, kernel_id, invert=True)
                except BaseException as e:  # noqa: BLE001
                    box["error"] = e

            worker = threading.Thread(target=interior)
            worker.start()
        else:
            near = None
            sm = self.tree1.knn_smooth(k, None, kernel_id) if self.n_own >= k else None
            self._tick("knn_owned")
        self._sm_stage1 = sm
        self._share_costs()
        if self.world == 1:
            if sm is None:
                raise ValueError("Number of smoothing particles exceeds number of particles in tree")
            self._sm_own, self._k = sm, k
            return sm
        if sm is None:                                              # fewer than k owned particles: everything is boundary
            r_ub = torch.full((self.n_own,), float("inf"), dtype=self.dt, device=self.dev)
            sm = torch.zeros(self.n_own, dtype=self.dt, device=self.dev)
        else:
            r_ub = 2 * sm
        flagged = r_ub * (1 + 1e-6) >= self._boundary_distance()
        if near is not None:
            flagged &= near                                         # (phase A computed only these; the rest is interior)
        period = self.boxsize
        # boundary queries -> the ranks whose domain their ball touches
        fidx = torch.nonzero(flagged).squeeze(1)
        fpos, frad = self.pos[fidx], r_ub[fidx] * (1 + 1e-6)      # cover rounding of r = 2h = sqrt(d2_k)
        # (query, rank) pairs for every domain the ball touches -- all ranks in one pass (a block of ranks at a time), one
        # size read-back and one exchange of rows that already lie destination-major, instead of a test, a compaction and
        # a host synchronisation per rank
        lim = frad * frad * (1 + 1e-6)
        touch = torch.zeros((self.world, fpos.shape[0]), dtype=torch.bool, device=self.dev)
        for r0 in range(0, self.world, 8):
            rr = range(r0, min(self.world, r0 + 8))
            lo = torch.stack([self.domains[r][0] for r in rr]).to(self.dev)[:, None, :]          # [b][1][3]
            hi = torch.stack([self.domains[r][1] for r in rr]).to(self.dev)[:, None, :]
            touch[r0:r0 + len(rr)] = _box_gap2(fpos[None, :, :], lo, hi, period) <= lim[None, :]
        touch[self.rank] = False
        pairs = touch.nonzero()                                    # sorted by rank, then by query
        counts = touch.sum(dim=1).tolist()
        sel = pairs[:, 1]
        balls_flat, rcounts = _alltoall_rows(torch.cat([fpos[sel], frad[sel, None]], dim=1), counts, self.group)
        balls = list(balls_flat.split(rcounts))
        self._tick("select_and_send_boundary_queries")
        # which of my particles do they need?  One flag row per requesting rank, one compaction for all of them
        need = torch.zeros((self.world, self.n_own), dtype=torch.bool, device=self.dev)
        for r in range(self.world):
            b = balls[r]
            if r == self.rank or b.shape[0] == 0:
                continue
            rad = torch.where(torch.isfinite(b[:, 3]), b[:, 3], torch.full_like(b[:, 3], 1e30))
            need[r] = self.tree1.mark_in_balls(b[:, :3].contiguous(), rad).to(torch.bool)
        self._halo_flat = need.nonzero()[:, 1]                     # grouped by requesting rank, ascending within a rank
        self._halo_counts = need.sum(dim=1).tolist()
        self._halo_send = list(self._halo_flat.split(self._halo_counts))
        del need
        halo, self._halo_rcounts = _alltoall_rows(torch.cat([self.pos[self._halo_flat], self.mass[self._halo_flat, None].to(self.dt)],
                                                            dim=1), self._halo_counts, self.group)
        halo_pos, halo_mass = halo[:, :3], halo[:, 3]
        self.n_halo = int(halo_pos.shape[0])
        self._tick("mark_and_exchange_halo")
        self.stats = {"owned": self.n_own, "boundary_queries": int(fidx.numel()), "halo": self.n_halo}
        # re-search the boundary queries on (owned particles inside some boundary ball) + halo: everything a
        # boundary query can reach, and nothing else -- the second tree is a few per cent of the first
        self._near = torch.nonzero(self.tree1.mark_in_balls(fpos, frad)).squeeze(1) if fidx.numel() else fidx
        self.n_near = int(self._near.shape[0])
        self._f2 = torch.searchsorted(self._near, fidx)            # boundary queries' positions in tree2
        self.stats["near"] = self.n_near
        if fidx.numel() > 0:
            self.tree2 = self.backend.make_tree(torch.cat([self.pos[self._near], halo_pos]),
                                                torch.cat([self.mass[self._near], halo_mass]), self.boxsize)
            self._tick("build_halo_tree")
            if self.tree2.n < k:
                raise ValueError("Number of smoothing particles exceeds number of particles in tree")
            self._mask2 = torch.zeros(self.tree2.n, dtype=torch.bool, device=self.dev)
            self._mask2[self._f2] = True
            sm2 = self.tree2.knn_smooth(k, self._mask2, kernel_id)
        self._tick("knn_boundary")
        if worker is not None:
            worker.join()
            from .kdtree import kdmain
            kdmain.set_option("knn_blocks_per_sm", 0)
            if "error" in box:
                raise box["error"]
            sm = box["sm"]                                          # every owned particle, searched on owned particles
            self._sm_stage1 = sm
            self.tree1._mask(None)
            self._tick("wait_for_interior_knn")
        if fidx.numel() > 0:
            sm = sm.clone()
            sm[fidx] = sm2[self._f2]
        self._fidx = fidx
        self._interior = ~flagged
        self._sm_own, self._k = sm, k
        return sm

    def _tree2_values(self, owned_values, halo_fill=None):
        """Per-particle values in tree2's order: owned particles near the boundary, then the halo (halo
        exchange #2 unless a constant fill is enough)."""
        near = owned_values[self._near]
        if halo_fill is not None:
            halo = torch.full((self.n_halo,) + tuple(near.shape[1:]), halo_fill, dtype=near.dtype, device=self.dev)
        else:
            halo = self._halo_values(owned_values)
        return torch.cat([near, halo])

    def _share_costs(self):
        """Tell every rank how expensive this rank's particles were to search (see _COST_FIELD)."""
        ms = getattr(self.tree1, "last_knn_ms", lambda: None)()
        if self.world == 1 or ms is None or not os.environ.get("PNB_DIST_COST_FEEDBACK"):
            return
        mine = torch.tensor([ms / max(self.n_own, 1)], dtype=torch.float64)     # (a host tensor: any backend gathers it)
        allc = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allc, mine, group=self.group)
        costs = [float(c.item()) for c in allc]
        self.knn_cost_by_rank = [round(c * 1e6, 2) for c in costs]      # ms per million particles
        record_costs(self.world, self.domains, costs)

    def _halo_values(self, owned_values):
        """Halo exchange #2: per-particle values of the halo particles, in tree2's halo order."""
        return _alltoall_rows(owned_values[self._halo_flat].contiguous(), self._halo_counts, self.group, self._halo_rcounts)[0]

    def smooth_rho(self, k, kernel_id=0):
        """(smooth, rho) aligned with the input particles."""
        sm = self.smooth(k, kernel_id)
        if self.world == 1:
            rho = self.tree1.density(sm, k, kernel_id)
        else:
            # interior particles: their neighbours are all owned, so the owned-only tree is exact (and its
            # density rode along the kNN pass when masses are uniform); boundary queries: owned + halo
            if self._sm_stage1 is not None:
                rho = self.tree1.density(self._sm_stage1, k, kernel_id)
            else:
                rho = torch.zeros(self.n_own, dtype=self.dt, device=self.dev)
            if self._fidx.numel() > 0:
                rho2 = self.tree2.density(self._tree2_values(sm, 1.0), k, kernel_id, self._mask2)
                rho = rho.clone()
                rho[self._fidx] = rho2[self._f2]
        self._rho_own = rho
        self._tick("density")
        out = self._to_origin(sm), self._to_origin(rho)
        self._tick("return_results")
        return out

    def sph_op(self, op, qty, k=None, kernel_id=0):
        """mean | disp | div | curl of `qty` (input order), after smooth_rho(k)."""
        k = k or self._k
        q_own = self._from_origin(qty)
        if self.world == 1:
            out = self.tree1.reduce(op, self._sm_own, self._rho_own, q_own, k, kernel_id)
        else:
            # interior queries see only owned particles; boundary queries run on tree2 with the halo's rho / qty
            out = self.tree1.reduce(op, self._sm_own, self._rho_own, q_own, k, kernel_id, self._interior)
            # halo exchange #2 is a collective: EVERY rank takes part, also one without boundary queries of its
            # own (an isolated clump, a void at its split plane) -- the others still need its particles' values
            rho2 = self._tree2_values(self._rho_own)
            qty2 = self._tree2_values(q_own)
            if self._fidx.numel() > 0:
                out2 = self.tree2.reduce(op, self._tree2_values(self._sm_own, 1.0), rho2, qty2, k, kernel_id, self._mask2)
                out = out.clone()
                out[self._fidx] = out2[self._f2]
        if out.dim() == 2 and op in ("disp", "div"):
            out = out[:, 0]
        return self._to_origin(out)

    def owned(self):
        """(pos, mass, smooth, rho) of the particles this rank owns -- what it renders."""
        return self.pos, self.mass, self._sm_own, self._rho_own
