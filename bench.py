#!/usr/bin/env python
"""Benchmark of the SPH smoothing hot path: particles/s for ``sim['rho']`` (k = 64).

One *step* = one pass of the path pynbody runs for ``sim['rho']`` on one particle set:
kd-tree build -> k-nearest-neighbour smoothing lengths (``sim['smooth']``) -> SPH density, through
the reference-facing ``KDTree`` API (``KDTree(pos, mass)``, ``populate('hsm')``, ``populate('rho')``).

Workload: BASELINE.json configs[1] -- 16 777 216-particle (256^3) clustered Zel'dovich-like periodic box,
WendlandC2 kernel, 64 neighbours, float64, one B200 per rank.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

``value``   whole-job particles/s with positions and masses resident in HBM (torch CUDA tensors handed
            to the C ABI by pointer), timed with CUDA events, max over ranks.
``e2e``     the same step from HOST arrays in pinned memory through the same public API: host->device
            copies of pos/mass (and smooth/mass again for populate('rho'), as the reference ABI
            re-reads its borrowed arrays) and device->host copies of smooth/rho are inside the timed region.
``--impl reference``  the reference's own compiled C++ / Cython path (oracle/_ref, built from /root/reference by
            oracle/Makefile) on all host cores, on the SAME 256^3 particle set: about 10 s per pass, so the number of
            timed passes is bounded by wall time (CPU_BUDGET_S) instead of shrinking the workload; plus one exact
            2048^2 render_image of the same particles (the Mpix/s half of the metric).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_NEIGHBOURS = 64
KERNEL = "WendlandC2Kernel"
METRIC = "particles/sec for sim['rho'] (k=64)"
UNIT = "particles/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


KERNEL_SOURCES = {"k_knn": ("knn.cu", "traverse.cuh"), "k_render_tiles": ("render.cu",)}


def source_hash(kernel="k_knn"):
    """Hash of the CODE a kernel is compiled from (comments and white space do not count); ncu-derived constants are
    only quoted for the code they were captured on."""
    import hashlib
    import re
    h = hashlib.sha256()
    for f in KERNEL_SOURCES[kernel]:
        with open(os.path.join(ROOT, "pynbody_b200", "csrc", f), "r") as fh:
            text = fh.read()
        text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)          # block comments
        text = re.sub(r"//[^\n]*", " ", text)                        # line comments (no '//' occurs inside a string here)
        h.update(" ".join(text.split()).encode())
    return h.hexdigest()[:16]


def load_ncu(key):
    """A per-launch figure from the committed ncu capture (profiles/r2_ncu_constants.json: DRAM bytes, issue-slot
    utilisation, warp instructions per particle; keys are "<kernel>:<dtype>:<n_side>:<what>"), or None when the capture
    was taken on other sources of that kernel -- stale constants are not printed."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_constants.json")) as f:
            d = json.load(f)
        kernel = key.split(":")[0]
        if d.get("source_hashes", {}).get(kernel) != source_hash(kernel):
            return None
        return d.get(key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Median SM clock and throttle reasons of the samples taken between t0 and t1 (the timed region).  The
        sampler is started before the warm-up so that nvidia-smi's own start-up (hundreds of ms of driver
        queries) stays out of the timed steps."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.25)]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_particles(n_side, seed, dtype):
    from pynbody_b200 import synth
    return synth.zeldovich_box(n_side, seed, dtype)


# ------------------------------------------------------------------------------------------------
CPU_BUDGET_S = 100.0      # the CPU arm repeats the FULL workload until this much wall time is spent (>= 1 pass)


def reference_step(pos, mass, threads):
    """The reference's own path for sim['rho'] (sph/__init__.py:40-94), via its compiled modules."""
    from oracle import ref
    kd = ref.RefKDTree(pos, mass, 16, 1.0, threads, 1)
    sm = kd.smooth(K_NEIGHBOURS)
    rho = kd.rho(K_NEIGHBOURS)
    return sm, rho


def oracle_step(pos, mass, threads):
    from oracle import oracle
    kd = oracle.OracleKDTree(pos, mass, 16, 1.0, threads, 1)
    sm = kd.smooth(K_NEIGHBOURS)
    rho = kd.rho(K_NEIGHBOURS)
    return sm, rho


def render_args(res, pos, sm, rho, mass, kern):
    """render_image arguments of BASELINE.json configs[2]: projected density of the whole periodic box, with the
    3 x 3 wrap offsets ImageRenderer generates for width = boxsize (renderers.py:616-629)."""
    wrap = [-1.0, 0.0, 1.0]
    return (res, res, pos[:, 0], pos[:, 1], pos[:, 2], sm, 0.0, 1.0, 0.0, 1.0, 0.0, 0.5, rho, mass, rho,
            0.0, float("inf"), float("-inf"), float("inf"), 0.0, kern, wrap, wrap)


def cpu_render(pos, sm, rho, mass, res, threads):
    """The reference's _render.render_image (exact, threaded as ThreadedImageRenderer splits it,
    renderers.py:558-573) on all host cores."""
    from oracle import ref, oracle
    from pynbody_b200.sph import kernels
    kern = kernels.create_kernel(KERNEL).projection()
    a = render_args(res, pos, sm, rho, mass, kern)
    t0 = time.perf_counter()
    if ref.available():
        ref.render_image(*a, num_threads=threads)
        kind = "reference"
    else:
        oracle.render_image(*a)
        kind, threads = "port", 1
    dt = time.perf_counter() - t0
    return {"value": res * res / dt / 1e6, "unit": "Mpix/s", "cores": threads, "kind": kind, "seconds": dt,
            "sample": f"the full workload once: {res}x{res} projected density of all {len(pos)} particles, 3x3 periodic images, "
                      f"exact (approximate_fast=False), particle-strided over {threads} threads"}


def cpu_arm(n_side, dtype, steps, budget_s, render_res=0):
    """The reference's CPU implementation on all host cores, on the FULL workload (the same particle set the
    GPU arm runs).  One pass is ~10 s on 16 cores, so the number of timed passes is bounded by wall time, not
    by shrinking the particle set."""
    from oracle import ref
    threads = os.cpu_count() or 1
    pos, vel, mass = make_particles(n_side, 2, dtype)
    if ref.available():
        kind, fn = "reference", reference_step
    else:
        kind, fn = "port", oracle_step
    times = []
    t_start = time.perf_counter()
    sm = rho = None
    while len(times) < max(steps, 1):
        t0 = time.perf_counter()
        sm, rho = fn(pos, mass, threads)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start + times[-1] > budget_s:
            break
    dt = float(np.mean(times))
    out = {"value": len(pos) / dt, "unit": UNIT, "cores": threads, "kind": kind,
           "sample": f"the full workload: {n_side}^3 = {len(pos)} particles, {np.dtype(dtype).name}, k={K_NEIGHBOURS}, "
                     f"WendlandC2: tree build + smooth + rho, {len(times)} timed pass(es) of {dt:.1f} s",
           "seconds_per_step": dt, "passes": len(times)}
    if render_res:
        out["render"] = cpu_render(pos, sm, rho, mass, render_res, threads)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dtype = np.float64 if args.dtype == "f64" else np.float32
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if world > 1:
        # the N-GPU arm searches ONE global box of ~N x n_side^3 particles: so does this arm, on the host cores (same
        # recipe and size; the GPU arm draws its Gaussian field on the device, so the realisation differs).  One pass
        # of 134 M particles takes minutes: no render leg here.
        res = cpu_arm(global_side(args, world), dtype, args.steps, CPU_BUDGET_S, 0)
        cfg = distributed_config(args, world)
    else:
        res = cpu_arm(args.n_side, dtype, args.steps, CPU_BUDGET_S, 0 if args.no_render else args.render_res)
        cfg = workload_config(args)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "steps_timed": res["passes"], "gpu_launches": 0}
    if "render" in res:
        r = res["render"]
        line["render"] = {"image": {"resolution": args.render_res, "ms": r["seconds"] * 1e3, "mpix_per_s": r["value"],
                                    "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}}}
    print(json.dumps(line))


def global_side(args, world):
    """Side of the ONE global box the N-GPU arms run: about N x n_side^3 particles (weak scaling)."""
    return int(round((world * args.n_side ** 3) ** (1.0 / 3.0)))


def distributed_config(args, world):
    side = global_side(args, world)
    n_total = side ** 3
    cfg = workload_config(args)
    cfg["workload"] = (f"C4/C5-style: ONE global {n_total}-particle ({side}^3) clustered Zel'dovich-like periodic box "
                       f"decomposed over {world} GPUs (~{n_total // world} particles per GPU), sim['rho'] = domain decomposition + "
                       f"tree build + smooth + rho, k={K_NEIGHBOURS}, WendlandC2, {args.dtype}")
    cfg["parallelism"] = (f"spatial domains x{world} (coordinate bisection), NCCL all-to-all particle routing + need-based halo "
                          f"exchange + masked re-search; image = per-rank frames + NCCL reduce")
    cfg["global_particles"] = n_total
    return cfg


def workload_config(args):
    n = args.n_side ** 3
    return {"workload": f"C2: {n}-particle ({args.n_side}^3) clustered Zel'dovich-like periodic box per GPU, "
                        f"sim['rho'] = kd-tree build + smooth + rho, k={K_NEIGHBOURS}, WendlandC2, {args.dtype}",
            "particles_per_gpu": n, "k": K_NEIGHBOURS, "kernel": "WendlandC2", "boxsize": 1.0,
            "parallelism": "one independent box per GPU (no data-path collective)" if args.gpus > 1 else "single GPU",
            "l2_policy": "inputs (pos+mass, 32 B/particle f64) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pynbody_b200.kdtree import KDTree, kdmain

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype = np.float64 if args.dtype == "f64" else np.float32
    tsz = np.dtype(dtype).itemsize

    pos, vel, mass = make_particles(args.n_side, 2 + rank, dtype)
    n = len(pos)
    # pinned host copies (what the e2e leg reads / writes) and device-resident copies (the value leg)
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:1]).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    h_pos, h_mass = pinned(pos), pinned(mass)
    h_sm = torch.empty(n, dtype=h_mass.dtype, pin_memory=True)
    h_rho = torch.empty(n, dtype=h_mass.dtype, pin_memory=True)
    d_pos, d_mass = h_pos.cuda(), h_mass.cuda()
    d_sm, d_rho = torch.empty_like(d_mass), torch.empty_like(d_mass)
    del pos, mass

    stats = {"build_ms": [], "knn_ms": [], "knn_kernel_ms": [], "rho_ms": [], "launches": 0}

    def step(p, m, sm, rho, record):
        kd = KDTree(p, m, leafsize=16, boxsize=1.0)
        kd.set_kernel(KERNEL)
        b_ms, b_l = kdmain.last_device_ms(0), kdmain.last_launch_count(0)
        kd.set_array_ref("smooth", sm)
        kd.populate("hsm", K_NEIGHBOURS)
        k_ms, k_l, kk_ms = kdmain.last_device_ms(1), kdmain.last_launch_count(1), kdmain.last_kernel_ms(0)
        kd.set_array_ref("mass", m)
        kd.set_array_ref("rho", rho)
        kd.populate("rho", K_NEIGHBOURS)
        r_ms, r_l = kdmain.last_device_ms(1), kdmain.last_launch_count(1)
        if args.verbose and rank == 0:
            print(f"[step] build {b_ms:.1f} ms ({b_l} launches) | hsm {k_ms:.1f} ms (kernel {kk_ms:.1f}) | rho {r_ms:.1f} ms",
                  file=sys.stderr, flush=True)
        if record:
            stats["build_ms"].append(b_ms); stats["knn_ms"].append(k_ms); stats["knn_kernel_ms"].append(kk_ms)
            stats["rho_ms"].append(r_ms); stats["launches"] += b_l + k_l + r_l
        return kd

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step(d_pos, d_mass, d_sm, d_rho, False)
    t_region = time.time()
    total_ms = timed(lambda: step(d_pos, d_mass, d_sm, d_rho, True), args.steps)
    clocks = sampler.stop(t_region, time.time()) if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # end to end from pinned host memory through the same public API
    step(h_pos.numpy(), h_mass.numpy(), h_sm.numpy(), h_rho.numpy(), False)
    e2e_ms = timed(lambda: step(h_pos.numpy(), h_mass.numpy(), h_sm.numpy(), h_rho.numpy(), False), args.steps) / args.steps
    e2e_value = world * n / (e2e_ms * 1e-3)
    # the same from ordinary (pageable) numpy arrays, as a pynbody user would hand them over
    p_pos, p_mass = h_pos.numpy().copy(), h_mass.numpy().copy()
    p_sm, p_rho = np.empty(n, p_mass.dtype), np.empty(n, p_mass.dtype)
    step(p_pos, p_mass, p_sm, p_rho, False)
    e2e_pageable_ms = timed(lambda: step(p_pos, p_mass, p_sm, p_rho, False), 2) / 2
    del p_pos, p_mass, p_sm, p_rho
    h2d = n * tsz * (3 + 1) + n * tsz * 2        # pos + mass at build; smooth + mass re-read by populate('rho')
    d2h = n * tsz * 2                            # smooth, rho

    # other stages of the path, for the record (not part of `value`)
    kd = step(d_pos, d_mass, d_sm, d_rho, False)
    d_vel = torch.from_numpy(vel).cuda()
    extra = {}
    for name, fn in (("v_mean", kd.sph_mean), ("v_disp", kd.sph_dispersion), ("v_div", kd.sph_divergence),
                     ("v_curl", kd.sph_curl)):
        for _ in range(2):                       # the first call pays for the result tensor's first allocation
            out = fn(d_vel, K_NEIGHBOURS)
            torch.cuda.synchronize()
            del out
        extra[name + "_ms"] = round(kdmain.last_device_ms(1), 3)
        extra[name + "_kernel_ms"] = round(kdmain.last_kernel_ms(1), 3)
    # sim['rho'] when the masses are NOT all equal (any real gas snapshot): the density cannot ride in the kNN
    # epilogue; it is summed over the resident neighbour lists instead (gather_list.cu)
    d_mass2 = d_mass * (1.0 + 0.25 * torch.sin(torch.arange(n, device="cuda", dtype=d_mass.dtype)))
    for _ in range(2):
        step(d_pos, d_mass2, d_sm, d_rho, False)
    uneq_ms = timed(lambda: step(d_pos, d_mass2, d_sm, d_rho, False), 3) / 3
    extra["unequal_masses_step_ms"] = round(uneq_ms, 3)
    extra["unequal_masses_rho_ms"] = round(kdmain.last_device_ms(1), 3)
    extra["unequal_masses_rho_kernel_ms"] = round(kdmain.last_kernel_ms(1), 3)
    # and with a foreign smooth array (not this tree's own): the tree-walking fixed-radius gather of the reference
    sm2 = d_sm.clone(); sm2[0] = sm2[0] * (1 + 1e-7)
    kd.set_array_ref("smooth", sm2); kd.set_array_ref("rho", d_rho)
    kd.populate("rho", K_NEIGHBOURS)
    extra["rho_tree_walk_ms"] = round(kdmain.last_device_ms(1), 3)
    extra["rho_tree_walk_kernel_ms"] = round(kdmain.last_kernel_ms(1), 3)
    del d_mass2, sm2

    # C3: projected density image and 3-D grid of the same particles (BASELINE.json configs[2]),
    # device-resident inputs, periodic wrap offsets as ImageRenderer generates them (renderers.py:616-629)
    render = {}
    if not args.no_render:
        from pynbody_b200.sph import _render, kernels
        k3 = kernels.create_kernel(KERNEL)
        k2 = k3.projection()
        wrap = [-1.0, 0.0, 1.0]
        res = args.render_res
        def once():
            return _render.render_image(*render_args(res, d_pos, d_sm, d_rho, d_mass, k2))
        once()
        ms = timed(once, 3) / 3
        call_ms, kern_ms = _render.last_render_ms()
        # end to end: the seven particle arrays in pinned HOST memory (x, y, z = column views of pos), image back to the host
        hp, hs, hr, hm = h_pos.numpy(), h_sm.numpy(), h_rho.numpy(), h_mass.numpy()
        hs[...] = d_sm.cpu().numpy(); hr[...] = d_rho.cpu().numpy()
        def once_host():
            return _render.render_image(*render_args(res, hp, hs, hr, hm, k2))
        once_host()
        e2e_img_ms = timed(once_host, 3) / 3
        # compulsory bytes N x 7 x s_in + 4 nx ny out + ~16 B/particle binning traffic
        img_bytes = n * 7 * tsz + 4 * res * res + n * 16
        peak_r, peak_src_r = load_peaks()
        render["image"] = {"resolution": res, "mode": "projected density, 3x3 periodic images, exact (approximate_fast=False)",
                           "ms": ms, "mpix_per_s": world * res * res / (ms * 1e-3) / 1e6, "tile_kernel_ms": kern_ms,
                           "e2e": {"value": world * res * res / (e2e_img_ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms": e2e_img_ms,
                                   "h2d_bytes_per_step": n * 7 * tsz, "d2h_bytes_per_step": 4 * res * res},
                           "roofline": {"bound": "hbm", "kernel": "k_render_tiles", "achieved": img_bytes / (kern_ms * 1e-3) / 1e9,
                                        "peak": peak_r, "unit": "GB/s", "frac": img_bytes / (kern_ms * 1e-3) / 1e9 / peak_r,
                                        "traffic": load_ncu(f"k_render_tiles:{args.dtype}:{args.n_side}:dram_bytes"),
                                        "algorithmic_bytes_per_launch": img_bytes, "peak_source": peak_src_r, "kernel_ms": kern_ms,
                                        "pixel_updates_per_launch": load_ncu(f"k_render_tiles:{args.dtype}:{args.n_side}:pixel_updates"),
                                        "note": "work-bound (pixel updates), not byte-bound: SURVEY.md 8(d)"}}
        gres = args.grid_res
        def gonce():
            return _render.to_3d_grid(gres, gres, gres, d_pos[:, 0], d_pos[:, 1], d_pos[:, 2], d_sm, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0,
                                      d_rho, d_mass, d_rho, 0.0, float("inf"), k3, wrap, wrap, wrap)
        gonce()
        ms = timed(gonce, 1)
        call_ms, kern_ms = _render.last_render_ms()
        render["grid"] = {"resolution": gres, "ms": ms, "mvox_per_s": world * gres ** 3 / (ms * 1e-3) / 1e6, "tile_kernel_ms": kern_ms}
        # healpix map (projected column per solid angle) seen from the box centre, and the tree-less sphere selection
        nside = 256
        c_pos = d_pos - 0.5
        shell = torch.nonzero(((c_pos * c_pos).sum(dim=1) > 0.09) & ((c_pos * c_pos).sum(dim=1) < 0.25)).squeeze(1)
        h_pos, h_rho, h_mass, h_sm = c_pos[shell].contiguous(), d_rho[shell], d_mass[shell], d_sm[shell]
        def honce():
            return _render.render_spherical_image_core(h_rho, h_mass, h_rho, h_pos[:, 0], h_pos[:, 1], h_pos[:, 2], h_sm, nside, k2)
        honce()
        ms = timed(honce, 1)
        render["healpix"] = {"nside": nside, "particles": int(shell.numel()), "selection": "0.3 < r < 0.5 around the box centre",
                             "ms": ms, "mpix_per_s": 12 * nside * nside / (ms * 1e-3) / 1e6,
                             "kernel_ms": _render.last_render_ms()[1]}
        del h_pos, h_rho, h_mass, h_sm, shell
        from pynbody_b200.filt import geometry_selection
        cp = d_pos.contiguous()
        sel = lambda: geometry_selection.selection(cp, "sphere", (0.5, 0.5, 0.5, 0.25), 1.0)
        sel()
        ms = timed(sel, 3) / 3
        kms = geometry_selection.last_selection_ms()[1]
        render["sphere_selection"] = {"ms": ms, "kernel_ms": kms, "kernel_gb_per_s": n * (3 * tsz + 1) / (kms * 1e-3) / 1e9}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    knn_kernel_ms = float(np.mean(stats["knn_kernel_ms"]))
    alg_bytes = n * 6 * tsz                      # read pos 3s + mass s, write smooth s + rho s
    achieved = alg_bytes / (knn_kernel_ms * 1e-3) / 1e9
    gather_bytes = n * K_NEIGHBOURS * 4 * tsz + alg_bytes
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "pageable_host_arrays_ms_per_step": e2e_pageable_ms},
        "gpu_launches": stats["launches"],
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_knn (kNN smoothing lengths + fused density)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": load_ncu(f"k_knn:{args.dtype}:{args.n_side}:dram_bytes"), "algorithmic_bytes_per_launch": alg_bytes,
                     "peak_source": peak_src, "algorithmic_bytes_per_particle": 6 * tsz,
                     "kernel_ms": knn_kernel_ms,
                     "logical_gather_gbs": gather_bytes / (knn_kernel_ms * 1e-3) / 1e9,
                     "neighbour_interactions_per_s": n * K_NEIGHBOURS / (knn_kernel_ms * 1e-3),
                     "ncu_issue_active_pct": load_ncu(f"k_knn:{args.dtype}:{args.n_side}:issue_active_pct"),
                     "ncu_warp_inst_per_particle": load_ncu(f"k_knn:{args.dtype}:{args.n_side}:warp_inst_per_particle"),
                     "ncu_source_hash": source_hash("k_knn"),
                     "note": "irregular tree search: bound by instruction issue (on-chip neighbour-queue updates), not by "
                             "HBM bytes (SURVEY.md 8d); traffic / issue-slot figures come from the committed ncu capture "
                             "(profiles/r2_ncu_constants.json) and are null unless it was taken on these very sources"},
        "stages_ms": {"build": float(np.mean(stats["build_ms"])), "smooth_incl_fused_rho": float(np.mean(stats["knn_ms"])),
                      "knn_kernel": knn_kernel_ms, "rho_readback": float(np.mean(stats["rho_ms"])), **extra},
        "unequal_masses": {"ms_per_step": extra["unequal_masses_step_ms"], "value": world * n / (extra["unequal_masses_step_ms"] * 1e-3),
                           "unit": UNIT, "note": "same step with particle masses that differ: density from the resident "
                                                 "neighbour lists instead of the kNN epilogue"},
        "render": render,
    }
    if world == 1 and not args.no_cpu_baseline:
        res = cpu_arm(args.n_side, dtype, 1, 0.0, 0 if args.no_render else args.render_res)
        line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if "render" in res and "image" in render:
            render["image"]["cpu_baseline"] = {k: res["render"][k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_ours_distributed(args):
    """N > 1: ONE global periodic box of about N x n_side^3 particles, spatially decomposed over the
    ranks (pynbody_b200.distributed): coordinate-bisection domains, NCCL all-to-all of the particles,
    need-based halo exchange, masked re-search of boundary queries.  Weak scaling: the per-GPU share
    is fixed, the global box grows with N."""
    import torch
    import torch.distributed as dist
    from pynbody_b200 import synth
    from pynbody_b200.distributed import DistributedKDTree
    from pynbody_b200.kdtree import kdmain

    rank, world, local = (int(os.environ.get(v, "0")) for v in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype = np.float64 if args.dtype == "f64" else np.float32
    tsz = np.dtype(dtype).itemsize
    side = global_side(args, world)
    n_total = side ** 3
    pos, vel, mass = synth.zeldovich_box_device(side, 2, dtype, "cuda", rank, world)
    n_local = int(pos.shape[0])
    h_pos = torch.empty(pos.shape, dtype=pos.dtype, pin_memory=True); h_pos.copy_(pos)
    h_mass = torch.empty(mass.shape, dtype=mass.dtype, pin_memory=True); h_mass.copy_(mass)
    h_sm = torch.empty(n_local, dtype=mass.dtype, pin_memory=True)
    h_rho = torch.empty(n_local, dtype=mass.dtype, pin_memory=True)
    last = {}

    def step(p, m):
        d = DistributedKDTree(p, m, boxsize=1.0)
        sm, rho = d.smooth_rho(K_NEIGHBOURS, 1)
        last["d"] = d
        if args.verbose and rank == 0:
            print("[dist step] owned=%d knn ms/Mparticle by rank: %s" % (d.n_own, getattr(d, "knn_cost_by_rank", None)),
                  file=sys.stderr, flush=True)
        return sm, rho

    def step_e2e():
        p = h_pos.cuda(non_blocking=True)
        m = h_mass.cuda(non_blocking=True)
        sm, rho = step(p, m)
        h_sm.copy_(sm, non_blocking=True)
        h_rho.copy_(rho, non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step(pos, mass)
    l0 = kdmain.total_launches()
    t_region = time.time()
    total_ms = timed(lambda: step(pos, mass), args.steps)
    launches = kdmain.total_launches() - l0
    clocks = sampler.stop(t_region, time.time()) if rank == 0 else None
    ms_per_step = total_ms / args.steps
    step_e2e()
    e2e_ms = timed(step_e2e, args.steps) / args.steps
    d = last["d"]
    # the dominant kernel of the decomposed step is the search of a rank's OWNED particles; later launches (the masked
    # re-search of the boundary queries) overwrite its timer, so it is run once more, untimed, on the last step's tree
    d.tree1.knn_smooth(K_NEIGHBOURS, None, 1)
    torch.cuda.synchronize()
    knn_owned_ms = kdmain.last_kernel_ms(0)
    if d.timings:                                   # PNB_DIST_PROFILE=1: phase table of every rank
        names = list(d.timings)
        mine = torch.tensor([d.timings[k] for k in names], device="cuda", dtype=torch.float64)
        allt = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine)
        if rank == 0:
            for i, nm in enumerate(names):
                print("[dist phase] %-36s %s" % (nm, " ".join("%7.1f" % float(t[i]) for t in allt)), file=sys.stderr, flush=True)
    if args.verbose and rank == 0 and d.timings:
        print("[dist phases ms] " + ", ".join(f"{k}={v:.1f}" for k, v in d.timings.items()), file=sys.stderr, flush=True)
        print("[dist stats] " + json.dumps(d.stats), file=sys.stderr, flush=True)
    st = torch.tensor([d.n_own, d.stats.get("boundary_queries", 0), d.stats.get("halo", 0)], device="cuda", dtype=torch.float64)
    allst = [torch.empty_like(st) for _ in range(world)]
    dist.all_gather(allst, st)

    # C4: the gradient operators on the decomposed box (halo exchange #2 carries rho and the velocities)
    stages = {}
    for op in ("div", "curl", "mean"):
        d.sph_op(op, vel, K_NEIGHBOURS, 1)
        stages["v_%s_ms" % op] = round(timed(lambda: d.sph_op(op, vel, K_NEIGHBOURS, 1), 2) / 2, 3)

    # C5-style image: every rank renders the particles it owns, frames are summed over NVLink
    render = {}
    if not args.no_render:
        from pynbody_b200.distributed import reduce_image
        from pynbody_b200.sph import _render, kernels
        k2 = kernels.create_kernel(KERNEL).projection()
        wrap = [-1.0, 0.0, 1.0]
        opos, omass, osm, orho = d.owned()
        res = args.render_res
        def once():
            img = _render.render_image(res, res, opos[:, 0], opos[:, 1], opos[:, 2], osm, 0.0, 1.0, 0.0, 1.0, 0.0, 0.5,
                                       orho, omass, orho, 0.0, float("inf"), float("-inf"), float("inf"), 0.0, k2, wrap, wrap)
            return reduce_image(img)
        for res in sorted({args.render_res, 4096}):
            once()
            ms = timed(once, 2) / 2
            render["image" if res == args.render_res else "image_%d" % res] = {
                "resolution": res, "mode": "projected density, 3x3 periodic images, per-rank frames summed with an NCCL reduce",
                "ms": ms, "mpix_per_s": res * res / (ms * 1e-3) / 1e6}
    if rank == 0:
        peak, peak_src = load_peaks()
        algo = 6 * tsz * d.n_own                        # 6 s bytes per particle incl. the fused density
        achieved = algo / (knn_owned_ms * 1e-3) / 1e9 if knn_owned_ms and knn_owned_ms > 0 else None
        cfg = distributed_config(args, world)          # (identical to the reference arm's config for the same N)
        line = {"metric": METRIC, "value": n_total / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": cfg,
                "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": n_local * tsz * 4, "d2h_bytes_per_step": n_local * tsz * 2},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "k_knn (rank 0: search of the rank's %d owned particles, re-run once after "
                                                         "the timed steps)" % d.n_own,
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                             "traffic": None, "algorithmic_bytes_per_launch": algo, "peak_source": peak_src,
                             "kernel_ms": knn_owned_ms,
                             "note": "irregular tree search, bound by instruction issue (SURVEY.md 8d); DRAM traffic of the "
                                     "kernel is captured by the N=1 run"},
                "per_rank_owned_boundary_halo": [[int(v) for v in st_.tolist()] for st_ in allst],
                "stages_ms": stages, "render": render}
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--n-side", type=int, default=256, help="particles per GPU = n_side^3 (256 -> 16.7M)")
    ap.add_argument("--dtype", choices=["f64", "f32"], default="f64")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true", help="per-step stage times on stderr")
    ap.add_argument("--no-render", action="store_true", help="skip the C3 image / grid timings")
    ap.add_argument("--render-res", type=int, default=2048)
    ap.add_argument("--grid-res", type=int, default=256)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            run_ours_distributed(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
