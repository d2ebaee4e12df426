"""GPU parity tests for the kd-tree / smoothing path (run with ``-m gpu`` on a B200).

Everything goes through the C ABI (``pynbody_b200.kdtree.KDTree`` -> ``kdmain`` shim -> ctypes ->
``libpnb_b200.so``) and is compared with

* the committed golden vectors produced by the reference's own compiled modules
  (tests/golden/make_golden.py), and
* the CPU oracle (oracle/oracle.c, pinned to those goldens by tests/test_oracle.py) on larger
  seeded inputs.

Bars (BASELINE.json north_star): smoothing lengths and neighbour sets bit-exact; rho / v_mean /
v_disp / v_div / v_curl within 1e-5 relative in float64, 1e-4 in float32.
"""
import warnings

import numpy as np
import pytest

import cases
from pynbody_b200 import synth
from pynbody_b200.kdtree import kdmain

pytestmark = pytest.mark.gpu


def _tol(dtype):
    return 1e-5 if np.dtype(dtype) == np.float64 else 1e-4


def assert_close(got, want, dtype, what=""):
    """rtol of the north star, with an absolute floor tied to the array's scale so that components
    that cancel to ~0 (e.g. a mean velocity) are judged against the size of the quantity."""
    tol = _tol(dtype)
    scale = float(np.max(np.abs(want))) if want.size else 1.0
    np.testing.assert_allclose(got, want, rtol=tol, atol=tol * scale * 1e-2, err_msg=what)


def _tree(pos, mass, boxsize, kernel_id=0, leafsize=16):
    from pynbody_b200.kdtree import KDTree
    kd = KDTree(pos, mass, leafsize=leafsize, boxsize=boxsize)
    kd.set_kernel("CubicSplineKernel" if kernel_id == 0 else "WendlandC2Kernel")
    return kd


def _smooth(kd, pos, k):
    sm = np.empty(len(pos), dtype=pos.dtype)
    kd.set_array_ref("smooth", sm)
    kd.populate("hsm", k)
    return sm


def _rho(kd, pos, mass, sm, k):
    rho = np.empty(len(pos), dtype=pos.dtype)
    kd.set_array_ref("smooth", sm)
    kd.set_array_ref("mass", mass)
    kd.set_array_ref("rho", rho)
    kd.populate("rho", k)
    return rho


# ---------------------------------------------------------------------------------------------
def test_library_reports_device():
    from pynbody_b200 import _lib
    assert _lib.lib().pnb_device_count() >= 1


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 4097, 100_003])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_tree_invariants(n, dtype):
    """Device tree: buckets partition the particles, hold exactly 32 each (the last one the remainder,
    trailing ones nothing), and their bounds are tight."""
    from pynbody_b200 import _lib
    rng = np.random.default_rng(n)
    pos = rng.normal(size=(n, 3)).astype(dtype)
    kd = _tree(pos, np.ones(n, dtype), None)
    L = _lib.lib()
    h = kd.kdtree.handle
    order = np.empty(n, np.int64)
    _lib.check(L.pnb_tree_order(h, order.ctypes.data))
    assert np.array_equal(np.sort(order), np.arange(n))
    nb = L.pnb_tree_num_buckets(h)
    start = np.empty(nb + 1, np.int64)
    bounds = np.empty((nb, 6))
    _lib.check(L.pnb_tree_buckets(h, start.ctypes.data, bounds.ctypes.data))
    assert start[0] == 0 and start[-1] == n
    counts = np.diff(start)
    full = n // 32
    assert np.all(counts[:full] == 32) and counts[full:].sum() == n - 32 * full and nb < 2 * max(1, -(-n // 32))
    for b in range(0, nb, max(1, nb // 50)):
        p = pos[order[start[b]:start[b + 1]]].astype(np.float64)
        if len(p):
            assert np.array_equal(bounds[b, :3], p.min(axis=0))
            assert np.array_equal(bounds[b, 3:], p.max(axis=0))
        else:
            assert np.all(bounds[b, :3] == np.inf) and np.all(bounds[b, 3:] == -np.inf)


@pytest.mark.parametrize("build", [0, 1, 2, 3, 4], ids=["morton", "kd_hybrid", "kd_lists_onchip_tail", "kd_levelwise_cub",
                                                        "kd_levelwise_fused"])
@pytest.mark.parametrize("n,dtype", [(40, np.float64), (2048, np.float64), (2049, np.float64), (4096, np.float32),
                                     (70_001, np.float32), (300_000, np.float64)])
def test_both_tree_builds_give_a_valid_tree_and_the_same_answers(build, n, dtype):
    """Device tree from the Morton sort + shared-memory kd levels (default) and from the level-by-level build:
    every node's bounds are the union of its children's (the searches prune on them), buckets are tight, and
    the smoothing lengths -- which must not depend on the tree -- equal the oracle's bit for bit."""
    from oracle import oracle
    from pynbody_b200 import _lib
    pos, vel, mass = synth.clumpy_box(n, 3 + n % 7, dtype)
    boxsize = 1.0 if n > 1000 else None          # (a 40-particle periodic box has balls of half the box: SURVEY 8a item 4)
    kdmain.set_option("tree_build", build)
    try:
        kd = _tree(pos, mass, boxsize)
    finally:
        kdmain.set_option("tree_build", 2)
    L = _lib.lib()
    h = kd.kdtree.handle
    nb = L.pnb_tree_num_buckets(h)
    boxes = np.empty((2 * nb, 6))
    _lib.check(L.pnb_tree_node_bounds(h, boxes.ctypes.data))
    order = np.empty(n, np.int64)
    _lib.check(L.pnb_tree_order(h, order.ctypes.data))
    assert np.array_equal(np.sort(order), np.arange(n))
    for node in range(nb - 1, 0, -1):
        lo = np.minimum(boxes[2 * node, :3], boxes[2 * node + 1, :3])
        hi = np.maximum(boxes[2 * node, 3:], boxes[2 * node + 1, 3:])
        assert np.array_equal(boxes[node, :3], lo) and np.array_equal(boxes[node, 3:], hi), node
    p64 = pos.astype(np.float64)
    assert np.array_equal(boxes[1, :3], p64.min(axis=0)) and np.array_equal(boxes[1, 3:], p64.max(axis=0))
    for b in range(0, nb, max(1, nb // 40)):
        p = p64[order[32 * b:32 * b + 32]]
        if len(p):
            assert np.array_equal(boxes[nb + b, :3], p.min(axis=0)) and np.array_equal(boxes[nb + b, 3:], p.max(axis=0))
    k = min(32, n)
    sm = _smooth(kd, pos, k)
    assert np.array_equal(sm, oracle.OracleKDTree(pos, mass, 16, boxsize, None, 0).smooth(k))


@pytest.mark.parametrize("n,leafsize,dtype", [(100_000, 16, np.float64), (65_537, 16, np.float32), (5000, 32, np.float64),
                                              (777, 8, np.float32), (40, 64, np.float64), (200_000, 1, np.float64)])
def test_kdnodes_equal_the_reference_tree(n, leafsize, dtype):
    """KDTree.kdnodes / particle_offsets against the compiled reference's own build (kd.cpp:113-238): split axis from
    the inherited box, fSplit = coordinate of the median-rank particle, tight bounds from the up-pass, inclusive
    [pLower, pUpper] ranges -- record for record; particle offsets as sets per leaf (quickselect leaves the order
    inside a leaf unspecified, kd.cpp:31-66)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref (compiled reference) not built")
    pos, vel, mass = synth.clumpy_box(n, 17 + leafsize, dtype)
    want = ref.RefKDTree(pos, mass, leafsize, None, 4)
    kd = _tree(pos, mass, None, leafsize=leafsize)
    got_nodes, got_off = kd.kdnodes, kd.particle_offsets
    assert got_nodes.shape == want.kdnodes.shape and got_nodes.dtype.itemsize == 80
    w, g = want.kdnodes[1:], got_nodes[1:]
    for field in ("iDim", "pLower", "pUpper", "fSplit"):
        assert np.array_equal(g[field], w[field]), field
    assert np.array_equal(g["bnd"]["fMin"], w["bnd"]["fMin"]) and np.array_equal(g["bnd"]["fMax"], w["bnd"]["fMax"])
    leaves = np.flatnonzero(want.kdnodes["iDim"] == -1)
    leaves = leaves[leaves >= 1]
    for node in leaves[:: max(1, len(leaves) // 500)]:
        lo, hi = int(want.kdnodes["pLower"][node]), int(want.kdnodes["pUpper"][node])
        assert set(got_off[lo:hi + 1]) == set(want.particle_offsets[lo:hi + 1]), node
    assert np.array_equal(np.sort(got_off), np.arange(n))


@pytest.mark.parametrize("n,leafsize,dtype", [(5000, 16, np.float64), (5000, 32, np.float32), (777, 8, np.float64),
                                              (40, 64, np.float64)])
def test_serialize_is_a_valid_reference_tree(n, leafsize, dtype):
    """KDTree.serialize() (kdtree/__init__.py:136-163): reference-layout KDNode records + offsets.

    Checked structurally, and by handing the serialized tree to the REFERENCE's own C++ code
    (kdmain.import_prebuilt) and letting it smooth with it (the reference's test_kdtree_from_existing_kdtree,
    tests/kdtree_test.py:284-294)."""
    from oracle import ref
    rng = np.random.default_rng(n)
    pos = rng.normal(1.0, size=(n, 3)).astype(dtype)
    mass = rng.uniform(size=n).astype(dtype)
    kd = _tree(pos, mass, None, leafsize=leafsize)
    leaf, boxsize, nodes, offsets, kernel_id = kd.serialize()
    assert leaf == leafsize and boxsize is None and kernel_id == 0
    assert np.array_equal(np.sort(offsets), np.arange(n))
    nsplit = len(nodes) // 2
    l = 1
    m = n
    while m > leafsize:
        m >>= 1
        l <<= 1
    assert nsplit == l                                                       # kd.cpp:98-111
    assert nodes["pLower"][1] == 0 and nodes["pUpper"][1] == n - 1
    for i in range(1, nsplit):
        lo, hi = nodes["pLower"][i], nodes["pUpper"][i]
        mid = (lo + hi) // 2
        assert (nodes["pLower"][2 * i], nodes["pUpper"][2 * i]) == (lo, mid)          # kd.cpp:176-199
        assert (nodes["pLower"][2 * i + 1], nodes["pUpper"][2 * i + 1]) == (mid + 1, hi)
        d = nodes["iDim"][i]
        assert 0 <= d < 3
        p = pos[offsets[lo:hi + 1]].astype(np.float64)
        assert np.array_equal(nodes["bnd"]["fMin"][i], p.min(axis=0)) and np.array_equal(nodes["bnd"]["fMax"][i], p.max(axis=0))
        assert p[:mid - lo + 1, d].max() <= nodes["fSplit"][i] <= p[mid - lo + 1:, d].min()
    assert np.all(nodes["iDim"][nsplit:] == -1)
    if ref.available():
        theirs = ref.RefKDTree(pos, mass, leafsize, None, 1)
        want = theirs.smooth(min(32, n))
        kdc = ref.kdmain.init(pos, mass, leafsize)
        ref.kdmain.import_prebuilt(kdc, nodes, offsets, 0)
        sm = np.empty(n, dtype)
        ref.kdmain.set_arrayref(kdc, 0, sm)
        smx = ref.kdmain.nn_start(kdc, min(32, n), -1.0)
        ref.kdmain.domain_decomposition(kdc, 1)
        ref.kdmain.populate(kdc, smx, 1, 0, 0)
        ref.kdmain.nn_stop(kdc, smx)
        ref.kdmain.free(kdc)
        assert np.array_equal(sm, want), "the reference must get its own answers from our serialized tree"
    # and the round trip through our own deserialize
    from pynbody_b200.kdtree import KDTree
    kd2 = KDTree.deserialize(pos, mass, kd.serialize())
    assert np.array_equal(_smooth(kd2, pos, min(32, n)), _smooth(kd, pos, min(32, n)))


@pytest.mark.parametrize("name", list(cases.SMOOTH_CASES))
def test_matches_reference_golden(name):
    case = cases.SMOOTH_CASES[name]
    gold = cases.load_golden("smooth", name)
    pos, vel, mass = case.make()
    dt = pos.dtype
    kd = _tree(pos, mass, case.boxsize, case.kernel_id, case.leafsize)
    sm = _smooth(kd, pos, case.k)
    assert np.array_equal(sm, gold["smooth"]), "smoothing lengths must be bit-exact"
    rho = _rho(kd, pos, mass, sm, case.k)
    assert_close(rho, gold["rho"], dt, "rho")

    # the general gather path (foreign smooth array => no fused density): nudge one value and restore
    sm2 = sm.copy()
    sm2[0] = np.nextafter(sm2[0], np.inf)
    rho2 = _rho(kd, pos, mass, sm2, case.k)
    assert_close(rho2[1:], gold["rho"][1:], dt, "rho via gather")
    # v_* through the reference's own call sequence (derived.py:126-154), with the golden rho
    # (bit-identical inputs to what the reference used)
    g_rho = gold["rho"]
    kd.set_array_ref("rho", g_rho)
    kd.set_array_ref("smooth", sm)
    kd.set_array_ref("mass", mass)
    for op, fn in (("mean", kd.sph_mean), ("curl", kd.sph_curl), ("div", kd.sph_divergence)):
        got = fn(vel, case.k)
        assert got.shape == gold["v_" + op].shape
        assert_close(got, gold["v_" + op], dt, "v_" + op)
    disp = kd.sph_dispersion(vel, case.k)
    assert_close(disp[:, 0], gold["v_disp"], dt, "v_disp")
    vx = np.ascontiguousarray(vel[:, 0])
    assert_close(kd.sph_mean(vx, case.k), gold["vx_mean"], dt, "vx_mean")
    assert_close(kd.sph_dispersion(vx, case.k), gold["vx_disp"], dt, "vx_disp")
    other = np.float32 if dt == np.float64 else np.float64
    got = kd.sph_mean(vel.astype(other), case.k)
    assert got.dtype == other
    assert_close(got, gold["v_mean_othertype"], np.float32, "v_mean other dtype")
    # strided quantity view (column of an N x 3 array), as pynbody passes sim['vx']
    assert_close(kd.sph_mean(vel[:, 0], case.k), gold["vx_mean"], dt, "vx_mean strided")

    # neighbour lists
    idx, d2, h = kd.all_nn_arrays(case.k)
    assert np.array_equal(h, gold["smooth"])
    assert idx.shape == (len(pos), case.k)
    for i, want, d2max in zip(gold["nn_i"], gold["nn_idx"], gold["nn_d2max"]):
        row = idx[int(i)]
        assert d2[int(i)].max() == d2max
        if not np.array_equal(np.sort(row), want):
            # only entries tied with the k-th distance may differ (pq.h:110 strict '<')
            diff = np.setxor1d(row, want)
            dd = ((pos[diff].astype(np.float64) - pos[int(i)].astype(np.float64)) ** 2).sum(axis=1)
            assert np.allclose(dd, float(d2max), rtol=1e-6), "neighbour set differs beyond exact ties"
        assert int(i) in row and d2[int(i)][list(row).index(int(i))] == 0.0
    ball = kd.particles_in_sphere(case.ball_centre, case.ball_radius)
    assert np.array_equal(ball, gold["ball"])


def test_nn_generator_contract():
    """kdtree_test.py:174-212 of the reference: every list has exactly k entries incl. self at d2=0."""
    pos, vel, mass = synth.uniform_box(500, 3, np.float64)
    kd = _tree(pos, mass, 1.0)
    sm = np.empty(500)
    kd.set_array_ref("smooth", sm)
    seen = 0
    for i, h, nbrs, d2 in kd.nn(32):
        assert len(nbrs) == 32 and len(d2) == 32
        assert i in nbrs and d2[nbrs.index(i)] == 0.0
        assert h == 0.5 * np.sqrt(max(d2))
        seen += 1
    assert seen == 500
    assert np.all(sm > 0)


@pytest.mark.parametrize("gen,n,dtype,boxsize,k,kernel_id", [
    ("uniform", 200_000, np.float64, 1.0, 32, 0),
    ("uniform", 200_000, np.float32, 1.0, 32, 0),
    ("zeldovich", 64 ** 3, np.float64, 1.0, 64, 1),
    ("zeldovich", 64 ** 3, np.float32, 1.0, 64, 1),
    ("clumpy", 150_000, np.float32, 1.0, 48, 1),
    ("blobs", 100_000, np.float64, None, 32, 0),
])
def test_matches_oracle_medium(gen, n, dtype, boxsize, k, kernel_id):
    from oracle import oracle
    if gen == "uniform":
        pos, vel, mass = synth.uniform_box(n, 101, dtype)
    elif gen == "zeldovich":
        pos, vel, mass = synth.zeldovich_box(int(round(n ** (1 / 3))), 102, dtype)
    elif gen == "clumpy":
        pos, vel, mass = synth.clumpy_box(n, 103, dtype)
    else:
        pos, vel, mass = synth.gaussian_blobs(n, 104, dtype)
    ref = oracle.OracleKDTree(pos, mass, 16, boxsize, None, kernel_id)
    sm_ref = ref.smooth(k)
    rho_ref = ref.rho(k)
    kd = _tree(pos, mass, boxsize, kernel_id)
    sm = _smooth(kd, pos, k)
    nbad = int(np.count_nonzero(sm != sm_ref))
    assert nbad == 0, f"{nbad} smoothing lengths differ from the oracle"
    rho = _rho(kd, pos, mass, sm, k)
    assert_close(rho, rho_ref, dtype, "rho")
    kd.set_array_ref("rho", rho_ref)
    for op, fn in (("mean", kd.sph_mean), ("div", kd.sph_divergence), ("curl", kd.sph_curl)):
        assert_close(fn(vel, k), ref.qty_op(op, vel, k), dtype, op)
    assert_close(kd.sph_dispersion(vel, k)[:, 0], ref.qty_op("disp", vel, k), dtype, "disp")


@pytest.mark.parametrize("npart", [1, 10, 100, 1000, 100000])
@pytest.mark.parametrize("offset", [0.0, 0.2, 0.5])
@pytest.mark.parametrize("radius", [0.1, 0.3, 1.0])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_particles_in_sphere(npart, offset, radius, dtype):
    """The reference's own test (tests/kdtree_test.py:258-282): index set == brute force after wrap."""
    np.random.seed(1337)
    pos = np.random.uniform(low=-0.5, high=0.5, size=(npart, 3)).astype(dtype)
    kd = _tree(pos, np.ones(npart, dtype), 1.0)
    got = kd.particles_in_sphere([offset, 0.0, 0.0], radius)
    p = pos.copy()
    p[:, 0] -= dtype(offset)
    p[p > 0.5] -= 1.0
    p[p < -0.5] += 1.0
    want = np.where(np.sqrt((p ** 2).sum(axis=1)) < radius)[0]
    assert np.array_equal(np.sort(got), want)


def test_k_larger_than_n_raises():                       # kdtree_test.py:116-121
    pos = np.random.default_rng(0).uniform(size=(16, 3))
    kd = _tree(pos, np.ones(16), None)
    sm = np.empty(16)
    kd.set_array_ref("smooth", sm)
    with pytest.raises(ValueError):
        kd.populate("hsm", 32)


def test_boxsize_too_small_warns_and_disables_periodicity():   # kdtree_test.py:322-329
    rng = np.random.default_rng(0)
    pos = rng.normal(scale=1.0, size=(1000, 3))
    mass = np.ones(1000)
    kd = _tree(pos, mass, 0.1)
    sm = np.empty(1000)
    kd.set_array_ref("smooth", sm)
    with pytest.warns(RuntimeWarning, match="span a region larger than the specified boxsize"):
        kd.populate("hsm", 32)
    kd2 = _tree(pos, mass, None)
    assert np.array_equal(sm, _smooth(kd2, pos, 32))


def test_float64_rounding_at_box_edge():                  # kdtree_test.py:332-345
    rng = np.random.default_rng(5)
    pos = rng.uniform(size=(1000, 3))
    pos[0, 0] = np.nextafter(1.0, -1.0)
    pos[1, 0] = 0.0
    kd = _tree(pos, np.ones(1000), float(np.nextafter(1.0, -1.0)))
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        sm = _smooth(kd, pos, 32)
    assert np.all(np.isfinite(sm)) and np.all(sm > 0)


def test_dtype_errors():
    pos = np.random.default_rng(1).uniform(size=(100, 3))
    from pynbody_b200.kdtree import KDTree
    with pytest.raises(ValueError):
        KDTree(pos.astype(np.int32), np.ones(100, np.int32))
    kd = _tree(pos, np.ones(100), None)
    with pytest.raises(TypeError):
        kd.set_array_ref("smooth", np.empty(100, np.float32))
    with pytest.raises(ValueError):
        kd.set_array_ref("nonsense", np.empty(100))


def test_gather_overflow_raises_runtime_error():          # smooth.h:253-267, kdmain.cpp:795-800
    pos, vel, mass = synth.uniform_box(5000, 9, np.float64)
    kd = _tree(pos, mass, 1.0)
    sm = np.full(5000, 0.2)                                # ~ 2000 particles inside 2h >> k + 500
    with pytest.raises(RuntimeError, match="Buffer overflow"):
        _rho(kd, pos, mass, sm, 32)


def test_device_resident_inputs():
    """torch CUDA tensors are accepted as device buffers (no host round trip)."""
    import torch
    pos, vel, mass = synth.uniform_box(50_000, 4, np.float32)
    kd_h = _tree(pos, mass, 1.0)
    sm_h = _smooth(kd_h, pos, 32)
    rho_h = _rho(kd_h, pos, mass, sm_h, 32)
    tp, tm = torch.from_numpy(pos).cuda(), torch.from_numpy(mass).cuda()
    kd = _tree(tp, tm, 1.0)
    sm = torch.empty(len(pos), dtype=torch.float32, device="cuda")
    kd.set_array_ref("smooth", sm)
    kd.populate("hsm", 32)
    rho = torch.empty_like(sm)
    kd.set_array_ref("mass", tm)
    kd.set_array_ref("rho", rho)
    kd.populate("rho", 32)
    torch.cuda.synchronize()
    assert np.array_equal(sm.cpu().numpy(), sm_h)
    assert np.array_equal(rho.cpu().numpy(), rho_h)


def _bruteforce_smooth(pos, k, L):
    """Exact minimum-image kNN with the library's arithmetic: per axis the nearest of q, fl(q+L), fl(q-L);
    d2 = (dx*dx + dy*dy) + dz*dz in the array dtype; h = 0.5*sqrt(d2_k)."""
    dt = pos.dtype.type
    q = pos[:, None, :]
    x = pos[None, :, :]
    d = q - x
    if L is not None:
        dp = (q + dt(L)) - x
        dm = (q - dt(L)) - x
        d = np.where(np.abs(dp) < np.abs(d), dp, d)
        d = np.where(np.abs(dm) < np.abs(d), dm, d)
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    kth = np.sort(d2, axis=1)[:, k - 1]
    return (dt(0.5) * np.sqrt(kth)).astype(pos.dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_small_and_ragged_cases(dtype):
    """Ragged sizes around the bucket size, k from 1 to N, periodic and open, every reduction.

    Open boxes: against the oracle.  Tiny periodic boxes: the reference's per-cell image choice makes its
    answer depend on its own tree there, so the bar is the exact minimum-image
    neighbour search (cells wider than half the box switch to a per-particle image)."""
    from oracle import oracle
    rng = np.random.default_rng(77)
    for n in (2, 5, 31, 32, 33, 64, 65, 97, 300):
        pos = rng.uniform(0.0, 1.0, (n, 3)).astype(dtype)
        mass = rng.uniform(0.5, 1.5, n).astype(dtype)
        vel = rng.normal(size=(n, 3)).astype(dtype)
        for k in sorted({1, 2, min(n, 7), min(n, 40), n}):
            kd = _tree(pos, mass, 1.0, 1)
            assert np.array_equal(_smooth(kd, pos, k), _bruteforce_smooth(pos, k, 1.0)), (n, k, "periodic")
            ref = oracle.OracleKDTree(pos, mass, 16, None, 1, 1)
            kd = _tree(pos, mass, None, 1)
            sm_ref = ref.smooth(k)
            sm = _smooth(kd, pos, k)
            assert np.array_equal(sm, sm_ref), (n, k, "open")
            assert np.array_equal(sm, _bruteforce_smooth(pos, k, None))
            if k < 2:
                continue            # h = 0: densities are inf/nan in both implementations
            rho_ref = ref.rho(k)
            rho = _rho(kd, pos, mass, sm, k)
            assert_close(rho, rho_ref, dtype, f"rho n={n} k={k}")
            if k < 7:
                continue            # a couple of neighbours sitting at the kernel edge: sums are rounding noise
            kd.set_array_ref("rho", rho_ref)
            assert_close(kd.sph_mean(vel, k), ref.qty_op("mean", vel, k), dtype, "mean")
            assert_close(kd.sph_divergence(vel, k), ref.qty_op("div", vel, k), dtype, "div")


def test_large_k():
    """k well above the bucket size (generic-depth queue path) and the k limit."""
    from oracle import oracle
    pos, vel, mass = synth.uniform_box(20_000, 21, np.float64)
    ref = oracle.OracleKDTree(pos, mass, 16, 1.0, None, 0)
    kd = _tree(pos, mass, 1.0)
    for k in (200, 400):
        assert np.array_equal(_smooth(kd, pos, k), ref.smooth(k)), k
    sm = np.empty(len(pos))
    kd.set_array_ref("smooth", sm)
    with pytest.raises(ValueError, match="too large"):
        kd.populate("hsm", 5000)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_full_size_box_properties(dtype):
    """BASELINE.json's headline workload at full size (256^3 = 16 777 216 particles, clustered periodic box,
    k = 64, float64, device-resident), checked through size-independent properties:

    * exactness on a sample: for 256 random particles the k-th smallest minimum-image distance over ALL
      particles (torch shortlist, then the reference's operation order in numpy) gives h bit for bit;
    * permutation invariance: the same particles in a shuffled order give the same h bit for bit and the
      same density to rounding (the tree, the buckets and the visit order all change);
    * the fused density equals the density of the general gather pass (forced by non-uniform masses that are
      equal to rounding) to 1e-12.
    """
    import torch
    k, L = 64, 1.0
    pos, _, mass = synth.zeldovich_box_device(256, 2, np.float64)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    pos, mass = pos.to(tdt).contiguous(), mass.to(tdt).contiguous()
    rtol = 1e-12 if dtype == np.float64 else 2e-5
    n = pos.shape[0]
    assert n == 1 << 24

    def run(p, m):
        kd = _tree(p, m, L)
        kd.set_kernel("WendlandC2Kernel")
        sm = torch.empty(n, dtype=tdt, device="cuda")
        rho = torch.empty_like(sm)
        kd.set_array_ref("smooth", sm)
        kd.populate("hsm", k)
        kd.set_array_ref("mass", m)
        kd.set_array_ref("rho", rho)
        kd.populate("rho", k)
        torch.cuda.synchronize()
        return sm, rho

    sm, rho = run(pos, mass)
    assert bool(torch.isfinite(sm).all()) and bool((sm > 0).all()) and bool((rho > 0).all())

    # exactness on a sample
    g = torch.Generator(device="cpu").manual_seed(11)
    sample = torch.randint(0, n, (256,), generator=g)
    for i in sample.tolist():
        d = pos - pos[i]
        d = d - torch.round(d / L) * L                         # shortlist only: exact values follow below
        cand = torch.topk((d * d).sum(dim=1), k + 64, largest=False).indices
        q, x = pos[i].cpu().numpy(), pos[cand].cpu().numpy()     # (the shortlist contains the query itself)
        dd = _bruteforce_d2(q, x, L)
        want = dtype(0.5) * np.sqrt(np.sort(dd)[k - 1])
        assert float(sm[i]) == float(want), (i, float(sm[i]), float(want))

    # permutation invariance
    perm = torch.randperm(n, generator=torch.Generator(device="cuda").manual_seed(5), device="cuda")
    sm_p, rho_p = run(pos[perm].contiguous(), mass[perm].contiguous())
    assert bool((sm_p == sm[perm]).all())
    assert float(((rho_p - rho[perm]).abs() / rho[perm]).max()) < rtol

    # fused density vs the general gather pass
    m2 = mass.clone()
    m2[::2] = torch.nextafter(m2[::2], torch.full_like(m2[::2], 2.0))     # masses no longer identical
    _, rho_g = run(pos, m2)
    assert float(((rho_g - rho).abs() / rho).max()) < rtol


@pytest.mark.parametrize("n,dtype", [(33, np.float64), (2047, np.float32), (2049, np.float64), (65_537, np.float64),
                                     (150_000, np.float32), (1_000_003, np.float64)])
def test_levelwise_build_variants_give_the_same_tree(n, dtype):
    """tree_build 2 (default: global fused passes + the lowest levels from the presorted lists in shared memory), 3 (mark +
    CUB prefix sum + partition per level) and 4 (fused pass for every level) are the same algorithm: identical particle
    order and identical bounds of every node."""
    from pynbody_b200 import _lib
    pos, vel, mass = synth.clumpy_box(n, 5 + n % 11, dtype)
    if n == 65_537:
        # coordinates closer than 2^-32 of the extent, exact duplicates and signed zeros: the runs of equal 32-bit sort keys
        # that the float64 build finishes by rank counting (tree_build.cu: k_fix_ties)
        pos[1000:2000] = pos[:1000] + 1e-13
        pos[2000:2500, 0] = pos[:500, 0]
        pos[3000:3010, 1] = 0.0
        pos[3010:3020, 1] = -0.0
    if n == 150_000:
        pos = pos.astype(np.float64)            # a lattice along x and y: long runs of equal keys -> 64-bit sort for those axes
        mass = mass.astype(np.float64)
        pos[:, 0] = np.round(pos[:, 0] * 40) / 40
        pos[:, 1] = np.round(pos[:, 1] * 3000) / 3000
    L = _lib.lib()
    got = {}
    for build in (2, 3, 4):
        kdmain.set_option("tree_build", build)
        try:
            kd = _tree(pos, mass, 1.0 if n > 1000 else None)
        finally:
            kdmain.set_option("tree_build", 2)
        h = kd.kdtree.handle
        nb = L.pnb_tree_num_buckets(h)
        boxes = np.empty((2 * nb, 6))
        _lib.check(L.pnb_tree_node_bounds(h, boxes.ctypes.data))
        order = np.empty(n, np.int64)
        _lib.check(L.pnb_tree_order(h, order.ctypes.data))
        got[build] = (order, boxes[1:])
    for build in (3, 4):
        assert np.array_equal(got[2][0], got[build][0]), build
        assert np.array_equal(got[2][1], got[build][1]), build


def _bruteforce_d2(q, x, L):
    """d2 of query q[3] to particles x[m,3] with the library's arithmetic (nearest of q, fl(q+L), fl(q-L) per axis)."""
    dt = x.dtype.type
    d = q[None, :] - x
    dp = (q[None, :] + dt(L)) - x
    dm = (q[None, :] - dt(L)) - x
    d = np.where(np.abs(dp) < np.abs(d), dp, d)
    d = np.where(np.abs(dm) < np.abs(d), dm, d)
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


def test_pinned_host_arrays_take_the_overlapped_paths():
    """Page-locked host arrays (what bench.py's e2e leg uses): the masses are uploaded behind the tree build and
    the fused density starts its way back before the caller's smooth array has been verified.  Results must
    be those of ordinary numpy arrays, also when the smooth array was changed in between (the early copy
    is then overwritten by the gather pass)."""
    import torch
    n, k = 700_000, 32                                    # > 4 MiB per array: below that the plain path is used
    pos, vel, mass = synth.clumpy_box(n, 21, np.float64)

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.copy_(torch.from_numpy(a))
        return t.numpy()

    def run(p, m, sm, rho, perturb):
        kd = _tree(p, m, 1.0)
        kd.set_array_ref("smooth", sm)
        kd.populate("hsm", k)
        if perturb:
            sm[::1000] *= 1.01
        kd.set_array_ref("mass", m)
        kd.set_array_ref("rho", rho)
        kd.set_array_ref("smooth", sm)
        kd.populate("rho", k)
        return sm.copy(), rho.copy()

    for perturb in (False, True):
        sm_a, rho_a = run(pos, mass, np.empty(n), np.empty(n), perturb)
        sm_b, rho_b = run(pinned(pos), pinned(mass), pinned(np.empty(n)), pinned(np.zeros(n)), perturb)
        assert np.array_equal(sm_a, sm_b)
        assert np.array_equal(rho_a, rho_b)
    # and the perturbed run really differs from the unperturbed one where h was changed
    assert not np.array_equal(rho_a[::1000], run(pos, mass, np.empty(n), np.empty(n), False)[1][::1000])
