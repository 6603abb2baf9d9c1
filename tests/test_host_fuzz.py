"""Seeded random sweep of the plan layer's host logic (dry plans on simulated ranks, CPU only):
random global sizes, rank counts, process grids, uneven user pencils and brick cuts.  Every case
must reproduce the oracle's decomposition / exchange tables and, for every transposition and
reshape, the MPI-datatype truth (global-array slicing) -- the same checks as tests/test_plan_host.py
and tests/test_reshape_shortcuts.py, on shapes nobody picked by hand."""
import numpy as np
import pytest

from dtfft_b200.plan import Config, Executor, Layout, Pencil, PlanC2C, PlanR2C, Precision
from oracle import datatype_path as D
from oracle import layout as L
from oracle import pipeline as P
from tests.test_plan_host import LAYOUT_OF_PENCIL, dry_world, replay_fused
from tests.test_reshape_shortcuts import check_brick_case

RANKS = [1, 2, 3, 4, 6, 8]


def _factor_pairs(n):
    return [(a, n // a) for a in range(1, n + 1) if n % a == 0]


def _default_cases(n_cases, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_cases):
        nd = int(rng.choice([2, 3], p=[0.25, 0.75]))
        nranks = int(rng.choice(RANKS))
        dims = [int(v) for v in rng.integers(max(2, nranks), 40, size=nd)]
        z_slab = bool(rng.integers(0, 2))
        cart = None
        if nd == 3 and nranks > 1 and rng.integers(0, 2):
            g1, g2 = _factor_pairs(nranks)[int(rng.integers(0, len(_factor_pairs(nranks))))]
            cart = [1, g1, g2]
        pipelined = bool(rng.integers(0, 2))
        out.append((dims, nranks, z_slab, cart, pipelined))
    return out


@pytest.mark.parametrize("dims,nranks,z_slab,cart,pipelined", _default_cases(40, 20261017))
def test_random_default_and_cart_decompositions(dims, nranks, z_slab, cart, pipelined):
    nd = len(dims)
    cfg = Config(enable_z_slab=z_slab, backend=27 if pipelined else 24)
    plans = dry_world(nranks, lambda r, c: PlanC2C(list(dims), comm=c, config=cfg, dry=True), cart_dims=cart)
    comm_dims = plans[0].grid_dims
    if cart is not None:
        assert comm_dims == cart
    else:
        want_dims, is_z, _ = L.choose_grid(list(dims), nranks, cuda=True, z_slab=z_slab, y_slab=False)
        assert comm_dims == want_dims and plans[0].z_slab_enabled == is_z
    is_z = plans[0].z_slab_enabled
    G = P.global_array(dims, np.complex128, kind="index")
    for r, plan in enumerate(plans):
        gold = L.make_pencils(list(dims), comm_dims, r)
        for d in range(nd):
            got = plan.get_pencil(LAYOUT_OF_PENCIL[d])
            assert (got.starts, got.counts) == (gold[d].starts, gold[d].counts), (r, d)
    ttypes = [1, -1] if nd == 2 else [1, -1, 2, -2] + ([3, -3] if is_z else [])
    for t in ttypes:
        _, geos = L.plan_geometry(list(dims), comm_dims, t, pipelined=pipelined)
        for r, plan in enumerate(plans):
            d, g = plan.describe_exchange(t), geos[r]
            assert d["members"] == g.members and d["me"] == g.comm_rank and d["pack_kernel"] == g.pack_kernel
            if g.comm_size > 1:
                assert d["unpack_kernel"] == g.unpack_kernel
                assert np.array_equal(d["send_nd"], g.send_nd) and np.array_equal(d["recv_nd"], g.recv_nd)
                assert d["send_counts"].tolist() == g.send_counts and d["send_displs"].tolist() == g.send_displs
                assert d["recv_counts"].tolist() == g.recv_counts and d["recv_displs"].tolist() == g.recv_displs
        src = P.scatter_input(G, list(dims), comm_dims, t)
        want = P.transpose_datatype(G, list(dims), comm_dims, t)
        # one-kernel NVLink path
        got = replay_fused(plans, t, src, [w.size for w in want], np.complex128)
        # three-step NCCL path with the (identical) reference tables
        gen = P.transpose_generic(src, list(dims), comm_dims, t, pipelined=pipelined)
        # the reference's host MPI-datatype path itself (its derived datatypes restated)
        allp = [L.make_pencils(list(dims), comm_dims, r) for r in range(nranks)]
        si, ri = L.transpose_pencil_ids(t)
        groups = [L.comm_members(r, comm_dims, L.transpose_comm_id(t)) for r in range(nranks)]
        dt = D.exchange(src, [p[si] for p in allp], [p[ri] for p in allp], groups, 16, ttype=t,
                        mode=D.PACK if pipelined else D.UNPACK)
        for r in range(nranks):
            assert np.array_equal(got[r], want[r]), ("fused", L.TRANSPOSE_NAMES[t], r)
            assert np.array_equal(gen[r], want[r]), ("generic", L.TRANSPOSE_NAMES[t], r)
            assert np.array_equal(dt[r], want[r]), ("datatype", L.TRANSPOSE_NAMES[t], r)
    Config()._commit()


def _random_cuts(rng, total, parts):
    """`parts` positive extents summing to `total`, uneven."""
    if parts == 1:
        return [total]
    marks = sorted(rng.choice(np.arange(1, total), size=parts - 1, replace=False).tolist())
    edges = [0] + marks + [total]
    return [edges[i + 1] - edges[i] for i in range(parts)]


def _user_pencil_cases(n_cases, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_cases):
        nranks = int(rng.choice([2, 3, 4, 6, 8]))
        py, pz = _factor_pairs(nranks)[int(rng.integers(0, len(_factor_pairs(nranks))))]
        dims = [int(rng.integers(4, 24)), int(rng.integers(max(py, 2), 24)), int(rng.integers(max(pz, 2), 24))]
        ycuts, zcuts = _random_cuts(rng, dims[1], py), _random_cuts(rng, dims[2], pz)
        out.append((dims, ycuts, zcuts))
    return out


@pytest.mark.parametrize("dims,ycuts,zcuts", _user_pencil_cases(25, 7))
def test_random_user_pencils(dims, ycuts, zcuts):
    ye, ze = np.concatenate([[0], np.cumsum(ycuts)]), np.concatenate([[0], np.cumsum(zcuts)])
    boxes = [([0, int(ye[j]), int(ze[k])], [dims[0], int(ycuts[j]), int(zcuts[k])])
             for k in range(len(zcuts)) for j in range(len(ycuts))]
    n = len(boxes)
    cfg = Config(enable_z_slab=False)
    plans = dry_world(n, lambda r, c: PlanC2C(Pencil(*boxes[r]), comm=c, config=cfg, dry=True))
    starts, counts = [b[0] for b in boxes], [b[1] for b in boxes]
    ggrid, coords = L.grid_from_boxes(starts, counts)
    assert ggrid == [1, len(ycuts), len(zcuts)]
    G = P.global_array(dims, np.float64, kind="index")
    pencils = []
    for r, plan in enumerate(plans):
        assert plan.dims == list(dims) and plan.grid_dims == ggrid
        gold = L.pencils_from_x(list(dims), ggrid, coords[r], starts[r], counts[r])
        pencils.append(gold)
        for d in range(3):
            got = plan.get_pencil(LAYOUT_OF_PENCIL[d])
            assert (got.starts, got.counts) == (gold[d].starts, gold[d].counts), (r, d)
    # partition check of the reference's Python test (tests/python/test_pencil_api_py.py:76-86): every layout
    # tiles the global grid
    for lay in LAYOUT_OF_PENCIL:
        assert sum(p.get_pencil(lay).size for p in plans) == int(np.prod(dims)), lay
    for t in (1, -1, 2, -2):
        si, ri = L.transpose_pencil_ids(t)
        src = [P.pencil_slice(G, pencils[r][si]) for r in range(n)]
        want = P.redistribute(G, [pencils[r][ri] for r in range(n)])
        got = replay_fused(plans, t, src, [w.size for w in want], np.float64)
        for r in range(n):
            assert np.array_equal(got[r], want[r]), (t, r)
    Config()._commit()


def _brick_cases(n_cases, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n_cases:
        nd = int(rng.choice([2, 3], p=[0.3, 0.7]))
        shape = [int(rng.integers(2, 5))] + [int(rng.integers(1, 3)) for _ in range(nd - 1)]
        if int(np.prod(shape)) > 8:
            continue
        # long axes now and then, so that the z split (> 32 * bricks along x) and the y split occur
        tot = [int(rng.integers(8 * shape[0], 40))] + [int(rng.choice([rng.integers(8 * shape[d], 48), rng.integers(140, 200)]))
                                                       for d in range(1, nd)]
        cuts = [_random_cuts(rng, tot[d], shape[d]) for d in range(nd)]
        if min(min(c) for c in cuts) < 4:
            continue
        out.append((cuts, bool(rng.integers(0, 2))))
    return out


@pytest.mark.parametrize("cuts,pipelined", _brick_cases(25, 11))
def test_random_bricks(cuts, pipelined):
    check_brick_case(cuts, pipelined)


def _r2c_cases(n_cases, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_cases):
        nd = int(rng.choice([2, 3]))
        nranks = int(rng.choice([1, 2, 4, 6]))
        dims = [int(v) for v in rng.integers(2 * max(2, nranks), 36, size=nd)]
        out.append((dims, nranks, bool(rng.integers(0, 2)), bool(rng.integers(0, 2))))
    return out


@pytest.mark.parametrize("dims,nranks,z_slab,single", _r2c_cases(16, 3))
def test_random_r2c_plans(dims, nranks, z_slab, single):
    """R2C: the real X pencil keeps the user's dims, every complex pencil has nx/2 + 1 points along x
    (src/dtfft_plan.F90:2626-2630); sizes are counted in REAL elements (:1350-1355, 1868-1876) and every
    transposition of the complex side reproduces the datatype path."""
    nd = len(dims)
    prec = Precision.SINGLE if single else Precision.DOUBLE
    cfg = Config(enable_z_slab=z_slab)
    plans = dry_world(nranks, lambda r, c: PlanR2C(list(dims), comm=c, precision=prec, executor=Executor.CUFFT,
                                                   config=cfg, dry=True))
    cdims = [dims[0] // 2 + 1] + list(dims[1:])
    comm_dims = plans[0].grid_dims
    want_dims, is_z, _ = L.choose_grid(cdims, nranks, cuda=True, z_slab=z_slab, y_slab=False)
    assert comm_dims == want_dims and plans[0].z_slab_enabled == is_z
    es_real = 4 if single else 8
    for r, plan in enumerate(plans):
        gold = L.make_pencils(cdims, comm_dims, r)
        real = plan.get_pencil(Layout.X_PENCILS)
        assert real.counts == [dims[0]] + gold[0].counts[1:] and real.starts == gold[0].starts
        four = plan.get_pencil(Layout.X_PENCILS_FOURIER)
        assert (four.starts, four.counts) == (gold[0].starts, gold[0].counts)
        for d in range(1, nd):
            got = plan.get_pencil(LAYOUT_OF_PENCIL[d])
            assert (got.starts, got.counts) == (gold[d].starts, gold[d].counts), (r, d)
        assert plan.element_size == es_real
        assert plan.alloc_size == max(int(np.prod(real.counts)), 2 * max(p.size for p in gold))
        assert plan.alloc_bytes == plan.alloc_size * es_real
    G = P.global_array(cdims, np.complex128, kind="index")
    ttypes = [1, -1] if nd == 2 else [1, -1, 2, -2] + ([3, -3] if is_z else [])
    for t in ttypes:
        src = P.scatter_input(G, cdims, comm_dims, t)
        want = P.transpose_datatype(G, cdims, comm_dims, t)
        got = replay_fused(plans, t, src, [w.size for w in want], np.complex128)
        for r in range(nranks):
            assert np.array_equal(got[r], want[r]), (L.TRANSPOSE_NAMES[t], r)
    Config()._commit()
