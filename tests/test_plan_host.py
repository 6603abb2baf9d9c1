"""Host logic of the plan layer on CPU: decomposition, sizes and exchange geometry computed by
the C++ library (metadata-only "dry" plans, dtfftb_plan_create_dry) against the oracle, at
1..8 simulated ranks (tests/fake_comm.py).

Ground truth for every exchange is what the reference's host MPI-datatype path delivers
(src/dtfft_reshape_handle_datatype.F90:436-847) = slicing a global array; the product's fused
boxes are replayed in numpy (oracle.pipeline.apply_boxes) and must reproduce it bit for bit.
"""
import os
import numpy as np
import pytest

from dtfft_b200.plan import (Config, DtfftError, Executor, Layout, Pencil, PlanC2C, PlanR2C, PlanR2R, Precision,
                             Reshape, Transpose)
from oracle import kernels as K
from oracle import layout as L
from oracle import pipeline as P
from tests.fake_comm import ThreadWorld

LAYOUT_OF_PENCIL = [Layout.X_PENCILS, Layout.Y_PENCILS, Layout.Z_PENCILS]


def dry_world(nranks, make_plan, cart_dims=None):
    """Create one dry plan per simulated rank; returns them (keep alive!)."""
    if nranks == 1:
        return [make_plan(0, None)]
    return ThreadWorld(nranks).run(make_plan, cart_dims)


def collect(plans, fn):
    if len(plans) == 1:
        return [fn(plans[0])]
    # describe_exchange etc. are local calls: no collective inside
    return [fn(p) for p in plans]


def pencil_of(plan, layout):
    p = plan.get_pencil(layout)
    return L.Pencil(p.dim, p.starts, p.counts)


DEFAULT_CASES = [((64, 64, 64), 4), ((512, 512, 512), 8), ((512, 512, 512), 1), ((48, 21, 36), 6), ((13, 7, 9), 6),
                 ((40, 33, 28), 4), ((129, 99, 33), 3), ((37, 22), 3), ((16384, 16384), 8), ((90, 57), 5),
                 ((66, 10, 12), 4), ((128, 64, 96), 8), ((1024, 1024, 1024), 8)]


@pytest.mark.parametrize("dims,nranks", DEFAULT_CASES)
@pytest.mark.parametrize("z_slab", [True, False])
def test_default_decomposition_matches_oracle(dims, nranks, z_slab):
    """Grid choice, pencils, local sizes (src/dtfft_transpose_plan.F90:170-203, 1084-1131;
    src/dtfft_pencil.F90:255-279, 438-463)."""
    cfg = Config(enable_z_slab=z_slab)
    plans = dry_world(nranks, lambda r, c: PlanC2C(list(dims), comm=c, config=cfg, dry=True))
    comm_dims, is_z, is_y = L.choose_grid(list(dims), nranks, cuda=True, z_slab=z_slab, y_slab=False)
    nd = len(dims)
    for r, plan in enumerate(plans):
        assert plan.grid_dims == comm_dims
        assert plan.z_slab_enabled == is_z and plan.y_slab_enabled == is_y
        gold = L.make_pencils(list(dims), comm_dims, r)
        for d in range(nd):
            got = plan.get_pencil(LAYOUT_OF_PENCIL[d])
            assert (got.starts, got.counts, got.dim) == (gold[d].starts, gold[d].counts, d + 1)
        ins, inc, outs, outc, alloc = plan.local_sizes
        assert (ins, inc) == (gold[0].starts, gold[0].counts)
        assert (outs, outc) == (gold[nd - 1].starts, gold[nd - 1].counts)
        assert alloc == max(p.size for p in gold)
        assert plan.alloc_bytes == alloc * 16 and plan.element_size == 16
    Config()._commit()


@pytest.mark.parametrize("dims,nranks", [c for c in DEFAULT_CASES if np.prod(c[0]) <= 2 ** 21])
@pytest.mark.parametrize("pipelined", [False, True])
def test_neighbor_data_matches_oracle(dims, nranks, pipelined):
    """neighbor_data / counts / displs / kernel kinds of reshape_handle_generic%create
    (src/dtfft_reshape_handle_generic.F90:291-640) for every transposition of the plan."""
    cfg = Config(enable_z_slab=True, backend=27 if pipelined else 24)
    plans = dry_world(nranks, lambda r, c: PlanC2C(list(dims), comm=c, config=cfg, dry=True))
    comm_dims, is_z, _ = L.choose_grid(list(dims), nranks)
    nd = len(dims)
    ttypes = [1, -1] if nd == 2 else [1, -1, 2, -2] + ([3, -3] if is_z else [])
    for t in ttypes:
        _, geos = L.plan_geometry(list(dims), comm_dims, t, pipelined=pipelined)
        for r, plan in enumerate(plans):
            d = plan.describe_exchange(t)
            g = geos[r]
            assert d["members"] == g.members and d["me"] == g.comm_rank
            assert d["pack_kernel"] == g.pack_kernel
            if g.comm_size > 1:
                assert d["unpack_kernel"] == g.unpack_kernel
                assert np.array_equal(d["send_nd"], g.send_nd) and np.array_equal(d["recv_nd"], g.recv_nd)
                assert d["send_counts"].tolist() == g.send_counts and d["send_displs"].tolist() == g.send_displs
                assert d["recv_counts"].tolist() == g.recv_counts and d["recv_displs"].tolist() == g.recv_displs
    if pipelined and nranks > 1:
        # aux = alloc + max(send, recv) bytes of the pipelined backend (dtfft_plan.F90:1398-1420)
        assert plans[0].aux_bytes > plans[0].alloc_bytes
    Config()._commit()


def replay_fused(plans, type_, src_bufs, dst_sizes, dtype):
    n = len(plans)
    dsts = [np.full(dst_sizes[r], -7, dtype) for r in range(n)]
    for r, plan in enumerate(plans):
        d = plan.describe_exchange(type_)
        P.apply_boxes(src_bufs[r], dsts, d["fused_boxes"], d["members"])
    return dsts


@pytest.mark.parametrize("dims,nranks", [c for c in DEFAULT_CASES if np.prod(c[0]) <= 2 ** 19])
def test_fused_boxes_reproduce_datatype_path(dims, nranks):
    """One-kernel NVLink path: replaying every rank's boxes must give exactly the destination
    pencils of the MPI-datatype path, for every transposition, both directions."""
    plans = dry_world(nranks, lambda r, c: PlanC2C(list(dims), comm=c, dry=True))
    comm_dims = plans[0].grid_dims
    nd = len(dims)
    G = P.global_array(dims, np.complex128, kind="index")
    ttypes = [1, -1] if nd == 2 else [1, -1, 2, -2] + ([3, -3] if plans[0].z_slab_enabled else [])
    for t in ttypes:
        src = P.scatter_input(G, list(dims), comm_dims, t)
        want = P.transpose_datatype(G, list(dims), comm_dims, t)
        got = replay_fused(plans, t, src, [w.size for w in want], np.complex128)
        for r in range(nranks):
            assert np.array_equal(got[r], want[r]), (L.TRANSPOSE_NAMES[t], r)


@pytest.mark.parametrize("dims,nranks,nchunks", [((48, 21, 36), 6, 3), ((40, 33, 28), 4, 8), ((37, 22), 3, 4),
                                                 ((64, 20, 18), 8, 5), ((16, 12, 5), 2, 8)])
def test_stage_overlap_chunks_tile_the_transposition(dims, nranks, nchunks):
    """Stage overlap of the FFT with the fused exchange (Plan::run_fft_transpose): the source pencil is cut
    along its slowest axis; the boxes of chunk k (dtfftb_plan_describe_chunk), applied to the chunk's part
    of the source, must together deliver exactly the destination pencils of the whole transposition --
    also when there are more chunks than planes (empty chunks) or the cut is uneven."""
    plans = dry_world(nranks, lambda r, c: PlanC2C(list(dims), comm=c, config=Config(enable_z_slab=False), dry=True))
    comm_dims = plans[0].grid_dims
    nd = len(dims)
    G = P.global_array(dims, np.complex128, kind="index")
    for t in ([1, -1] if nd == 2 else [1, -1, 2, -2]):
        src = P.scatter_input(G, list(dims), comm_dims, t)
        want = P.transpose_datatype(G, list(dims), comm_dims, t)
        dsts = [np.full(w.size, -7, np.complex128) for w in want]
        hits = [np.zeros(w.size, np.int32) for w in want]
        for r, plan in enumerate(plans):
            members = plan.describe_exchange(t)["members"]
            covered = 0
            for k in range(nchunks):
                d = plan.describe_chunk(t, k, nchunks)
                part = src[r][d["chunk_offset"]:]
                P.apply_boxes(part, dsts, d["boxes"], members)
                P.apply_boxes(np.ones(part.size, np.int32), hits, d["boxes"], members)  # overwrites: counted below
                covered += int(sum(b[0] * b[1] * b[2] for b in d["boxes"] if b[0] > 0))
            assert covered == src[r].size, (t, r)  # every source element belongs to exactly one chunk box
        for r in range(nranks):
            assert np.array_equal(dsts[r], want[r]), (L.TRANSPOSE_NAMES[t], r)
            assert np.all(hits[r] == 1)
    Config()._commit()


# ---------------------------------------------------------------------------------------------
# user pencils and bricks
# ---------------------------------------------------------------------------------------------
def split_uneven(n, parts):
    """cut points like the reference's C test (3/4 - 1/4 split, tests/c/test_c2c_3d_c.c:128-136)"""
    if parts == 1:
        return [(0, n)]
    if parts == 2:
        a = (3 * n) // 4
        return [(0, a), (a, n - a)]
    out, s = [], 0
    for i in range(parts):
        c = n // parts + (1 if i < n % parts else 0)  # remainder to the FIRST ranks: not dtFFT's own rule
        out.append((s, c))
        s += c
    return out


def user_x_pencils(dims, grid):
    """X pencils (x undistributed) over a py x pz process grid with uneven user cuts."""
    ys, zs = split_uneven(dims[1], grid[0]), split_uneven(dims[2], grid[1])
    boxes = []
    for iz in range(grid[1]):
        for iy in range(grid[0]):
            boxes.append(([0, ys[iy][0], zs[iz][0]], [dims[0], ys[iy][1], zs[iz][1]]))
    return boxes


@pytest.mark.parametrize("dims,grid", [((512, 64, 96), (2, 2)), ((20, 13, 17), (3, 2)), ((33, 9, 8), (1, 4)),
                                       ((16, 12, 40), (4, 1))])
def test_user_pencils(dims, grid):
    """Plans created from dtfft_pencil_t keep the user's split (src/dtfft_pencil.F90:136-163,
    777-908); every transposition still reproduces the datatype path."""
    boxes = user_x_pencils(dims, grid)
    n = len(boxes)
    cfg = Config(enable_z_slab=False)
    plans = dry_world(n, lambda r, c: PlanC2C(Pencil(*boxes[r]), comm=c, config=cfg, dry=True))
    starts, counts = [b[0] for b in boxes], [b[1] for b in boxes]
    ggrid, coords = L.grid_from_boxes(starts, counts)
    assert ggrid == [1, grid[0], grid[1]]
    G = P.global_array(dims, np.float64, kind="index")
    pencils = []
    for r, plan in enumerate(plans):
        assert plan.dims == list(dims) and plan.grid_dims == ggrid
        gold = L.pencils_from_x(list(dims), ggrid, coords[r], starts[r], counts[r])
        pencils.append(gold)
        for d in range(3):
            got = plan.get_pencil(LAYOUT_OF_PENCIL[d])
            assert (got.starts, got.counts) == (gold[d].starts, gold[d].counts), (r, d)
    for t in (1, -1, 2, -2):
        si, ri = L.transpose_pencil_ids(t)
        src = [P.pencil_slice(G, pencils[r][si]) for r in range(n)]
        want = P.redistribute(G, [pencils[r][ri] for r in range(n)])
        got = replay_fused(plans, t, src, [w.size for w in want], np.float64)
        for r in range(n):
            assert np.array_equal(got[r], want[r]), (t, r)
    Config()._commit()


def test_user_pencils_with_an_empty_rank():
    """A rank without points along a split axis (more ranks than points: dtFFT's own get_local_size
    gives rank 0 nothing for 3 points on 4 ranks) joins the 1-D communicator of the neighbour it
    shares its start with (create_1d_comm, src/dtfft_pencil.F90:1066-1068); its kernels are no-ops
    (abstract_kernel.F90:236-240) and every transposition still reproduces the datatype path."""
    dims = (16, 3, 8)
    boxes = [([0, 0, 0], [16, 0, 8]), ([0, 0, 0], [16, 1, 8]), ([0, 1, 0], [16, 1, 8]), ([0, 2, 0], [16, 1, 8])]
    n = len(boxes)
    cfg = Config(enable_z_slab=False)
    plans = dry_world(n, lambda r, c: PlanC2C(Pencil(*boxes[r]), comm=c, config=cfg, dry=True))
    starts, counts = [b[0] for b in boxes], [b[1] for b in boxes]
    ggrid, coords = L.grid_from_boxes(starts, counts)
    assert ggrid == [1, 4, 1] and [c[1] for c in coords] == [0, 1, 2, 3]
    G = P.global_array(dims, np.float64, kind="index")
    pencils = []
    for r, plan in enumerate(plans):
        assert plan.dims == list(dims) and plan.grid_dims == ggrid
        gold = L.pencils_from_x(list(dims), ggrid, coords[r], starts[r], counts[r])
        pencils.append(gold)
        for d in range(3):
            got = plan.get_pencil(LAYOUT_OF_PENCIL[d])
            assert (got.starts, got.counts) == (gold[d].starts, gold[d].counts), (r, d)
    assert pencils[0][0].size == 0 and pencils[0][1].size > 0
    for t in (1, -1, 2, -2):
        si, ri = L.transpose_pencil_ids(t)
        src = [P.pencil_slice(G, pencils[r][si]) for r in range(n)]
        want = P.redistribute(G, [pencils[r][ri] for r in range(n)])
        got = replay_fused(plans, t, src, [w.size for w in want], np.float64)
        for r in range(n):
            assert np.array_equal(got[r], want[r]), (t, r)
        # reference geometry: the empty rank sends / receives nothing
        d0 = plans[0].describe_exchange(t)
        if t == 1:
            assert d0["send_counts"].sum() == 0 and d0["recv_counts"].sum() == pencils[0][1].size
        if t == -1:
            assert d0["recv_counts"].sum() == 0
    Config()._commit()


def brick_boxes(cuts):
    """cuts = per axis list of extents; rank order x fastest."""
    edges = [np.concatenate([[0], np.cumsum(c)]) for c in cuts]
    boxes = []
    nd = len(cuts)
    if nd == 3:
        for k in range(len(cuts[2])):
            for j in range(len(cuts[1])):
                for i in range(len(cuts[0])):
                    boxes.append(([int(edges[0][i]), int(edges[1][j]), int(edges[2][k])],
                                  [int(cuts[0][i]), int(cuts[1][j]), int(cuts[2][k])]))
    else:
        for j in range(len(cuts[1])):
            for i in range(len(cuts[0])):
                boxes.append(([int(edges[0][i]), int(edges[1][j])], [int(cuts[0][i]), int(cuts[1][j])]))
    return boxes


BRICK_CASES = [
    # BASELINE config 5: 768x512x1024, brick grid 2x2x2, uneven non-power-of-two cuts
    ([[300, 468], [200, 312], [500, 524]], True),
    ([[30, 34], [20, 12], [70, 58]], False),          # z long enough at tile 32? no -> y / factorised split
    ([[10, 6, 8], [9, 11], [12, 8]], False),          # 3 bricks along x
    ([[10, 6], [40, 24], [5, 7]], False),
    ([[20, 13, 7], [16, 17]], False),                 # 2-D bricks
]


@pytest.mark.parametrize("cuts,metadata_only", BRICK_CASES)
def test_bricks_to_pencils(cuts, metadata_only):
    """from_bricks + the four reshapes (src/dtfft_pencil.F90:520-775, src/dtfft_reshape_plan.F90:
    150-182): X pencils, Z bricks and the reshape exchanges against the oracle."""
    boxes = brick_boxes(cuts)
    n = len(boxes)
    nd = len(cuts)
    cfg = Config(enable_fourier_reshape=True, enable_z_slab=False)
    plans = dry_world(n, lambda r, c: PlanR2R(Pencil(*boxes[r]), comm=c, config=cfg, dry=True))
    starts, counts = [b[0] for b in boxes], [b[1] for b in boxes]
    dims, comm_dims, coords, xs, xc, bgrid, _ = L.from_bricks(starts, counts)
    pencils = [L.pencils_from_x(dims, comm_dims, coords[r], xs[r], xc[r]) for r in range(n)]
    zb = L.z_bricks(dims, comm_dims, coords, [p[nd - 1] for p in pencils], bgrid)
    for r, plan in enumerate(plans):
        assert plan.dims == dims and plan.grid_dims == comm_dims
        b1 = plan.get_pencil(Layout.X_BRICKS)
        assert (b1.starts, b1.counts) == (starts[r], counts[r])
        for d in range(nd):
            got = plan.get_pencil(LAYOUT_OF_PENCIL[d])
            assert (got.starts, got.counts) == (pencils[r][d].starts, pencils[r][d].counts), (r, d)
        b2 = plan.get_pencil(Layout.Z_BRICKS)
        assert (b2.starts, b2.counts) == (zb[r].starts, zb[r].counts), r
        ins, inc, outs, outc, alloc = plan.local_sizes
        assert (ins, inc) == (starts[r], counts[r])
        assert alloc >= max(int(np.prod(counts[r])), max(p.size for p in pencils[r]), zb[r].size)
    if nd == 3 and cuts == BRICK_CASES[0][0]:
        # pencil grid 1x2x4, X pencils 768 x {200|312} x {250|250|262|262}
        assert comm_dims == [1, 2, 4]
        assert sorted({tuple(p[0].counts) for p in pencils}) == [(768, 200, 250), (768, 200, 262), (768, 312, 250),
                                                                  (768, 312, 262)]
        assert {tuple(p[1].counts) for p in pencils} == {(512, 250, 384), (512, 262, 384)}
        assert {tuple(p[2].counts) for p in pencils} == {(1024, 384, 128)}
    if metadata_only:
        Config()._commit()
        return
    G = P.global_array(dims, np.float64, kind="index")
    bricks1 = [L.Pencil(1, starts[r], counts[r]) for r in range(n)]
    xp = [p[0] for p in pencils]
    lastp = [p[nd - 1] for p in pencils]
    for rtype, src_l, dst_l in ((Reshape.X_BRICKS_TO_PENCILS, bricks1, xp), (Reshape.X_PENCILS_TO_BRICKS, xp, bricks1),
                                (Reshape.Z_PENCILS_TO_BRICKS, lastp, zb), (Reshape.Z_BRICKS_TO_PENCILS, zb, lastp)):
        src = P.redistribute(G, src_l)
        want = P.redistribute(G, dst_l)
        got = replay_fused(plans, rtype, src, [w.size for w in want], np.float64)
        for r in range(n):
            assert np.array_equal(got[r], want[r]), (rtype, r)
    Config()._commit()


# ---------------------------------------------------------------------------------------------
# R2C sizes, BASELINE configs, validation
# ---------------------------------------------------------------------------------------------
def test_r2c_sizes_config3():
    """BASELINE config 3: 1024^3 R2C fp32, 8 ranks (1x4x2): complex side 513x1024x1024
    (dtfft_plan.F90:1350-1355, 1868-1876, 2626-2630)."""
    cfg = Config(enable_z_slab=False)
    mk = lambda r, c: PlanR2C([1024, 1024, 1024], comm=c, precision=Precision.SINGLE, executor=Executor.CUFFT,
                              config=cfg, dry=True)
    # without a user grid the reference keeps the slab-shaped grid 1x1x8 even with the Z-slab
    # optimisation off (src/dtfft_transpose_plan.F90:193-195); the pencil grid needs a cart comm
    assert dry_world(8, mk)[0].grid_dims == [1, 1, 8]
    plans = dry_world(8, mk, cart_dims=[1, 4, 2])
    p0 = plans[0]
    assert p0.grid_dims == [1, 4, 2] and p0.element_size == 4
    ins, inc, outs, outc, alloc = p0.local_sizes
    assert inc == [1024, 256, 512]
    assert p0.get_pencil(Layout.X_PENCILS_FOURIER).counts == [513, 256, 512]
    ycounts = sorted({tuple(p.get_pencil(Layout.Y_PENCILS).counts) for p in plans})
    assert ycounts == [(1024, 512, 128), (1024, 512, 129)]
    zcounts = sorted({tuple(p.get_pencil(Layout.Z_PENCILS).counts) for p in plans})
    assert zcounts == [(1024, 128, 512), (1024, 129, 512)]
    # alloc in REAL elements = max(real volume, 2 * max complex pencil)
    for p in plans:
        cmax = max(int(np.prod(p.get_pencil(l).counts)) for l in (Layout.X_PENCILS_FOURIER, Layout.Y_PENCILS, Layout.Z_PENCILS))
        assert p.alloc_size == max(1024 * 256 * 512, 2 * cmax)
    Config()._commit()


def test_slab_config4_and_cart_grid():
    """BASELINE config 4: 16384^2 on 8 ranks = X slabs 16384 x 2048, per-peer block 2048 x 2048;
    a user process grid (the MPI_Cart_create case) is honoured."""
    plans = dry_world(8, lambda r, c: PlanC2C([16384, 16384], comm=c, dry=True))
    assert plans[0].grid_dims == [1, 8]
    d = plans[3].describe_exchange(Transpose.X_TO_Y)
    assert d["send_counts"].tolist() == [2048 * 2048] * 8
    plans = dry_world(8, lambda r, c: PlanC2C([64, 48, 40], comm=c, dry=True), cart_dims=[1, 2, 4])
    assert plans[0].grid_dims == [1, 2, 4] and not plans[0].z_slab_enabled
    plans = dry_world(4, lambda r, c: PlanC2C([64, 48, 40], comm=c, dry=True), cart_dims=[4])
    assert plans[0].grid_dims == [1, 1, 4] and plans[0].z_slab_enabled
    with pytest.raises(DtfftError) as e:
        dry_world(4, lambda r, c: PlanC2C([64, 48, 40], comm=c, dry=True), cart_dims=[2, 2, 1])
    assert e.value.code == 10  # DTFFT_ERROR_INVALID_COMM_FAST_DIM


def test_create_argument_validation():
    """Error codes of check_create_args (src/dtfft_plan.F90:2107-2219) and of the pencil checks
    (src/dtfft_pencil.F90:777-880; tests/fortran/test_pencils_f.F90:62-222)."""
    def code(fn):
        with pytest.raises(DtfftError) as e:
            fn()
        return e.value.code

    assert code(lambda: PlanC2C([8], dry=True)) == 3
    assert code(lambda: PlanC2C([8, 8, 8, 8], dry=True)) == 3
    assert code(lambda: PlanC2C([8, 0, 8], dry=True)) == 4
    assert code(lambda: PlanC2C([8, 8, 8], precision=7, dry=True)) == 6
    assert code(lambda: PlanC2C([8, 8, 8], executor=9, dry=True)) == 8
    assert code(lambda: PlanC2C([8, 8, 8], executor=Executor.FFTW3, dry=True)) == 401
    assert code(lambda: PlanR2C([8, 8, 8], dry=True)) == 13           # R2C transpose-only plan
    assert code(lambda: PlanR2R([8, 8, 8], executor=Executor.CUFFT, dry=True)) == 11  # missing kinds
    assert code(lambda: PlanC2C(Pencil([0, -1, 0], [4, 4, 4]), dry=True)) == 28
    assert code(lambda: PlanC2C(Pencil([0, 0, 0], [4, -4, 4]), dry=True)) == 27
    # overlapping / non-tiling user pencils on 2 ranks
    boxes = [([0, 0, 0], [8, 4, 8]), ([0, 3, 0], [8, 5, 8])]
    assert code(lambda: dry_world(2, lambda r, c: PlanC2C(Pencil(*boxes[r]), comm=c, dry=True))) == 30
    boxes = [([0, 0, 0], [8, 4, 8]), ([0, 5, 0], [8, 3, 8])]
    assert code(lambda: dry_world(2, lambda r, c: PlanC2C(Pencil(*boxes[r]), comm=c, dry=True))) == 31
    # a dry plan never executes
    p = PlanC2C([8, 8, 8], dry=True)
    assert code(lambda: p.mem_alloc(64)) == 203
    # config validation (src/dtfft_config.F90:677-764)
    assert code(lambda: Config(n_measure_iters=0)._commit()) == 34
    assert code(lambda: Config(platform=1)._commit()) == 400
    assert code(lambda: Config(backend=22)._commit()) == 402
    assert code(lambda: Config(backend=77)._commit()) == 202
    Config()._commit()


def _grid_candidates(dims, nranks):
    import ctypes as C

    from dtfft_b200 import _lib

    L_ = _lib.lib()
    L_.dtfftb_grid_candidates.restype = C.c_int32
    L_.dtfftb_grid_candidates.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    out = (C.c_int32 * 64)()
    n = L_.dtfftb_grid_candidates((C.c_int32 * 3)(*dims), nranks, 32, out)
    return [(out[2 * i], out[2 * i + 1]) for i in range(n)]


def test_grid_search_candidates():
    """Grids tried by the DTFFT_MEASURE / DTFFT_PATIENT search: every divisor pair in the reference's
    order (src/dtfft_transpose_plan.F90:456-500), minus those with fewer points than ranks along a
    split axis of any pencil (autotune_grid, :600-607)."""
    assert _grid_candidates((512, 512, 512), 8) == [(1, 8), (8, 1), (2, 4), (4, 2)]
    assert _grid_candidates((512, 512, 512), 4) == [(1, 4), (4, 1), (2, 2)]
    assert _grid_candidates((512, 512, 512), 6) == [(1, 6), (6, 1), (2, 3), (3, 2)]
    assert _grid_candidates((512, 512, 512), 1) == [(1, 1)]
    # y has 3 points: a grid dimension larger than 3 can never split it (X pencil: y over g1, Z pencil: y over g2)
    assert _grid_candidates((64, 3, 100), 8) == []
    assert _grid_candidates((64, 5, 100), 8) == [(2, 4), (4, 2)]
    # x has 2 points: it is split over g1 in the Y and Z pencils
    assert _grid_candidates((2, 64, 64), 8) == [(1, 8), (2, 4)]


@pytest.mark.parametrize("dims,nranks", [((128, 64, 96), 8), ((48, 21, 36), 6), ((40, 33, 28), 4)])
def test_grid_search_redecomposition_equals_cart_grid(dims, nranks):
    """Switching a default plan to the grid 1 x g1 x g2 (what the grid search does between two
    timings) gives, on every rank, exactly the pencils and exchange geometry of a plan CREATED on
    that process grid."""
    cfg = Config(enable_z_slab=False)
    for g1, g2 in _grid_candidates(dims, nranks):
        def switched(r, c):
            p = PlanC2C(list(dims), comm=c, config=cfg, dry=True)
            from dtfft_b200 import _lib
            from dtfft_b200.plan import _check

            _lib.lib().dtfftb_plan_dry_set_grid.restype = int
            _check(_lib.lib().dtfftb_plan_dry_set_grid(p._h, g1, g2), "dtfftb_plan_dry_set_grid")
            return p

        a = dry_world(nranks, switched)
        b = dry_world(nranks, lambda r, c: PlanC2C(list(dims), comm=c, config=cfg, dry=True), cart_dims=[1, g1, g2])
        for pa, pb in zip(a, b):
            assert pa.grid_dims == pb.grid_dims == [1, g1, g2]
            assert pa.local_sizes == pb.local_sizes and pa.alloc_bytes == pb.alloc_bytes
            for lay in LAYOUT_OF_PENCIL:
                ga, gb = pa.get_pencil(lay), pb.get_pencil(lay)
                assert (ga.starts, ga.counts) == (gb.starts, gb.counts)
            for t in (Transpose.X_TO_Y, Transpose.Y_TO_X, Transpose.Y_TO_Z, Transpose.Z_TO_Y):
                da, db = pa.describe_exchange(t), pb.describe_exchange(t)
                assert list(da["members"]) == list(db["members"])
                assert np.array_equal(da["fused_boxes"], db["fused_boxes"])
                for key in ("send_nd", "recv_nd", "send_counts", "send_displs", "recv_counts", "recv_displs"):
                    assert np.array_equal(da[key], db[key]), key
                assert (da["pack_kernel"], da["unpack_kernel"]) == (db["pack_kernel"], db["unpack_kernel"])


def test_python_api_surface_of_the_reference_module():
    """Names a user of the reference's Python package relies on (src/interfaces/python/__init__.py:
    25-72 build queries, 112-115 exception / version, 360-381 Request, 708-710 dtype) exist with the
    same meaning; values are those of this build (CUDA + NCCL + cuFFT, nothing else)."""
    import dtfft_b200 as d

    assert d.is_cuda_enabled() and d.is_cufft_enabled() and d.is_nccl_enabled()
    assert not (d.is_fftw_enabled() or d.is_mkl_enabled() or d.is_vkfft_enabled() or d.is_nvshmem_enabled()
                or d.is_compression_enabled() or d.is_transpose_only_enabled())
    assert d.get_backend_string(d.Backend.NCCL) == "NCCL"
    assert d.get_backend_string(d.Backend.NCCL_PIPELINED) == "NCCL_PIPELINED"
    assert d.Version.get() == d.Version.MAJOR * 100000 + d.Version.MINOR * 1000 + d.Version.PATCH
    assert d.dtfft_Exception is d.DtfftError
    assert (d.TransposeMode.PACK, d.TransposeMode.UNPACK, d.AccessMode.WRITE, d.AccessMode.READ) == (15, 16, -1, 1)
    assert PlanC2C([8, 8, 8], dry=True).dtype == np.complex128
    assert PlanC2C([8, 8, 8], precision=Precision.SINGLE, dry=True).dtype == np.complex64
    assert PlanR2R([8, 8, 8], dry=True).dtype == np.float64
    assert PlanR2C([8, 8, 8], precision=Precision.SINGLE, executor=Executor.CUFFT, dry=True).dtype == np.float32
    r = d.Request(0x1234, "Transpose.X_TO_Y")
    assert int(r) == r.handle == 0x1234 and r.kind == "Transpose.X_TO_Y" and "0x1234" in repr(r)
    # a request nobody started is rejected (CHECK_REQUEST, src/dtfft_plan.F90:75-84)
    p = PlanC2C([8, 8, 8], dry=True)
    for bad in (d.Request(0, "x"), d.Request(0xdead0, "x")):
        with pytest.raises(DtfftError) as e:
            p.transpose_end(bad)
        assert e.value.code == 35  # DTFFT_ERROR_INVALID_REQUEST
    with pytest.raises(ValueError):
        p.get_ndarray(8, shape=(3, 3))
    # shape / order handling of get_ndarray (the allocation itself needs a GPU: tests/test_zz_python_api_gpu.py)
    import torch

    from dtfft_b200.plan import _shaped_view

    flat = torch.arange(40, dtype=torch.float64)
    f = _shaped_view(flat, (2, 3, 4), "F")
    assert tuple(f.shape) == (2, 3, 4) and f.stride() == (1, 2, 6) and f[1, 2, 3] == 1 + 2 * 2 + 3 * 6
    c = _shaped_view(flat, (2, 3, 4), "C")
    assert c.stride() == (12, 4, 1) and c[1, 2, 3] == 23
    assert f.data_ptr() == flat.data_ptr() == c.data_ptr()


def test_report_follows_the_reference_format(capfd):
    """dtfft_report: the reference's lines and labels (src/dtfft_plan.F90:1556-1631), incl. the initial /
    final grids of a plan created from bricks (src/dtfft_reshape_plan.F90:155-158, 199-204)."""
    boxes = brick_boxes([[30, 34], [20, 12], [70, 58]])
    cfg = Config(enable_fourier_reshape=True, enable_z_slab=False)
    plans = dry_world(len(boxes), lambda r, c: PlanR2R(Pencil(*boxes[r]), comm=c, config=cfg, dry=True))
    capfd.readouterr()
    for p in plans:
        p.report()  # only rank 0 prints
    out = capfd.readouterr().out.splitlines()
    assert out[0] == "dtFFT: **Plan report**" and out[-1] == "dtFFT: **End of report**"
    assert sum(l == "dtFFT: **Plan report**" for l in out) == 1
    body = {l.split(":", 2)[1].strip(): l.split(":", 2)[2].strip() for l in out[1:-1]}
    assert body["dtFFT Version"] == "3.2.0" and body["Number of dimensions"] == "3"
    assert body["Global dimensions"] == "64x32x128" and body["Grid decomposition"] == "1x2x4"
    assert body["Initial grid"] == "2x2x2" and body["Final grid"] == "2x2x2" and body["Final reshape enabled"] == "True"
    assert body["Execution platform"] == "CUDA" and body["Plan type"] == "Real-to-Real"
    assert body["Plan precision"] == "Double" and body["FFT Executor type"] == "None"
    assert body["Z-slab enabled"] == "False" and body["Y-slab enabled"] == "False"
    assert body["Backend"] == "NCCL" and body["Reshape Backend"] == "NCCL"
    Config()._commit()


@pytest.mark.parametrize("dims,nranks", [((128, 96, 40), 2), ((256, 128, 10), 4), ((64, 64, 64), 2), ((80, 70, 33), 2)])
def test_y_slab_decomposition(dims, nranks):
    """Y-slab optimisation (enable_y_slab, src/dtfft_transpose_plan.F90:182-191): when both x and y are long
    enough the grid is 1 x P x 1, the plan ends in Y pencils (execute = X -> Y only, dtfft_plan.F90:856-862,
    get_local_sizes :1868-1876 with is_y_slab) and in-place transpose-only execution is refused (:800-804)."""
    cfg = Config(enable_z_slab=False, enable_y_slab=True)
    plans = dry_world(nranks, lambda r, c: PlanC2C(list(dims), comm=c, config=cfg, dry=True))
    comm_dims, is_z, is_y = L.choose_grid(list(dims), nranks, cuda=True, z_slab=False, y_slab=True)
    G = P.global_array(dims, np.complex128, kind="index")
    for r, plan in enumerate(plans):
        assert plan.grid_dims == comm_dims and plan.y_slab_enabled == is_y and not plan.z_slab_enabled
        gold = L.make_pencils(list(dims), comm_dims, r)
        ins, inc, outs, outc, alloc = plan.local_sizes
        last = gold[1] if is_y else gold[2]
        assert (ins, inc) == (gold[0].starts, gold[0].counts) and (outs, outc) == (last.starts, last.counts)
        assert alloc == max(p.size for p in gold)
    if is_y:
        assert comm_dims == [1, nranks, 1]
    for t in (1, -1, 2, -2):
        src = P.scatter_input(G, list(dims), comm_dims, t)
        want = P.transpose_datatype(G, list(dims), comm_dims, t)
        got = replay_fused(plans, t, src, [w.size for w in want], np.complex128)
        for r in range(nranks):
            assert np.array_equal(got[r], want[r]), (t, r)
    Config()._commit()


# ---------------------------------------------------------------------------------------------
# copy-engine form of the direct-store exchange (geometry.h: DmaBlock) and its peer-by-peer pipelines
# ---------------------------------------------------------------------------------------------
def _dma_copy(staging, dst, cp, pack_off):
    """What cudaMemcpy3DAsync does with the parameters handle.cu: dma_send builds from a DmaBlock."""
    run, rows, planes = cp["run"], cp["rows"], cp["planes"]
    for pl in range(planes):
        for rw in range(rows):
            s0 = pack_off + (pl * rows + rw) * run
            d0 = cp["dst_off"] + rw * cp["dst_pitch"] + pl * cp["dst_plane_rows"] * cp["dst_pitch"]
            dst[d0: d0 + run] = staging[s0: s0 + run]


@pytest.fixture(params=[None, "2048"], ids=["whole-blocks", "sliced"])
def sub_bytes(request):
    """DTFFTB_DMA_SUB_BYTES: None = default (16 MiB: test-sized blocks travel whole), 2048 = blocks cut into slices."""
    old = os.environ.get("DTFFTB_DMA_SUB_BYTES")
    if request.param is None:
        os.environ.pop("DTFFTB_DMA_SUB_BYTES", None)
    else:
        os.environ["DTFFTB_DMA_SUB_BYTES"] = request.param
    yield request.param
    if old is None:
        os.environ.pop("DTFFTB_DMA_SUB_BYTES", None)
    else:
        os.environ["DTFFTB_DMA_SUB_BYTES"] = old


@pytest.mark.parametrize("dims,grid", [((24, 20, 36), [1, 1, 4]), ((17, 13, 29), [1, 1, 3]), ((16, 12, 10), [1, 2, 2]),
                                       ((33, 9, 14), [1, 3, 2]), ((40, 36), [1, 4]), ((21, 10), [1, 3]),
                                       ((96, 80, 72), [1, 1, 2])])
def test_dma_blocks_deliver_every_transposition(dims, grid, sub_bytes):
    """Every transposition as pack -> one strided 3-D copy per (peer, slice) (+ the self block stored directly): the
    members' destinations must equal the datatype truth, every element written exactly once, and the staging slices
    must tile the workspace the handle asks for."""
    nranks = int(np.prod(grid))
    cfg = Config(enable_z_slab=False)
    plans = dry_world(nranks, lambda r, c: PlanC2C(list(dims), comm=c, config=cfg, dry=True), cart_dims=grid)
    comm_dims = plans[0].grid_dims
    G = P.global_array(dims, np.complex128, kind="index")
    SENT = np.complex128(-7 - 7j)
    sliced = 0
    for t in ([1, -1] if len(dims) == 2 else [1, -1, 2, -2]):
        src = P.scatter_input(G, list(dims), comm_dims, t)
        want = P.transpose_datatype(G, list(dims), comm_dims, t)
        out = [np.full(w.size, SENT) for w in want]
        hits = [np.zeros(w.size, np.int32) for w in want]
        for r, plan in enumerate(plans):
            d = plan.describe_dma(t)
            members, me = d["members"], d["me"]
            assert members[me] == r
            remote = sum(int(np.prod(e["fused"][:3])) for e in d["entries"] if e["member"] != me and e["fused"][0] > 0)
            staging = np.full(max(remote, 1), SENT)
            off = 0
            for e in d["entries"]:
                f, peer = e["fused"], members[e["member"]]
                sliced += e["nsub"] > 1
                if f[0] <= 0:
                    continue
                if e["member"] == me:
                    P.apply_boxes(src[r], out, [f], [peer])
                    P.apply_boxes(np.ones(src[r].size, np.int32), hits, [f], [peer])
                    continue
                cp = e["copy"]
                assert cp["ok"] == 1, (t, r, e)
                assert e["pack"][4] == off  # slices are packed back to back
                vol = int(np.prod(f[:3]))
                assert cp["run"] * cp["rows"] * cp["planes"] == vol
                P.apply_boxes(src[r], [staging], [e["pack"]], [0])
                assert not np.any(staging[off: off + vol] == SENT)
                _dma_copy(staging, out[peer], cp, off)
                _dma_copy(np.ones(staging.size, np.int32), hits[peer], cp, off)
                off += vol
            assert off == remote
        for r in range(nranks):
            assert np.array_equal(out[r], want[r]), (t, r)
            assert np.all(hits[r] == 1), (t, r)
    if sub_bytes and tuple(dims) == (96, 80, 72):
        assert sliced > 0  # the big case really was cut
    Config()._commit()


@pytest.mark.parametrize("dims,nranks", [((24, 20, 36), 4), ((17, 13, 29), 3), ((32, 8, 16), 8), ((12, 10, 7), 2),
                                         ((96, 80, 72), 2), ((67, 5, 9), 4), ((70, 33, 11), 5), ((33, 64, 6), 6),
                                         ((130, 7, 40), 3)])
def test_peer_by_peer_pair_pipelines(dims, nranks, sub_bytes):
    """Plan::run_transpose_pair with copy-engine exchanges on a slab-shaped grid 1 x 1 x P.  Forward: the local X->Y
    runs in pieces cut by the (peer, slice) blocks of the Y->Z exchange; after a piece the pack of that slice must
    find exactly its source elements written (the rest of the Y pencil may still hold the sentinel).  Backward: the
    local Y->X runs in pieces cut by the senders' slices; a piece may only read what that slice delivered.  All
    pieces together = the whole transposition, every element exactly once."""
    cfg = Config(enable_z_slab=False)
    plans = dry_world(nranks, lambda r, c: PlanC2C(list(dims), comm=c, config=cfg, dry=True), cart_dims=[1, 1, nranks])
    comm_dims = plans[0].grid_dims
    G = P.global_array(dims, np.complex128, kind="index")
    SENT = np.complex128(-7 - 7j)
    wantY = P.transpose_datatype(G, list(dims), comm_dims, 1)
    # ---- forward: X -> Y local pieces feeding the packs of Y -> Z -----------------------------------------
    X = P.scatter_input(G, list(dims), comm_dims, 1)
    wantZ = P.transpose_datatype(G, list(dims), comm_dims, 2)
    outZ = [np.full(w.size, SENT) for w in wantZ]
    for r, plan in enumerate(plans):
        d = plan.describe_dma(2)
        members, me = d["members"], d["me"]
        by_member = {}
        for e in d["entries"]:
            by_member.setdefault(e["member"], []).append(e)
        mid = np.full(wantY[r].size, SENT)
        hits = np.zeros(wantY[r].size, np.int32)
        staging = np.full(max(1, sum(int(np.prod(e["fused"][:3])) for e in d["entries"] if e["member"] != me and e["fused"][0] > 0)), SENT)
        order = [(me + k) % len(members) for k in range(1, len(members))] + [me]
        for p in order:
            for e in by_member[p]:
                piece, nsub = plan.describe_peer_piece(1, 2, 0, p, e["sub"])
                assert nsub == e["nsub"]
                P.apply_boxes(X[r], [mid], piece, [0])
                P.apply_boxes(np.ones(X[r].size, np.int32), [hits], piece, [0])
                f = e["fused"]
                if f[0] <= 0:
                    continue
                if p == me:
                    P.apply_boxes(mid, outZ, [f], [members[p]])
                else:
                    P.apply_boxes(mid, [staging], [e["pack"]], [0])
                    vol, off = int(np.prod(f[:3])), int(e["pack"][4])
                    assert not np.any(staging[off: off + vol] == SENT), (r, p, e["sub"])  # the piece produced all the pack reads
                    _dma_copy(staging, outZ[members[p]], e["copy"], off)
        assert np.array_equal(mid, wantY[r]) and np.all(hits == 1), r
    for r in range(nranks):
        assert np.array_equal(outZ[r], wantZ[r]), r
    # ---- backward: slices of Z -> Y land one at a time, the local Y -> X consumes them ---------------------
    Z = P.scatter_input(G, list(dims), comm_dims, -2)
    wantX = P.transpose_datatype(G, list(dims), comm_dims, -1)
    mids = [np.full(w.size, SENT) for w in wantY]
    outX = [np.full(w.size, SENT) for w in wantX]
    hitsX = [np.zeros(w.size, np.int32) for w in wantX]
    descr = [plan.describe_dma(-2) for plan in plans]
    P_ = nranks
    for r, plan in enumerate(plans):  # self blocks first (handle.cu: dma_self precedes the waits)
        d = descr[r]
        me = d["me"]
        for e in d["entries"]:
            if e["member"] == me and e["fused"][0] > 0:
                P.apply_boxes(Z[r], mids, [e["fused"]], [d["members"][me]])
        piece, _ = plan.describe_peer_piece(-1, -2, 1, me, 0)
        P.apply_boxes(mids[r], [outX[r]], piece, [0])
        P.apply_boxes(np.ones(mids[r].size, np.int32), [hitsX[r]], piece, [0])
    for k in range(1, P_):  # step k: every sender's block for its k-th target lands slice by slice
        for r in range(nranks):
            d = descr[r]
            me = d["me"]
            p = (me + k) % P_
            ents = [e for e in d["entries"] if e["member"] == p]
            for e in ents:
                f = e["fused"]
                if f[0] > 0:
                    vol = int(np.prod(f[:3]))
                    staging = np.full(int(e["pack"][4]) + vol, SENT)
                    P.apply_boxes(Z[r], [staging], [e["pack"]], [0])
                    _dma_copy(staging, mids[d["members"][p]], e["copy"], int(e["pack"][4]))
                # ... and the receiver consumes exactly that slice right away (everything later is still the sentinel)
                recv = d["members"][p]
                rplan, rme = plans[recv], descr[recv]["me"]
                src_member = (rme - k + P_) % P_
                assert descr[recv]["members"][src_member] == r
                piece, nsub = rplan.describe_peer_piece(-1, -2, 1, src_member, e["sub"])
                assert nsub == e["nsub"]
                P.apply_boxes(mids[recv], [outX[recv]], piece, [0])
                P.apply_boxes(np.ones(mids[recv].size, np.int32), [hitsX[recv]], piece, [0])
    for r in range(nranks):
        assert np.array_equal(outX[r], wantX[r]), r  # a piece that read ahead of its slice would have copied the sentinel
        assert np.all(hitsX[r] == 1), r
    Config()._commit()


def _random_dma_cases(n_cases, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_cases):
        nd_ = int(rng.choice([2, 3, 3]))
        if nd_ == 2:
            p = int(rng.choice([2, 3, 4, 5]))
            grid = [1, p]
            dims = [int(rng.integers(p, 60)), int(rng.integers(p, 60))]
        else:
            g1, g2 = int(rng.choice([1, 2, 3])), int(rng.choice([1, 2, 3, 4]))
            if g1 * g2 == 1:
                g2 = 2
            grid = [1, g1, g2]
            dims = [int(rng.integers(max(g1, 2), 40)), int(rng.integers(max(g1, g2, 2), 40)), int(rng.integers(max(g2, 2), 40))]
        out.append((tuple(dims), grid, str(int(rng.choice([64, 512, 4096])))))
    return out


@pytest.mark.parametrize("dims,grid,sub", _random_dma_cases(24, 2026))
def test_dma_blocks_fuzz(dims, grid, sub):
    """Random extents (primes, extents barely above the grid, uneven splits) and grids: whenever every block of a
    transposition is expressible as one strided 3-D copy (DmaBlock::ok -- the handle falls back to the direct-store
    kernel otherwise, for the whole group alike), packs + copies must deliver exactly the datatype truth."""
    old = os.environ.get("DTFFTB_DMA_SUB_BYTES")
    os.environ["DTFFTB_DMA_SUB_BYTES"] = sub
    try:
        nranks = int(np.prod(grid))
        cfg = Config(enable_z_slab=False)
        plans = dry_world(nranks, lambda r, c: PlanC2C(list(dims), comm=c, config=cfg, dry=True), cart_dims=grid)
        comm_dims = plans[0].grid_dims
        G = P.global_array(dims, np.complex128, kind="index")
        SENT = np.complex128(-7 - 7j)
        replayed = 0
        for t in ([1, -1] if len(dims) == 2 else [1, -1, 2, -2]):
            descr = [plan.describe_dma(t) for plan in plans]
            if len(descr[0]["members"]) == 1:
                continue
            if not all(e["copy"]["ok"] == 1 for d in descr for e in d["entries"]):
                continue  # not one pitched copy per block: this transposition keeps the direct-store kernel
            src = P.scatter_input(G, list(dims), comm_dims, t)
            want = P.transpose_datatype(G, list(dims), comm_dims, t)
            out = [np.full(w.size, SENT) for w in want]
            hits = [np.zeros(w.size, np.int32) for w in want]
            for r, d in enumerate(descr):
                members, me = d["members"], d["me"]
                remote = sum(int(np.prod(e["fused"][:3])) for e in d["entries"] if e["member"] != me and e["fused"][0] > 0)
                staging = np.full(max(remote, 1), SENT)
                for e in d["entries"]:
                    f, peer = e["fused"], members[e["member"]]
                    if f[0] <= 0:
                        continue
                    if e["member"] == me:
                        P.apply_boxes(src[r], out, [f], [peer])
                        P.apply_boxes(np.ones(src[r].size, np.int32), hits, [f], [peer])
                        continue
                    off = int(e["pack"][4])
                    P.apply_boxes(src[r], [staging], [e["pack"]], [0])
                    _dma_copy(staging, out[peer], e["copy"], off)
                    _dma_copy(np.ones(staging.size, np.int32), hits[peer], e["copy"], off)
            for r in range(nranks):
                assert np.array_equal(out[r], want[r]), (t, r)
                assert np.all(hits[r] == 1), (t, r)
            replayed += 1
        assert replayed >= 1 or nranks == 1
    finally:
        if old is None:
            os.environ.pop("DTFFTB_DMA_SUB_BYTES", None)
        else:
            os.environ["DTFFTB_DMA_SUB_BYTES"] = old
        Config()._commit()
