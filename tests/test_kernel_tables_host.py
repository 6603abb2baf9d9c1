"""CPU check of what the GPU kernels are TOLD to do: the work-item tables built by
Kernel::rebuild_tables (dtfft_b200/csrc/kernel_object.cu) for dry kernels are run through a host
emulation of transpose_tiles_kernel / rows_copy_kernel (tests/kernel_emu.py) and must move exactly
what the oracle's kernels move (oracle/kernels.py, restating src/include/_dtfft_kernel_host_routines.inc
and src/dtfft_nvrtc_module.F90:494-578) -- every destination element written exactly once, nothing
else touched.  Covers every kernel kind the plan layer launches, the three element sizes, all access
widths of the row-copy family, per-peer and all-peer tables, the fused NVLink tables with their peer
interleaving, and the brick-reshape tables.  The device code itself is covered by the -m gpu tests."""
import numpy as np
import pytest

from dtfft_b200.kernel import Kernel
from dtfft_b200.plan import Config, Pencil, PlanC2C, PlanR2R, Reshape
from oracle import kernels as K
from oracle import layout as L
from oracle import pipeline as P
from tests import kernel_emu as E
from tests.test_plan_host import brick_boxes, dry_world

DT = {4: np.float32, 8: np.float64, 16: np.complex128}
FAMILY = {"transpose": "T", "rows": "R"}


def _rand(n, es, seed=5):
    rng = np.random.default_rng(seed)
    a = rng.random(n) + (1j * rng.random(n) if es == 16 else 0)
    return a.astype(DT[es])


def _emulate(kern, es, src, out_size, neighbor, unit=None, grid=None, fill=-5):
    """Run the dumped table of `kern` on `src`; returns (out, write counts) as element arrays."""
    fam = FAMILY[kern.info()["family"]]
    u = es if fam == "T" else unit
    table = kern.dump_table(unit=u, neighbor=neighbor)
    out = np.full(out_size, fill, DT[es])
    s2 = src.view(np.uint8).reshape(-1, u)
    o2 = out.view(np.uint8).reshape(-1, u)
    cnt = np.zeros(o2.shape[0], np.int32)
    E.run_table(table, fam, s2, {-1: o2}, {-1: cnt}, grid=grid)
    return out, np.repeat(cnt, u).reshape(-1, es), table  # write count of every byte, grouped per element


def _check(kern, es, kt, dims, src, out_size, nd, neighbor, grid=None):
    gold = np.full(out_size, -5, DT[es])
    K.execute(kt, dims, src, gold, nd, neighbor if neighbor else None)
    touched = np.full(out_size, False)
    probe = np.zeros(out_size, np.int8)
    K.execute(kt, dims, np.ones(src.size, np.int8), probe, nd, neighbor if neighbor else None)
    touched = probe == 1
    fam = FAMILY[kern.info()["family"]]
    units = [es] if fam == "T" else [u for u in (4, 8, 16) if u <= es or True]
    ran = 0
    for u in units:
        if fam == "R" and kern.dump_table(unit=u, neighbor=neighbor)["total_items"] == 0 and touched.any():
            continue  # this access width does not divide the geometry
        out, cnt, _ = _emulate(kern, es, src, out_size, neighbor, unit=u, grid=grid)
        assert np.array_equal(out.view(np.uint8), gold.view(np.uint8)), (kt, dims, es, u, neighbor)
        assert np.array_equal(cnt.min(axis=1) == 1, touched) and cnt.max() <= 1, (kt, dims, es, u)
        ran += 1
    assert ran >= 1 or not touched.any()


PERMUTE_DIMS = [[33, 77, 21], [18, 33, 155], [90, 57], [18, 155], [64, 64, 64], [1, 40, 3], [65, 1, 33], [129, 99, 33]]


@pytest.mark.parametrize("dims", PERMUTE_DIMS)
@pytest.mark.parametrize("es", [4, 8, 16])
def test_whole_pencil_permutes(dims, es):
    """forward / backward / backward_start on the reference's own test shapes
    (src/tests/test_host_kernels.F90:8-22, src/tests/test_device_kernels.F90:27-40) and odd ones."""
    n = int(np.prod(dims))
    src = _rand(n, es)
    kinds = [K.KERNEL_PERMUTE_FORWARD, K.KERNEL_PERMUTE_BACKWARD]
    if len(dims) == 3:
        kinds.append(K.KERNEL_PERMUTE_BACKWARD_START)
    for kt in kinds:
        k = Kernel().create_dry(dims, es, kt)
        _check(k, es, kt, dims, src, n, None, 0, grid=None if n < 50000 else 211)
        k.destroy()


PLAN_CASES = [((48, 21, 36), (1, 3, 2)), ((13, 7, 9), (1, 2, 3)), ((40, 33, 28), (1, 1, 4)), ((37, 22), (1, 3)),
              ((66, 10, 12), (1, 2, 2)), ((33, 34, 35), (1, 4, 1))]


@pytest.mark.parametrize("gdims,grid", PLAN_CASES)
@pytest.mark.parametrize("es", [4, 8, 16])
@pytest.mark.parametrize("mode", ["plain", "pipelined", "fused"])
def test_kernels_of_every_transposition(gdims, grid, es, mode):
    """pack-side and unpack-side kernels of every transposition of a plan, with the reference's
    neighbor_data (uneven splits): all-peer launches for the looped kinds, one launch per peer for the
    pipelined / fused kinds (src/dtfft_kernel_device.F90:141-174)."""
    nd_ = len(gdims)
    ttypes = [1, -1] if nd_ == 2 else [1, -1, 2, -2]
    if nd_ == 3 and grid[1] == 1:
        ttypes += [3, -3]
    for t in ttypes:
        pencils, geos = L.plan_geometry(list(gdims), list(grid), t, pipelined=mode == "pipelined", fused=mode == "fused")
        for r in (0, len(geos) - 1):
            g = geos[r]
            if g.comm_size == 1:
                continue
            alloc = max(p.size for p in pencils[r])
            src = _rand(alloc, es, seed=r + 3)
            for kt, kdims, nd in ((g.pack_kernel, g.send_dims, g.send_nd), (g.unpack_kernel, g.recv_dims, g.recv_nd)):
                per_peer = kt in K.PER_NEIGHBOR_KERNELS
                k = Kernel().create_dry(kdims, es, kt, nd if kt in K.PER_NEIGHBOR_KERNELS or kt in K.LOOPED else None)
                for nb in (range(1, g.comm_size + 1) if per_peer else [0]):
                    _check(k, es, kt, kdims, src, alloc, nd, nb)
                k.destroy()


shuffled = [0]


def _fused_case(plans, t, src, want, es, family_T, grid):
    n = len(plans)
    dsts = [np.full(w.size, -7, DT[es]) for w in want]
    cnts = [np.zeros(w.size, np.int32) for w in want]
    for r, plan in enumerate(plans):
        d = plan.describe_exchange(t)
        assert d["fused_transposing"] == family_T or not np.any(d["fused_boxes"][:, 0] > 0)
        k = Kernel().create_boxes_dry(2 if family_T else 3, es, d["fused_boxes"], remote_peers=True)
        if k.info()["family"] == "none":
            continue
        units = [es] if family_T else [u for u in (16, 8, 4) if k.dump_table(unit=u)["total_items"] > 0][:1]
        u = units[0]
        table = k.dump_table(unit=u)
        sh = int(table["blocks"][0][E.SHUFFLE])
        if sh > 1:  # peers are interleaved by a multiplier coprime with the item count
            assert np.gcd(sh, table["total_items"]) == 1
            shuffled[0] += 1
        s2 = src[r].view(np.uint8).reshape(-1, u)
        views = {i: dsts[m].view(np.uint8).reshape(-1, u) for i, m in enumerate(d["members"])}
        cviews = {i: np.zeros(views[i].shape[0], np.int32) for i in views}
        E.run_table(table, "T" if family_T else "R", s2, views, cviews, grid=grid)
        for i, m in enumerate(d["members"]):
            cnts[m] += np.repeat(cviews[i], u).reshape(-1, es).max(axis=1)
        k.destroy()
    for r in range(n):
        assert np.array_equal(dsts[r].view(np.uint8), want[r].view(np.uint8)), (t, r)
        assert np.all(cnts[r] == 1), (t, r)


@pytest.mark.parametrize("dims,nranks,cart", [((48, 21, 36), 6, [1, 3, 2]), ((40, 33, 28), 4, None), ((37, 22), 3, None),
                                              ((64, 20, 18), 8, [1, 2, 4])])
@pytest.mark.parametrize("es", [8, 16])
def test_fused_nvlink_tables(dims, nranks, cart, es):
    """The one-kernel NVLink path: each rank's table (one box per peer, destination = the peer's array,
    peers interleaved across CTAs) stores every element of every destination pencil exactly once."""
    plans = dry_world(nranks, lambda r, c: PlanC2C(list(dims), comm=c, dry=True,
                                                   precision=1 if es == 16 else 0), cart_dims=cart)
    comm_dims = plans[0].grid_dims
    nd_ = len(dims)
    G = P.global_array(dims, DT[es] if es == 16 else np.complex64, kind="random")
    ttypes = [1, -1] if nd_ == 2 else [1, -1, 2, -2] + ([3, -3] if plans[0].z_slab_enabled else [])
    for t in ttypes:
        src = P.scatter_input(G, list(dims), comm_dims, t)
        want = P.transpose_datatype(G, list(dims), comm_dims, t)
        if es == 8:
            src = [s.view(np.float64) for s in src]
            want = [w.view(np.float64) for w in want]
        _fused_case(plans, t, src, want, es, True, grid=97)
    assert shuffled[0] > 0


@pytest.mark.parametrize("cuts", [[[30, 34], [20, 12], [70, 58]], [[10, 6], [40, 24], [5, 7]], [[20, 13, 7], [16, 17]]])
def test_brick_reshape_tables(cuts):
    """Row-copy tables of the brick <-> pencil reshapes: fused (peer destinations) and the NCCL
    pack / unpack kernels, all access widths the geometry allows."""
    boxes = brick_boxes(cuts)
    n = len(boxes)
    cfg = Config(enable_fourier_reshape=True, enable_z_slab=False)
    plans = dry_world(n, lambda r, c: PlanR2R(Pencil(*boxes[r]), comm=c, config=cfg, dry=True))
    # fused path: row-copy family with peer destinations
    starts, counts = [b[0] for b in boxes], [b[1] for b in boxes]
    nd_ = len(cuts)
    dims, comm_dims, coords, xs, xc, bgrid, _ = L.from_bricks(starts, counts)
    pencils = [L.pencils_from_x(dims, comm_dims, coords[r], xs[r], xc[r]) for r in range(n)]
    zb = L.z_bricks(dims, comm_dims, coords, [p[nd_ - 1] for p in pencils], bgrid)
    bricks1 = [L.Pencil(1, starts[r], counts[r]) for r in range(n)]
    xp, lastp = [p[0] for p in pencils], [p[nd_ - 1] for p in pencils]
    G = P.global_array(dims, np.float64, kind="index")
    for rtype, src_l, dst_l in ((Reshape.X_BRICKS_TO_PENCILS, bricks1, xp), (Reshape.X_PENCILS_TO_BRICKS, xp, bricks1),
                                (Reshape.Z_PENCILS_TO_BRICKS, lastp, zb), (Reshape.Z_BRICKS_TO_PENCILS, zb, lastp)):
        _fused_case(plans, rtype, P.redistribute(G, src_l), P.redistribute(G, dst_l), 8, False, grid=53)
    for rtype in (Reshape.X_BRICKS_TO_PENCILS, Reshape.X_PENCILS_TO_BRICKS, Reshape.Z_PENCILS_TO_BRICKS,
                  Reshape.Z_BRICKS_TO_PENCILS):
        for r in (0, n - 1):
            d = plans[r].describe_reshape(rtype)
            if len(d["members"]) == 1:
                continue
            alloc = plans[r].alloc_size
            src = _rand(alloc, 8, seed=r)
            for which in ("pack_boxes", "unpack_boxes"):
                gold = np.full(alloc, -5.0)
                P.apply_local_boxes(src, gold, d[which])
                k = Kernel().create_boxes_dry(3, 8, d[which])
                ran = 0
                for u in (4, 8, 16):
                    for nb in [0] + list(range(1, len(d["members"]) + 1)):
                        table = k.dump_table(unit=u, neighbor=nb)
                        if table["total_items"] == 0:
                            continue
                        if nb == 0:
                            out, cnt, _ = _emulate(k, 8, src, alloc, 0, unit=u)
                            assert np.array_equal(out, gold), (rtype, which, u)
                            assert cnt.max() == 1
                            ran += 1
                        else:
                            one = np.full(alloc, -5.0)
                            P.apply_local_boxes(src, one, d[which][nb - 1: nb])
                            out, _, _ = _emulate(k, 8, src, alloc, nb, unit=u)
                            assert np.array_equal(out, one), (rtype, which, u, nb)
                assert ran >= 1
                k.destroy()
    Config()._commit()


def _random_kernel_cases(n_cases, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_cases):
        nd_ = int(rng.choice([2, 3]))
        hi = 70 if nd_ == 3 else 300
        dims = [int(v) for v in rng.integers(1, hi, size=nd_)]
        out.append((dims, int(rng.choice([4, 8, 16])), int(rng.choice([0, 7, 64, 1000]))))
    return out


@pytest.mark.parametrize("dims,es,grid", _random_kernel_cases(30, 99))
def test_random_permute_tables(dims, es, grid):
    """Random shapes (extents of 1, primes, smaller than a tile) and grid sizes: one CTA per item or a
    grid-stride loop must give the same, exact result."""
    n = int(np.prod(dims))
    src = _rand(n, es, seed=n)
    kinds = [K.KERNEL_PERMUTE_FORWARD, K.KERNEL_PERMUTE_BACKWARD] + ([K.KERNEL_PERMUTE_BACKWARD_START] if len(dims) == 3 else [])
    for kt in kinds:
        k = Kernel().create_dry(dims, es, kt)
        _check(k, es, kt, dims, src, n, None, 0, grid=grid or None)
        k.destroy()


ALL_TILES = [(1, 1, 4), (1, 1, 8), (1, 1, 16), (2, 1, 8), (1, 2, 8), (2, 2, 8), (2, 2, 16), (1, 4, 16), (4, 1, 16)]


@pytest.mark.parametrize("tile", ALL_TILES)
@pytest.mark.parametrize("es", [4, 16])
def test_every_tile_configuration(tile, es):
    """Every compiled (KA, KB, ROWS) instantiation of the transpose family -- the candidates of the
    DTFFT_EXHAUSTIVE kernel autotune (src/dtfft_kernel_device.F90:338-397) -- gets a table that moves
    the right elements on shapes that do not divide the tile."""
    for dims in ([70, 45, 3], [33, 130, 2], [129, 31]):
        n = int(np.prod(dims))
        src = _rand(n, es, seed=7)
        for kt in (K.KERNEL_PERMUTE_FORWARD, K.KERNEL_PERMUTE_BACKWARD):
            k = Kernel().create_dry(dims, es, kt)
            k.set_tile(*tile)
            assert k.dump_table(unit=es)["launch"] == list(tile)
            _check(k, es, kt, dims, src, n, None, 0, grid=61)
            k.destroy()


@pytest.mark.parametrize("es,zcounts,ny,nx", [(8, [40, 56], 64, 24), (16, [36, 60], 32, 16), (4, [72, 88], 64, 8), (8, [250, 262], 32, 3)])
def test_tile_grid_is_aligned_to_destination_lines(es, zcounts, ny, nx):
    """Uneven splits make destination runs start off a 128-byte line (BASELINE config 5: 250 x 8 B).  The table builder
    then starts the tile grid BSHIFT elements before the box so that every tile boundary along the output-contiguous axis
    falls on a line of the destination (kernels.cu: b0 = t1 * TB - bshift, first tile masked at its start).  Both the
    shifted grid and the launch-time fallback for misaligned base pointers (`noshift`) must move exactly the box's
    elements, each once."""
    G = len(zcounts)
    nz = sum(zcounts)
    L = 128 // es
    nyl = ny // G
    rng = np.random.default_rng(3)
    for d in range(G):
        z0 = sum(zcounts[:d])
        boxes = [[nyl, zcounts[d], nx, p * nyl, z0, ny, ny * zcounts[d], nz * nx, 1, nz] for p in range(G)]
        k = Kernel().create_boxes_dry(2, es, boxes, remote_peers=True)
        table = k.dump_table(unit=es)
        tb = 32 * table["launch"][1]
        src = rng.integers(1, 2 ** 30, size=ny * zcounts[d] * nx).astype(np.int64)
        want = [np.zeros(nz * nx * nyl, np.int64) for _ in range(G)]
        P.apply_boxes(src, want, boxes, list(range(G)))
        shifted = 0
        for blk in table["blocks"]:
            assert blk[E.BSHIFT] == blk[E.OUT_OFF] % L  # rows of these boxes are all misaligned alike
            shifted += blk[E.BSHIFT] > 0
            # every tile boundary inside the run is a multiple of a line in destination elements
            for t1 in range(1, int(blk[E.TILES1])):
                assert (blk[E.OUT_OFF] + t1 * tb - blk[E.BSHIFT]) % L == 0
            assert blk[E.TILES1] == -(-(blk[E.N1] + blk[E.BSHIFT]) // tb)
        if z0 % L:
            assert shifted == len(table["blocks"])
        for noshift in (False, True):
            dsts = {p: np.zeros(nz * nx * nyl, np.int64) for p in range(G)}
            cnts = {p: np.zeros(nz * nx * nyl, np.int32) for p in range(G)}
            E.run_table(table, "T", src, dsts, cnts, grid=53, noshift=noshift)
            for p in range(G):
                assert np.array_equal(dsts[p], want[p]), (d, p, noshift)
                assert np.array_equal(cnts[p] == 1, want[p] != 0) and cnts[p].max() <= 1
        k.destroy()
