"""Brick <-> pencil reshapes over the NCCL backends on CPU: the product's pack / exchange / unpack
geometry (dtfftb_plan_describe_reshape, dry plans on simulated ranks) against the oracle's
restatement of reshape_handle_generic for reshapes (src/dtfft_reshape_handle_generic.F90:261-289,
343-404, 479-484, 536-611) and against the MPI-datatype truth (global-array slicing), including the
pack-free / unpack-free shortcuts and their buffer choreography (:695-759)."""
import numpy as np
import pytest

from dtfft_b200.plan import Config, Pencil, PlanR2R, Reshape
from oracle import kernels as K
from oracle import layout as L
from oracle import pipeline as P
from tests.test_plan_host import brick_boxes, dry_world

CASES = [
    # cuts per axis                                    expected (strat of X reshapes, pack-free)
    ([[30, 34], [20, 12], [70, 58]], 1, True),        # z long enough (> 32 * 2): z split
    ([[10, 6], [40, 24], [5, 7]], None, None),        # y / factorised split
    ([[10, 6, 8], [9, 11], [12, 8]], None, None),     # 3 bricks along x, short axes
    ([[6, 6, 6, 6], [70, 70], [70, 70]], None, None),  # 4 bricks along x: 2 x 2 factorisation candidate
    ([[12, 20], [16, 16], [140, 140]], 1, True),      # even z split
    ([[20, 13, 7], [16, 17]], 1, True),               # 2-D bricks: always pack-free (:262)
    ([[9, 9], [5, 5, 6]], 1, True),
]


def product_schedule(descs, inputs, out_sizes, buf_sizes, pipelined):
    """ReshapeHandle::execute (dtfft_b200/csrc/handle.cu) replayed with numpy: same kernels
    (boxes), same buffers, same order as the CUDA path."""
    n = len(descs)
    dtype = inputs[0].dtype
    a = [np.zeros(buf_sizes[r], dtype) for r in range(n)]
    b = [np.full(buf_sizes[r], -7, dtype) for r in range(n)]
    w = [np.full(buf_sizes[r], -9, dtype) for r in range(n)]
    for r in range(n):
        a[r][: inputs[r].size] = inputs[r]

    def exchange(src, dst):
        for r, d in enumerate(descs):
            for i, peer in enumerate(d["members"]):
                dp = descs[peer]
                j = dp["members"].index(r)
                cnt = int(d["recv_counts"][i])
                assert cnt == int(dp["send_counts"][j])
                so, ro = int(dp["send_displs"][j]), int(d["recv_displs"][i])
                dst[r][ro: ro + cnt] = src[peer][so: so + cnt]

    pack = lambda src, dst: [P.apply_local_boxes(src[r], dst[r], descs[r]["pack_boxes"]) for r in range(n)]
    unpack = lambda src, dst: [P.apply_local_boxes(src[r], dst[r], descs[r]["unpack_boxes"]) for r in range(n)]
    d0 = descs[0]
    if d0["is_pack_free"]:            # in -> aux exchange, aux -> out unpack (both flavours)
        exchange(a, w)
        unpack(w, b)
    elif d0["is_unpack_free"]:        # in -> aux pack, aux -> out exchange
        pack(a, w)
        exchange(w, b)
    elif pipelined:                   # in -> aux pack, aux -> in exchange, in -> out unpack
        pack(a, w)
        exchange(w, a)
        unpack(a, b)
    else:                             # in -> out pack, out -> in exchange, in -> out unpack
        pack(a, b)
        exchange(b, a)
        unpack(a, b)
    return [b[r][: out_sizes[r]] for r in range(n)]


@pytest.mark.parametrize("cuts,want_strat,want_free", CASES)
@pytest.mark.parametrize("pipelined", [False, True])
def test_reshape_nccl_geometry_and_shortcuts(cuts, want_strat, want_free, pipelined):
    check_brick_case(cuts, pipelined, want_strat, want_free)


def check_brick_case(cuts, pipelined, want_strat=None, want_free=None):
    boxes = brick_boxes(cuts)
    n, nd = len(boxes), len(cuts)
    cfg = Config(enable_fourier_reshape=True, enable_z_slab=False, backend=27 if pipelined else 24,
                 reshape_backend=27 if pipelined else 24)
    plans = dry_world(n, lambda r, c: PlanR2R(Pencil(*boxes[r]), comm=c, config=cfg, dry=True))
    starts, counts = [b[0] for b in boxes], [b[1] for b in boxes]
    dims, comm_dims, coords, xs, xc, bgrid, bcoords = L.from_bricks(starts, counts)
    pencils = [L.pencils_from_x(dims, comm_dims, coords[r], xs[r], xc[r]) for r in range(n)]
    zb = L.z_bricks(dims, comm_dims, coords, [p[nd - 1] for p in pencils], bgrid)
    bricks1 = [L.Pencil(1, starts[r], counts[r]) for r in range(n)]
    xp, lastp = [p[0] for p in pencils], [p[nd - 1] for p in pencils]
    G = P.global_array(dims, np.float64, kind="index")
    from dtfft_b200.plan import Layout
    for lay in [Layout.X_BRICKS, Layout.X_PENCILS, Layout.Y_PENCILS, Layout.Z_BRICKS] + ([Layout.Z_PENCILS] if nd == 3 else []):
        # every layout tiles the global grid (tests/python/test_pencil_api_py.py:76-86)
        assert sum(p.get_pencil(lay).size for p in plans) == int(np.prod(dims)), lay
    alloc = [plans[r].alloc_size for r in range(n)]
    want_aux = [0] * n

    for rtype, src_l, dst_l in ((Reshape.X_BRICKS_TO_PENCILS, bricks1, xp), (Reshape.X_PENCILS_TO_BRICKS, xp, bricks1),
                                (Reshape.Z_PENCILS_TO_BRICKS, lastp, zb), (Reshape.Z_BRICKS_TO_PENCILS, zb, lastp)):
        geos, descs = [], []
        for r in range(n):
            members = L.reshape_members(r, int(rtype), bgrid, bcoords, coords, xp)
            me = members.index(r)
            g = L.reshape_geometry(int(rtype), [src_l[m] for m in members], [dst_l[m] for m in members], me, members,
                                   pipelined=pipelined)
            d = plans[r].describe_reshape(rtype)
            geos.append(g)
            descs.append(d)
            # communicator, exchange tables, strategy and shortcut flags == reference formulas
            assert d["members"] == members and d["me"] == me, (rtype, r)
            if len(members) == 1:
                continue
            assert d["send_counts"].tolist() == g.send_counts and d["send_displs"].tolist() == g.send_displs
            assert d["recv_counts"].tolist() == g.recv_counts and d["recv_displs"].tolist() == g.recv_displs
            assert d["reshape_strat"] == g.reshape_strat, (rtype, r)
            assert (d["is_pack_free"], d["is_unpack_free"]) == (g.is_pack_free, g.is_unpack_free), (rtype, r)
            es = 8  # fp64 R2R: both sides of every reshape move 8-byte elements
            if pipelined:   # abstract_backend.F90:196-201
                want_aux[r] = max(want_aux[r], es * max(sum(g.send_counts), sum(g.recv_counts)))
            if g.is_pack_free or g.is_unpack_free:  # reshape_handle_generic.F90:684-686
                want_aux[r] = max(want_aux[r], es * max(src_l[r].size, dst_l[r].size))
        if len(geos[0].members) == 1:
            continue
        src = P.redistribute(G, src_l)
        want = P.redistribute(G, dst_l)
        out_sizes = [w_.size for w_ in want]
        # the product's pack / unpack boxes move exactly what the reference's pack / unpack kernels
        # move with the reference's neighbor_data (the wire format is the reference's)
        for r in range(n):
            g, d = geos[r], descs[r]
            a = np.zeros(alloc[r]); a[: src[r].size] = src[r]
            ref, got = np.full(alloc[r], -3.0), np.full(alloc[r], -3.0)
            K.execute(K.KERNEL_PACK, g.send_dims, a, ref, g.send_nd)
            P.apply_local_boxes(a, got, d["pack_boxes"])
            assert np.array_equal(ref, got), ("pack", rtype, r)
            if g.is_pack_free:  # the packed buffer IS the source array
                assert np.array_equal(ref[: src[r].size], src[r])
            slots = np.arange(alloc[r], dtype=np.float64) + 0.25
            ref, got = np.full(alloc[r], -3.0), np.full(alloc[r], -3.0)
            P.apply_local_boxes(slots, got, d["unpack_boxes"])
            if g.is_unpack_free:
                # the reference never runs an unpack kernel here (KERNEL_DUMMY, :623) and its
                # neighbor_data(5) of the z split (running sum of n1*n2, :594-597) is not a valid
                # scatter offset; what must hold is that the received slots ARE the brick
                assert np.array_equal(got[: out_sizes[r]], slots[: out_sizes[r]])
            else:
                K.execute(K.KERNEL_UNPACK, g.recv_dims, slots, ref, g.recv_nd)
                assert np.array_equal(ref, got), ("unpack", rtype, r)
        # reference schedule (with its shortcuts) and product schedule == datatype truth
        ref_out = P.reshape_generic(src, geos, out_sizes, alloc)
        got_out = product_schedule(descs, src, out_sizes, alloc, pipelined)
        for r in range(n):
            assert np.array_equal(ref_out[r], want[r]), ("oracle schedule", rtype, r)
            assert np.array_equal(got_out[r], want[r]), ("product schedule", rtype, r)
        if want_strat is not None and rtype == Reshape.X_BRICKS_TO_PENCILS:
            assert geos[0].reshape_strat == want_strat and geos[0].is_pack_free == want_free
        if want_strat is not None and rtype == Reshape.X_PENCILS_TO_BRICKS:
            assert geos[0].is_unpack_free == want_free
    for r in range(n):
        assert plans[r].aux_bytes_reshape == want_aux[r], r
    Config()._commit()
