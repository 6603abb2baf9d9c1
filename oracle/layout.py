"""Oracle (test infrastructure): decomposition and per-peer block geometry.

Restates, in plain Python over *all ranks at once* (no MPI -- every Allgather of the
reference becomes a loop over simulated ranks):

* ``get_local_size`` remainder rule            -- src/dtfft_pencil.F90:237-279
* pencil axis/communicator permutations         -- src/dtfft_transpose_plan.F90:1046-1082
* default grid choice (Z-slab / Y-slab / dims)  -- src/dtfft_transpose_plan.F90:170-203
* pencils of a cartesian grid                   -- src/dtfft_transpose_plan.F90:1084-1131
* which 1-D communicator a transposition uses   -- src/dtfft_abstract_reshape_handle.F90:156-188
* ``neighbor_data`` / counts / displs / kinds   -- src/dtfft_reshape_handle_generic.F90:291-640
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import kernels as K

# dtfft_transpose_t values (include/dtfft_config.h.in:38-43)
X_TO_Y, Y_TO_X, Y_TO_Z, Z_TO_Y, X_TO_Z, Z_TO_X = 1, -1, 2, -2, 3, -3
TRANSPOSE_NAMES = {1: "X_TO_Y", -1: "Y_TO_X", 2: "Y_TO_Z", -2: "Z_TO_Y", 3: "X_TO_Z", -3: "Z_TO_X"}
# dtfft_backend_t values used on this path (include/dtfft_config.h.in:153-169)
BACKEND_NCCL, BACKEND_NCCL_PIPELINED = 24, 27
DEF_TILE_SIZE = 32  # src/dtfft_parameters.F90 (Z-slab heuristic on CUDA)


def local_size(n_global: int, comm_dim: int, comm_rank: int):
    """(start, count) of rank ``comm_rank`` of ``comm_dim`` along an axis of ``n_global`` points.
    Remainder goes to the LAST ``n mod p`` ranks (src/dtfft_pencil.F90:255-267)."""
    if comm_dim == 1:
        return 0, n_global
    res = n_global % comm_dim
    base = n_global // comm_dim

    def cnt(r):
        return base + 1 if r >= comm_dim - res else base

    return sum(cnt(q) for q in range(comm_rank)), cnt(comm_rank)


def dims_create(nnodes: int, ndims: int, dims):
    """MPI_Dims_create for the shapes dtFFT asks for: fill the zero entries of ``dims``
    with factors as close to each other as possible, in non-increasing order."""
    dims = list(dims)
    free = [i for i, d in enumerate(dims) if d == 0]
    fixed = 1
    for d in dims:
        if d > 0:
            fixed *= d
    rem = nnodes // fixed
    if len(free) == 0:
        return dims
    if len(free) == 1:
        dims[free[0]] = rem
        return dims
    if len(free) == 2:
        b = int(np.floor(np.sqrt(rem)))
        while rem % b:
            b -= 1
        dims[free[0]], dims[free[1]] = rem // b, b
        return dims
    raise NotImplementedError("dims_create with >2 free dims is not needed by dtFFT")


def choose_grid(dims, comm_size: int, cuda: bool = True, z_slab: bool = True, y_slab: bool = False):
    """Default (non-cartesian communicator, DTFFT_ESTIMATE) grid choice.
    Returns ``(comm_dims, is_z_slab, is_y_slab)`` -- src/dtfft_transpose_plan.F90:170-203."""
    ndims = len(dims)
    comm_dims = [0] * ndims
    comm_dims[0] = 1
    cond1 = comm_size <= dims[-1]
    cond2 = comm_size <= dims[0] and comm_size <= dims[1]
    if cuda:
        cond1 = DEF_TILE_SIZE <= dims[-1] // comm_size
        cond2 = DEF_TILE_SIZE <= dims[0] // comm_size and DEF_TILE_SIZE <= dims[1] // comm_size
    is_z = is_y = False
    if ndims == 3:
        if cond1 and z_slab:
            comm_dims[1], comm_dims[2], is_z = 1, comm_size, True
        elif cond2 and y_slab:
            comm_dims[1], comm_dims[2], is_y = comm_size, 1, True
        elif cond1:
            comm_dims[1], comm_dims[2] = 1, comm_size
        elif cond2:
            comm_dims[1], comm_dims[2] = comm_size, 1
    comm_dims = dims_create(comm_size, ndims, comm_dims)
    return comm_dims, is_z, is_y


def cart_coords(rank: int, comm_dims):
    """Row-major cartesian coordinates (MPI_Cart_create, reorder treated as identity)."""
    coords = []
    for d in reversed(comm_dims):
        coords.append(rank % d)
        rank //= d
    return list(reversed(coords))


def cart_rank(coords, comm_dims):
    r = 0
    for c, d in zip(coords, comm_dims):
        r = r * d + c
    return r


def comm_members(rank: int, comm_dims, comm_id: int):
    """Global ranks of the 1-D communicator ``comm_id`` (1-based like ``helper%comms``)
    that contains ``rank``, in sub-communicator rank order.  ``comm_id == 1`` is the whole
    cartesian communicator (src/dtfft_abstract_backend.F90:408 ``comms(1) = base_comm``);
    ``comm_id == d`` (d >= 2) is MPI_Cart_sub keeping grid dimension ``d``."""
    n = int(np.prod(comm_dims))
    if comm_id == 1:
        return list(range(n))
    coords = cart_coords(rank, comm_dims)
    out = []
    for c in range(comm_dims[comm_id - 1]):
        cc = list(coords)
        cc[comm_id - 1] = c
        out.append(cart_rank(cc, comm_dims))
    return out


def permutations(ndims: int):
    """(dperm, cperm), 0-based, per pencil d: local axis j of pencil d is global axis
    dperm[d][j] and is split over grid dimension cperm[d][j]
    (src/dtfft_transpose_plan.F90:1046-1082)."""
    if ndims == 2:
        return [[0, 1], [1, 0]], [[0, 1], [0, 1]]
    return [[0, 1, 2], [1, 2, 0], [2, 0, 1]], [[0, 1, 2], [0, 2, 1], [0, 1, 2]]


@dataclass
class Pencil:
    aligned_dim: int              # 1-based like the reference
    starts: list
    counts: list

    @property
    def size(self):
        return int(np.prod(self.counts))


def make_pencils(dims, comm_dims, rank: int):
    """The ndims pencils of ``rank`` (src/dtfft_transpose_plan.F90:1106-1128, no user pencil)."""
    ndims = len(dims)
    dperm, cperm = permutations(ndims)
    coords = cart_coords(rank, comm_dims)
    pencils = []
    for d in range(ndims):
        starts, counts = [], []
        for j in range(ndims):
            g = cperm[d][j]
            s, c = local_size(dims[dperm[d][j]], comm_dims[g], coords[g])
            starts.append(s)
            counts.append(c)
        pencils.append(Pencil(d + 1, starts, counts))
    return pencils


def transpose_comm_id(ttype: int) -> int:
    """src/dtfft_abstract_reshape_handle.F90:157-164."""
    return {1: 2, 2: 3, 3: 1}[abs(ttype)]


def transpose_pencil_ids(ttype: int):
    """(send pencil index, recv pencil index), 0-based."""
    return {X_TO_Y: (0, 1), Y_TO_X: (1, 0), Y_TO_Z: (1, 2), Z_TO_Y: (2, 1), X_TO_Z: (0, 2), Z_TO_X: (2, 0)}[ttype]


@dataclass
class HandleGeometry:
    """Everything ``reshape_handle_generic%create`` derives for ONE rank."""
    ttype: int
    comm_size: int
    comm_rank: int
    members: list                       # global ranks of the 1-D communicator
    send_dims: list
    recv_dims: list
    pack_kernel: int
    unpack_kernel: int | None
    send_nd: np.ndarray = field(default=None)   # (P,5) neighbor_data handed to the pack kernel
    recv_nd: np.ndarray = field(default=None)   # (P,5) neighbor_data handed to the unpack kernel
    send_counts: list = field(default_factory=list)   # elements
    send_displs: list = field(default_factory=list)
    recv_counts: list = field(default_factory=list)
    recv_displs: list = field(default_factory=list)
    is_pipelined: bool = False
    is_fused: bool = False


def transpose_geometry(ttype: int, send_by_member, recv_by_member, me: int, members,
                       pipelined: bool = False, fused: bool = False) -> HandleGeometry:
    """Block geometry of a transposition for sub-communicator member ``me``.

    ``send_by_member[i]`` / ``recv_by_member[i]`` are the send / recv :class:`Pencil` of
    member ``i`` (what the reference obtains with two ``MPI_Allgather`` calls,
    src/dtfft_reshape_handle_generic.F90:143-144).  Index comments use the reference's
    1-based dimension numbers."""
    p = len(members)
    send, recv = send_by_member[me], recv_by_member[me]
    ndims = len(send.counts)
    forward = ttype in (X_TO_Y, Y_TO_Z, Z_TO_X)
    kernel_type = K.KERNEL_PERMUTE_FORWARD if forward else K.KERNEL_PERMUTE_BACKWARD  # :232-239
    geo = HandleGeometry(ttype, p, me, list(members), list(send.counts), list(recv.counts), kernel_type, None)
    if p == 1:                                                                       # :246-253
        return geo

    S = lambda d, r: send_by_member[r].counts[d - 1]
    s = lambda d, r: send_by_member[r].starts[d - 1]
    D = lambda d, r: recv_by_member[r].counts[d - 1]
    d_ = lambda d, r: recv_by_member[r].starts[d - 1]

    def send_box(i, frm):
        """``in%ln(:, i)``, ``in%ls(:, i)`` as computed by rank ``frm`` (:295-336)."""
        if ndims == 2:
            return [D(2, i), S(2, frm)], [d_(2, i), s(2, frm)]
        if ttype == X_TO_Z:
            return [S(1, frm), D(3, i), S(3, frm)], [s(1, frm), d_(3, i), s(3, frm)]
        if ttype in (Z_TO_X, X_TO_Y, Y_TO_Z):
            return [D(3, i), S(2, frm), S(3, frm)], [d_(3, i), s(2, frm), s(3, frm)]
        return [D(2, i), S(2, frm), S(3, frm)], [d_(2, i), s(2, frm), s(3, frm)]

    send_nd = np.zeros((p, 5), dtype=np.int32)
    sdispl = 0
    for i in range(p):
        ln, ls = send_box(i, me)
        send_nd[i, 3] = ln[0] * ls[1] if ttype == X_TO_Z else ls[0]                 # :337-341
        send_nd[i, 0], send_nd[i, 1] = ln[0], ln[1]
        send_nd[i, 2] = ln[2] if ndims == 3 else 1
        send_nd[i, 4] = sdispl                                                       # :413
        cnt = int(np.prod(ln))
        geo.send_counts.append(cnt)
        geo.send_displs.append(sdispl)
        sdispl += cnt

    two_step = ttype in (Y_TO_X, Z_TO_Y) and ndims == 3 and not fused                # :430-433
    if two_step:
        kernel_type = K.KERNEL_PERMUTE_BACKWARD_START
    if fused:                                                                        # :447, abstract_kernel.F90:202-217
        kernel_type = {K.KERNEL_PERMUTE_FORWARD: K.KERNEL_PACK_FORWARD,
                       K.KERNEL_PERMUTE_BACKWARD: K.KERNEL_PACK_BACKWARD}.get(kernel_type, kernel_type)
    geo.pack_kernel = kernel_type
    geo.is_pipelined = pipelined or fused
    geo.is_fused = fused

    recv_nd = np.zeros((p, 5), dtype=np.int32)
    rdispl = 0
    for i in range(p):
        # what member i announces it sends to me (:421-422, 488-489)
        ln_i, _ = send_box(me, i)
        recvsize = int(np.prod(ln_i))
        ln, ls = [0] * ndims, [0] * ndims
        if recvsize > 0:                                                             # :490-531
            if ndims == 2:
                ln, ls = [S(2, i), D(2, me)], [s(2, i), d_(2, me)]
            elif ttype == X_TO_Z:
                ln, ls = [S(3, i), D(2, me), D(3, me)], [s(3, i), d_(2, me), d_(3, me)]
            elif ttype == Z_TO_X:
                ln, ls = [D(1, me), S(3, i), D(3, me)], [d_(1, me), s(3, i), d_(3, me)]
            elif ttype in (X_TO_Y, Y_TO_Z):
                ln, ls = [S(2, i), D(2, me), D(3, me)], [s(2, i), s(2, me), d_(3, me)]
            else:
                ln, ls = [S(3, i), D(2, me), D(3, me)], [s(3, i), s(2, me), d_(3, me)]
        recv_nd[i, 0], recv_nd[i, 1] = ln[0], ln[1]
        recv_nd[i, 2] = ln[2] if ndims == 3 else 1
        recv_nd[i, 3] = rdispl                                                       # :582
        recv_nd[i, 4] = ln[0] * ls[1] if ttype == Z_TO_X else ls[0]                  # :583-588
        geo.recv_counts.append(recvsize)
        geo.recv_displs.append(rdispl)
        rdispl += recvsize

    uk = K.KERNEL_UNPACK                                                             # :618-621
    if geo.is_pipelined:
        uk = K.KERNEL_UNPACK_PIPELINED
    if two_step:
        uk = K.KERNEL_PERMUTE_BACKWARD_END
    if geo.is_pipelined and two_step:
        uk = K.KERNEL_PERMUTE_BACKWARD_END_PIPELINED
    geo.unpack_kernel = uk
    geo.send_nd, geo.recv_nd = send_nd, recv_nd
    return geo


def plan_geometry(dims, comm_dims, ttype: int, pipelined=False, fused=False):
    """Geometry of transposition ``ttype`` for EVERY global rank of the grid.
    Returns ``(pencils_by_rank, geos_by_rank)``."""
    n = int(np.prod(comm_dims))
    pencils = [make_pencils(dims, comm_dims, r) for r in range(n)]
    si, ri = transpose_pencil_ids(ttype)
    cid = transpose_comm_id(ttype)
    geos = []
    for r in range(n):
        members = comm_members(r, comm_dims, cid)
        me = members.index(r)
        geos.append(transpose_geometry(ttype, [pencils[m][si] for m in members],
                                       [pencils[m][ri] for m in members], me, members,
                                       pipelined=pipelined, fused=fused))
    return pencils, geos


# --------------------------------------------------------------------------------------
# User pencils and bricks (plans created from dtfft_pencil_t)
# --------------------------------------------------------------------------------------
def grid_from_boxes(starts, counts):
    """Per-axis 1-D communicators of a user decomposition (``create_1d_comm``,
    src/dtfft_pencil.F90:1034-1081): for every rank r and axis d, the ranks that share start and
    extent on every OTHER axis, ordered by their start along d.  Returns ``(grid, coords)``:
    ``grid[d]`` = communicator size along d (taken from rank 0; the reference aborts unless the
    product equals the world size, :889-899), ``coords[r][d]`` = r's rank in that communicator."""
    P, nd = len(starts), len(starts[0])
    coords = [[0] * nd for _ in range(P)]
    grid = [1] * nd
    for r in range(P):
        for d in range(nd):
            line = sorted((starts[i][d], i) for i in range(P)
                          if i == r or (all(starts[i][j] == starts[r][j] and counts[i][j] == counts[r][j]
                                            for j in range(nd) if j != d)
                                        and (starts[i][d] != starts[r][d]
                                             # a rank without points along d may share its start (:1066-1068)
                                             or counts[i][d] == 0 or counts[r][d] == 0)))
            coords[r][d] = line.index((starts[r][d], r))
            if r == 0:
                grid[d] = len(line)
    return grid, coords


def pencils_from_x(dims, comm_dims, coords, x_starts, x_counts):
    """X / Y / Z pencils when the X pencil is prescribed (user pencil or from_bricks):
    ``create_pencils_and_comm`` with ``ipencil`` (src/dtfft_transpose_plan.F90:1113-1122) +
    the carry-over rule of ``pencil%create`` (src/dtfft_pencil.F90:136-163): the axis that stays
    distributed between consecutive pencils keeps its split, the newly distributed axis gets the
    standard ``get_local_size`` split."""
    nd = len(dims)
    X = Pencil(1, list(x_starts), list(x_counts))
    if nd == 2:
        s, c = local_size(dims[0], comm_dims[1], coords[1])
        return [X, Pencil(2, [0, s], [dims[1], c])]
    s, c = local_size(dims[0], comm_dims[1], coords[1])
    Y = Pencil(2, [0, X.starts[2], s], [dims[1], X.counts[2], c])
    s2, c2 = local_size(dims[1], comm_dims[2], coords[2])
    Z = Pencil(3, [0, Y.starts[2], s2], [dims[2], Y.counts[2], c2])
    return [X, Y, Z]


def from_bricks(brick_starts, brick_counts, cuda=True):
    """``pencil_init%from_bricks`` (src/dtfft_pencil.F90:520-775) for a non-cartesian communicator.
    Input: every rank's brick (x fastest).  Returns ``(dims, comm_dims, coords, x_starts, x_counts,
    brick_grid, brick_coords)`` of the X pencils the bricks are reshaped to: the ``P0`` ranks that
    share a (y, z) footprint split the brick's z (if it is long enough: > tile * P0), else y,
    else a ``y_size x z_size`` factorisation (:607-655); new grid = ``1 x P1*y_size x P2*z_size``."""
    P, nd = len(brick_starts), len(brick_starts[0])
    dims = [max(brick_starts[r][d] + brick_counts[r][d] for r in range(P)) for d in range(nd)]
    bgrid, bcoords = grid_from_boxes(brick_starts, brick_counts)
    fast = bgrid[0]
    tile = DEF_TILE_SIZE if cuda else 4
    c0 = brick_counts[0]  # rank 0 decides and broadcasts (:640-645)
    y_size = z_size = None
    if nd == 3 and c0[2] > tile * fast:
        y_size, z_size = 1, fast
    elif c0[1] > tile * fast or nd == 2:
        y_size, z_size = fast, 1
    else:
        for i in range(2, fast + 1):
            if fast % i:
                continue
            if c0[2] < tile * i or c0[2] % i:
                continue
            if c0[1] < tile * i or c0[1] % i:
                continue
            y_size, z_size = min(i, fast // i), max(i, fast // i)
            break
    if y_size is None:
        y_size, z_size = dims_create(fast, 2, [0, 0])
    comm_dims = [1, bgrid[1] * y_size, (bgrid[2] * z_size) if nd == 3 else 1][:nd]
    coords, xs, xc = [], [], []
    for r in range(P):
        a, b = bcoords[r][0], bcoords[r][1]
        c = bcoords[r][2] if nd == 3 else 0
        if y_size == 1:
            ay, az = 0, a
        elif z_size == 1:
            ay, az = a, 0
        else:  # MPI_Cart_create(y_size x z_size) over the x-line, row-major (:689-706)
            ay, az = a // z_size, a % z_size
        s1, n1 = local_size(brick_counts[r][1], y_size, ay)
        st, ct = [0, brick_starts[r][1] + s1], [dims[0], n1]
        co = [0, b * y_size + ay]
        if nd == 3:
            s2, n2 = local_size(brick_counts[r][2], z_size, az)
            st.append(brick_starts[r][2] + s2)
            ct.append(n2)
            co.append(c * z_size + az)
        coords.append(co)
        xs.append(st)
        xc.append(ct)
    return dims, comm_dims, coords, xs, xc, bgrid, bcoords


def z_bricks(dims, comm_dims, coords, last_pencils, brick_grid):
    """``bricks(2)`` of the reshape plan (src/dtfft_reshape_plan.F90:150-182): groups of ``c``
    consecutive ranks of the last grid dimension (c = brick-grid size along the last axis) split
    the pencil's aligned axis among themselves and pool their share of the slowest axis.
    ``last_pencils[r]`` = rank r's last pencil (Z pencil in 3-D, Y pencil in 2-D)."""
    P, nd = len(coords), len(dims)
    last = nd - 1
    csize = brick_grid[last]
    out = []
    for r in range(P):
        grp, k = divmod(coords[r][last], csize)
        s0, c0 = local_size(dims[last], csize, k)
        mates = [q for q in range(P) if coords[q][last] // csize == grp
                 and all(coords[q][d] == coords[r][d] for d in range(1, nd) if d != last)]
        lo = min(last_pencils[q].starts[nd - 1] for q in mates)
        cnt = sum(last_pencils[q].counts[nd - 1] for q in mates)
        L = last_pencils[r]
        if nd == 3:
            out.append(Pencil(3, [s0, L.starts[1], lo], [c0, L.counts[1], cnt]))
        else:
            out.append(Pencil(2, [s0, lo], [c0, cnt]))
    return out


# --------------------------------------------------------------------------------------
# Brick <-> pencil reshapes: neighbor_data, strategies, pack-free / unpack-free
# --------------------------------------------------------------------------------------
X_BRICKS_TO_PENCILS, X_PENCILS_TO_BRICKS, Z_PENCILS_TO_BRICKS, Z_BRICKS_TO_PENCILS = 11, 12, 13, 14


@dataclass
class ReshapeGeometry(HandleGeometry):
    is_pack_free: bool = False
    is_unpack_free: bool = False
    reshape_strat: int = 0


def reshape_geometry(rtype: int, send_by_member, recv_by_member, me: int, members,
                     pipelined: bool = False) -> ReshapeGeometry:
    """What ``reshape_handle_generic%create`` derives for a brick <-> pencil reshape
    (src/dtfft_reshape_handle_generic.F90): send boxes :343-376, send-side displacements by
    strategy :380-404, recv boxes :536-567, recv-side displacements :590-611, strategy :267-289,
    ``is_pack_free`` :261-266, ``is_unpack_free`` :479-484 (both are all-reduced over the
    communicator: every member's predicate must hold), kernel kinds :425-447, 618-624.
    Both sides of a reshape store their axes in the same order, so the formulas work on local
    axes 1..3 whatever the global axes are (X bricks / X pencils: x,y,z; Z bricks / Z pencils: z,x,y)."""
    p = len(members)
    send, recv = send_by_member[me], recv_by_member[me]
    ndims = len(send.counts)
    to_pencils = rtype in (X_BRICKS_TO_PENCILS, Z_BRICKS_TO_PENCILS)
    geo = ReshapeGeometry(0, p, me, list(members), list(send.counts), list(recv.counts), K.KERNEL_COPY, None)
    if p == 1:                                                                       # :246-253
        return geo
    S = lambda d, r: send_by_member[r].counts[d - 1]
    s = lambda d, r: send_by_member[r].starts[d - 1]
    D = lambda d, r: recv_by_member[r].counts[d - 1]
    d_ = lambda d, r: recv_by_member[r].starts[d - 1]
    dims_of = range(1, ndims + 1)

    def pack_free_of(m):                                                             # :261-262, 479-480
        return ndims == 2 or all(S(2, m) == D(2, i) for i in range(p))

    geo.is_pack_free = to_pencils and all(pack_free_of(m) for m in range(p))         # :264-266
    geo.is_unpack_free = (not to_pencils) and all(pack_free_of(m) for m in range(p))  # :481-484
    if ndims == 2:                                                                   # :289
        strat = 1
    else:                                                                            # :267-288
        if to_pencils:
            zslab = all(S(2, me) == D(2, i) for i in range(p))
            yslab = all(S(3, me) == D(3, i) for i in range(p))
        else:
            zslab = all(S(2, i) == D(2, me) for i in range(p))
            yslab = all(S(3, i) == D(3, me) for i in range(p))
        strat = 1 if zslab else (2 if yslab else 3)
    geo.reshape_strat = strat

    def send_box(i, frm):                                                            # :343-376
        if to_pencils:
            ln = [S(1, frm)] + [D(d, i) for d in dims_of if d > 1]
            ls = [s(1, frm)] + [d_(d, i) for d in dims_of if d > 1]
        else:
            ln = [D(1, i)] + [S(d, frm) for d in dims_of if d > 1]
            ls = [d_(1, i)] + [s(d, frm) for d in dims_of if d > 1]
        return ln, ls

    send_nd = np.zeros((p, 5), dtype=np.int32)
    sdispl = ssdispl = 0
    for i in range(p):
        ln, ls = send_box(i, me)
        if to_pencils:                                                               # :380-391
            if strat == 1:
                send_nd[i, 3] = ssdispl
                ssdispl += int(np.prod(ln))
            elif strat == 2:
                send_nd[i, 3] = ssdispl
                ssdispl += ln[0] * ln[1]
            else:
                ssdispl = (d_(3, i) - s(3, me)) * S(1, me) * S(2, me) + (d_(2, i) - s(2, me)) * S(1, me)
                send_nd[i, 3] = ssdispl
        else:                                                                        # :402
            send_nd[i, 3] = ls[0]
        send_nd[i, 0], send_nd[i, 1] = ln[0], ln[1]
        send_nd[i, 2] = ln[2] if ndims == 3 else 1
        send_nd[i, 4] = sdispl                                                       # :413
        cnt = int(np.prod(ln))
        geo.send_counts.append(cnt)
        geo.send_displs.append(sdispl)
        sdispl += cnt

    recv_nd = np.zeros((p, 5), dtype=np.int32)
    rdispl = rrdispl = 0
    for i in range(p):
        ln_i, _ = send_box(me, i)                                                    # :421-422, 488-489
        recvsize = int(np.prod(ln_i))
        ln, ls = [0] * ndims, [0] * ndims
        if recvsize > 0:                                                             # :536-567
            if to_pencils:
                ln = [S(1, i)] + [D(d, me) for d in dims_of if d > 1]
                ls = [s(1, i)] + [d_(d, me) for d in dims_of if d > 1]
            else:
                ln = [D(1, me)] + [S(d, i) for d in dims_of if d > 1]
                ls = [d_(1, me)] + [s(d, i) for d in dims_of if d > 1]
        recv_nd[i, 0], recv_nd[i, 1] = ln[0], ln[1]
        recv_nd[i, 2] = ln[2] if ndims == 3 else 1
        recv_nd[i, 3] = rdispl                                                       # :582
        if to_pencils:                                                               # :590-591
            recv_nd[i, 4] = ls[0]
        elif strat in (1, 2):                                                        # :594-597
            recv_nd[i, 4] = rrdispl
            rrdispl += ln[0] * ln[1]
        else:                                                                        # :598-600
            rrdispl = (abs(d_(3, me) - s(3, i)) * D(1, me) * D(2, me) + abs(d_(2, me) - s(2, i)) * D(1, me)) \
                if ndims == 3 else 0
            recv_nd[i, 4] = rrdispl
        geo.recv_counts.append(recvsize)
        geo.recv_displs.append(rdispl)
        rdispl += recvsize

    geo.is_pipelined = pipelined
    geo.pack_kernel = K.KERNEL_PACK                                                  # :437-441 (non-fused backends)
    if geo.is_pack_free:                                                             # :442
        geo.pack_kernel = K.KERNEL_DUMMY
    geo.unpack_kernel = K.KERNEL_UNPACK_PIPELINED if pipelined else K.KERNEL_UNPACK  # :618-619
    if geo.is_unpack_free:                                                           # :623
        geo.unpack_kernel = K.KERNEL_DUMMY
    geo.send_nd, geo.recv_nd = send_nd, recv_nd
    return geo


def reshape_members(rank: int, rtype: int, brick_grid, brick_coords, coords, x_pencils):
    """Global ranks of the 1-D communicator of a reshape, in communicator order
    (src/dtfft_reshape_plan.F90:160-171).  X side: the bricks of one x-line of the brick grid
    (``ipencil%comms(1)``), re-ranked by where their X pencil starts -- ``MPI_Comm_split`` key
    ``starts(2) + starts(3) * counts(2) * init_grid(2)``, ties by the old rank (= x coordinate).
    Z side: ``NEIGHBOR_GROUP`` of ``c`` consecutive ranks of the last grid dimension, c = size of
    the brick grid along the last axis (``create_custom_comm``)."""
    P, nd = len(coords), len(coords[0])
    last = nd - 1
    if rtype in (X_BRICKS_TO_PENCILS, X_PENCILS_TO_BRICKS):
        def key(q):
            xp = x_pencils[q]
            k = xp.starts[1]
            if nd == 3:
                k += xp.starts[2] * xp.counts[1] * brick_grid[1]
            return (k, brick_coords[q][0])
        line = [q for q in range(P) if all(brick_coords[q][d] == brick_coords[rank][d] for d in range(1, nd))]
        return sorted(line, key=key)
    csize = brick_grid[last]
    grp = [q for q in range(P) if coords[q][last] // csize == coords[rank][last] // csize
           and all(coords[q][d] == coords[rank][d] for d in range(1, nd) if d != last)]
    return sorted(grp, key=lambda q: coords[q][last])
