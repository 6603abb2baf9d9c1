"""Oracle (test infrastructure): the reference's host MPI-DATATYPE path, restated.

``reshape_handle_datatype`` (src/dtfft_reshape_handle_datatype.F90) moves a transposition or a brick
reshape with ONE ``MPI_Alltoall(w)`` whose send / receive datatypes do the packing and unpacking:
``create`` (:125-270) builds, per peer, the derived datatypes of ``create_transpose_2d`` (:436-476),
``create_forw_permutation`` (:478-526), ``create_back_permutation`` (:528-575), ``create_transpose_XZ``
(:577-628), ``create_transpose_ZX`` (:630-680), ``create_reshape_32`` (:682-737), ``create_reshape_23``
(:739-800), ``create_reshape_21`` (:802-825), ``create_reshape_12`` (:827-850) and the running byte
displacements; ``execute`` (:278-398) posts the exchange.

This module restates exactly that -- a small typemap algebra for the MPI constructors the reference
calls (``MPI_Type_vector``, ``_create_hvector``, ``_contiguous``, ``_create_indexed_block``,
``_create_resized``) and the all-to-all(w) on simulated ranks -- so that the "ground truth" the product
is compared with (global-array slicing, oracle/pipeline.py) is itself checked against the reference's
own datatype construction, in both ``DTFFT_TRANSPOSE_MODE_PACK`` and ``_UNPACK``.  MPI semantics used:
a derived datatype is a sequence of (byte offset) entries in typemap order plus an extent; sending
``count = 1`` of a type at displacement d reads the elements at d + offsets in that order, receiving
writes the incoming elements at d + offsets of the receive type in that order.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import layout as L

PACK, UNPACK = 15, 16  # dtfft_transpose_mode_t, include/dtfft_config.h.in:180-181


@dataclass
class Datatype:
    offsets: np.ndarray  # byte offsets in typemap order
    extent: int          # bytes


def base(es: int) -> Datatype:
    return Datatype(np.zeros(1, np.int64), es)


def vector(count: int, blocklength: int, stride: int, old: Datatype) -> Datatype:
    """MPI_Type_vector: `count` blocks of `blocklength` olds, block starts `stride` olds apart."""
    i = np.arange(count, dtype=np.int64)[:, None, None] * stride * old.extent
    j = np.arange(blocklength, dtype=np.int64)[None, :, None] * old.extent
    offs = (i + j + old.offsets[None, None, :]).reshape(-1)
    extent = ((count - 1) * stride + blocklength) * old.extent if count > 0 and blocklength > 0 else 0
    return Datatype(offs, extent)


def hvector(count: int, blocklength: int, stride_bytes: int, old: Datatype) -> Datatype:
    """MPI_Type_create_hvector: like vector with the stride given in bytes."""
    i = np.arange(count, dtype=np.int64)[:, None, None] * stride_bytes
    j = np.arange(blocklength, dtype=np.int64)[None, :, None] * old.extent
    offs = (i + j + old.offsets[None, None, :]).reshape(-1)
    extent = (count - 1) * stride_bytes + blocklength * old.extent if count > 0 and blocklength > 0 else 0
    return Datatype(offs, extent)


def contiguous(count: int, old: Datatype) -> Datatype:
    return vector(count, 1, 1, old)


def indexed_block(blocklength: int, displs, old: Datatype) -> Datatype:
    """MPI_Type_create_indexed_block: equally sized blocks at displacements given in olds."""
    d = np.asarray(displs, dtype=np.int64)[:, None, None] * old.extent
    j = np.arange(blocklength, dtype=np.int64)[None, :, None] * old.extent
    offs = (d + j + old.offsets[None, None, :]).reshape(-1)
    extent = int(offs.max() + old.extent - min(0, offs.min())) if offs.size else 0
    return Datatype(offs, extent)


def resized(old: Datatype, extent: int) -> Datatype:
    """MPI_Type_create_resized with lb = 0."""
    return Datatype(old.offsets, int(extent))


# ---- the reference's creators: (send pencil, peer's send counts, recv pencil, peer's recv counts) ----
# sc / rc = counts of the PEER's send / recv pencil (1-based dimension numbers as in the reference).
def transpose_2d(S, sc, R, rc, mode, es):                                    # :436-476
    b = base(es)
    send_displ, recv_displ = rc[1] * es, sc[1] * es
    if mode == UNPACK:
        sdt = resized(vector(S[1], rc[1], S[0], b), send_displ)
        t2 = resized(vector(R[1], 1, R[0], b), es)
        rdt = contiguous(sc[1], t2)
    else:
        t2 = resized(vector(S[1], 1, S[0], b), es)
        sdt = contiguous(rc[1], t2)
        rdt = resized(vector(R[1], sc[1], R[0], b), recv_displ)
    return sdt, send_displ, rdt, recv_displ


def forw_permutation(S, sc, R, rc, mode, es):                                # :478-526  X->Y, Y->Z
    b = base(es)
    send_displ, recv_displ = rc[2] * es, sc[1] * es
    if mode == UNPACK:
        sdt = resized(vector(S[1] * S[2], rc[2], S[0], b), send_displ)
        t2 = resized(vector(R[2], 1, R[0] * R[1], b), es)
        t3 = contiguous(sc[1], t2)
        t4 = hvector(R[1], 1, R[0] * es, t3)
        rdt = resized(t4, recv_displ)
    else:
        t2 = resized(vector(S[1] * S[2], 1, S[0], b), es)
        sdt = contiguous(rc[2], t2)
        t2 = resized(vector(R[1], sc[1], R[0], b), recv_displ)
        t3 = hvector(R[2], 1, R[0] * R[1] * es, t2)
        rdt = resized(t3, recv_displ)
    return sdt, send_displ, rdt, recv_displ


def back_permutation(S, sc, R, rc, mode, es):                                # :528-575  Y->X, Z->Y
    b = base(es)
    send_displ, recv_displ = rc[1] * es, sc[2] * es
    if mode == UNPACK:
        sdt = resized(vector(S[1] * S[2], rc[1], S[0], b), send_displ)
        t2 = resized(vector(R[1] * R[2], 1, R[0], b), es)
        t3 = contiguous(sc[2], t2)
        rdt = resized(t3, recv_displ)
    else:
        t2 = resized(vector(S[2], 1, S[0] * S[1], b), es)
        t3 = contiguous(rc[1], t2)
        t4 = hvector(S[1], 1, S[0] * es, t3)
        sdt = resized(t4, send_displ)
        rdt = resized(vector(R[1] * R[2], sc[2], R[0], b), recv_displ)
    return sdt, send_displ, rdt, recv_displ


def transpose_xz(S, sc, R, rc, mode, es):                                    # :577-628
    b = base(es)
    send_displ, recv_displ = S[0] * rc[2] * es, sc[2] * es
    if mode == UNPACK:
        t2 = resized(vector(S[2], S[0], S[0] * S[1], b), S[0] * es)
        sdt = contiguous(rc[2], t2)
        t2 = resized(vector(R[1], 1, R[0], b), es)
        t3 = contiguous(sc[2], t2)
        t4 = hvector(R[2], 1, R[0] * R[1] * es, t3)
        rdt = resized(t4, recv_displ)
    else:
        t2 = resized(vector(S[2], 1, S[0] * S[1], b), es)
        t3 = contiguous(S[0], t2)
        t4 = hvector(rc[2], 1, S[0] * es, t3)
        sdt = resized(t4, send_displ)
        rdt = resized(vector(R[1] * R[2], sc[2], R[0], b), recv_displ)
    return sdt, send_displ, rdt, recv_displ


def transpose_zx(S, sc, R, rc, mode, es):                                    # :630-680
    b = base(es)
    send_displ, recv_displ = rc[2] * es, R[0] * sc[2] * es
    if mode == UNPACK:
        sdt = resized(vector(S[1] * S[2], rc[2], S[0], b), send_displ)
        t2 = resized(vector(R[2], 1, R[0] * R[1], b), es)
        t3 = contiguous(R[0], t2)
        t4 = hvector(sc[2], 1, R[0] * es, t3)
        rdt = resized(t4, recv_displ)
    else:
        t2 = resized(vector(S[1] * S[2], 1, S[0], b), es)
        sdt = contiguous(rc[2], t2)
        rdt = resized(vector(R[2], R[0] * sc[2], R[0] * R[1], b), recv_displ)
    return sdt, send_displ, rdt, recv_displ


def reshape_32(S, s_starts, sc, R, rc, r_starts, es, strat):                 # :682-737 bricks -> pencils, 3-D
    b = base(es)
    if strat == 1:
        send_displ = S[0] * S[1] * rc[2] * es
        sdt = contiguous(S[0] * S[1] * rc[2], b)
    elif strat == 2:
        send_displ = S[0] * rc[1] * es
        sdt = resized(vector(rc[2], S[0] * rc[1], S[0] * S[1], b), send_displ)
    else:
        send_displ = 0
        displs = []
        dsp0 = (r_starts[2] - s_starts[2]) * S[0] * S[1] + (r_starts[1] - s_starts[1]) * S[0]
        for k in range(rc[2]):
            for j in range(rc[1]):
                displs.append(dsp0 + k * S[0] * S[1] + j * S[0])
        sdt = resized(indexed_block(S[0], displs, b), 0)
    if any(c == 0 for c in rc):
        send_displ = 0
    recv_displ = sc[0] * es
    rdt = resized(vector(R[1] * R[2], sc[0], R[0], b), recv_displ)
    if any(c == 0 for c in sc):
        recv_displ = 0
    return sdt, send_displ, rdt, recv_displ


def reshape_23(S, sc, s_starts_peer, R, r_starts, rc, es, strat, recv_displ_in):  # :739-800 pencils -> bricks, 3-D
    b = base(es)
    send_displ = rc[0] * es
    sdt = resized(vector(S[1] * S[2], rc[0], S[0], b), send_displ)
    if strat == 1:
        recv_displ = R[0] * R[1] * sc[2] * es
        rdt = contiguous(R[0] * R[1] * sc[2], b)
    elif strat == 2:
        recv_displ = R[0] * sc[1] * es
        rdt = resized(vector(sc[2], R[0] * sc[1], R[0] * R[1], b), recv_displ)
    else:
        dsp0 = (abs(r_starts[2] - s_starts_peer[2]) * R[0] * R[1] + abs(r_starts[1] - s_starts_peer[1]) * R[0]
                - recv_displ_in // es)
        displs = []
        for k in range(sc[2]):
            for j in range(sc[1]):
                displs.append(dsp0 + k * R[0] * R[1] + j * R[0])
        rdt = resized(indexed_block(R[0], displs, b), 0)
        recv_displ = R[0] * sc[1] * sc[2] * es
    if any(c == 0 for c in R):
        send_displ = 0
    return sdt, send_displ, rdt, recv_displ


def reshape_21(S, sc, R, rc, es):                                            # :802-825 bricks -> slabs, 2-D
    b = base(es)
    send_displ = S[0] * rc[1] * es
    sdt = contiguous(S[0] * rc[1], b)
    recv_displ = sc[0] * es
    rdt = resized(vector(R[1], sc[0], R[0], b), recv_displ)
    return sdt, send_displ, rdt, recv_displ


def reshape_12(S, sc, R, rc, es):                                            # :827-850 slabs -> bricks, 2-D
    b = base(es)
    send_displ = rc[0] * es
    sdt = resized(vector(S[1], rc[0], S[0], b), send_displ)
    recv_displ = R[0] * sc[1] * es
    rdt = contiguous(R[0] * sc[1], b)
    return sdt, send_displ, rdt, recv_displ


def create(send_by_member, recv_by_member, me, es, ttype=0, rtype=0, mode=PACK):
    """``reshape_handle_datatype%create`` (:125-270) for member ``me``: per-peer send / receive datatypes
    and byte displacements (the non-even form; the even form is the special case of equal entries)."""
    p = len(send_by_member)
    send, recv = send_by_member[me], recv_by_member[me]
    S, R = list(send.counts), list(recv.counts)
    nd = len(S)
    to_pencils = rtype in (L.X_BRICKS_TO_PENCILS, L.Z_BRICKS_TO_PENCILS)
    strat = 0
    if rtype and nd == 3:                                                    # :177-199
        if to_pencils:
            z = all(S[1] == recv_by_member[i].counts[1] for i in range(p))
            y = all(S[2] == recv_by_member[i].counts[2] for i in range(p))
        else:
            z = all(send_by_member[i].counts[1] == R[1] for i in range(p))
            y = all(send_by_member[i].counts[2] == R[2] for i in range(p))
        strat = 1 if z else (2 if y else 3)
    sdts, rdts, sdispls, rdispls = [], [], [0], [0]
    recv_displ = 0
    for i in range(p):
        sc, rc = list(send_by_member[i].counts), list(recv_by_member[i].counts)
        if ttype:
            if nd == 2:
                out = transpose_2d(S, sc, R, rc, mode, es)
            elif ttype in (L.X_TO_Y, L.Y_TO_Z):
                out = forw_permutation(S, sc, R, rc, mode, es)
            elif ttype in (L.Y_TO_X, L.Z_TO_Y):
                out = back_permutation(S, sc, R, rc, mode, es)
            elif ttype == L.X_TO_Z:
                out = transpose_xz(S, sc, R, rc, mode, es)
            else:
                out = transpose_zx(S, sc, R, rc, mode, es)
        elif to_pencils:
            out = reshape_21(S, sc, R, rc, es) if nd == 2 else \
                reshape_32(S, send.starts, sc, R, rc, recv_by_member[i].starts, es, strat)
        else:
            out = reshape_12(S, sc, R, rc, es) if nd == 2 else \
                reshape_23(S, sc, send_by_member[i].starts, R, recv.starts, rc, es, strat, recv_displ)
        sdt, send_displ, rdt, rd = out
        sdts.append(sdt)
        rdts.append(rdt)
        if i < p - 1:                                                        # :238-250
            sdispls.append(sdispls[i] + (0 if any(c == 0 for c in S) else send_displ))
            rdispls.append(rdispls[i] + (0 if any(c == 0 for c in R) else rd))
            recv_displ = rdispls[i + 1]
    return sdts, sdispls, rdts, rdispls


def exchange(inputs, send_layouts, recv_layouts, groups, es, ttype=0, rtype=0, mode=PACK, out_sizes=None):
    """``execute`` (:278-398): MPI_Alltoallw with count 1 per peer on simulated ranks.
    ``inputs[r]`` flat element array of global rank r, ``send_layouts[r]`` / ``recv_layouts[r]`` its
    source / destination :class:`oracle.layout.Pencil`, ``groups[r]`` the global ranks of its 1-D
    communicator in communicator order.  Returns the destination arrays."""
    n = len(inputs)
    dtype = inputs[0].dtype
    assert dtype.itemsize == es
    outs = [np.full(out_sizes[r] if out_sizes else recv_layouts[r].size, -7, dtype) for r in range(n)]
    made = {}
    for r in range(n):
        g = groups[r]
        made[r] = create([send_layouts[m] for m in g], [recv_layouts[m] for m in g], g.index(r), es, ttype, rtype, mode)
    for r in range(n):          # sender
        g = groups[r]
        sdts, sdispls, _, _ = made[r]
        for i, peer in enumerate(g):
            if send_layouts[r].size == 0 or recv_layouts[peer].size == 0:
                continue
            _, _, rdts, rdispls = made[peer]
            j = groups[peer].index(r)
            src_idx = (sdispls[i] + sdts[i].offsets) // es
            dst_idx = (rdispls[j] + rdts[j].offsets) // es
            assert src_idx.size == dst_idx.size, (r, peer, src_idx.size, dst_idx.size)
            outs[peer][dst_idx] = inputs[r][src_idx]
    return outs
