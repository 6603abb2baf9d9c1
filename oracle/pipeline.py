"""Oracle (test infrastructure): whole transpositions on simulated ranks.

Two routes to the same answer:

* :func:`transpose_generic` -- the 3-step schedule of ``reshape_handle_generic%execute``
  (src/dtfft_reshape_handle_generic.F90:695-759): pack kernel, all-to-all(v) with the
  float-unit displacements of ``abstract_backend`` (here: element slices copied between
  the simulated ranks' buffers, src/dtfft_backend_nccl.F90:108-119), unpack kernel.
* :func:`transpose_datatype` -- what the host MPI-datatype path delivers
  (src/dtfft_reshape_handle_datatype.F90:436-673): rank r's destination pencil is the
  global array restricted to r's destination box, stored in destination axis order.

They must agree bit for bit; that is the reference's own "generic == datatype" contract
and the pin the CUDA product is tested against.
"""
from __future__ import annotations

import numpy as np

from . import kernels as K
from . import layout as L


def global_array(dims, dtype, seed=1234, kind="random"):
    """Synthetic global field G[x,y,z] (Fortran order).  ``kind='index'`` encodes the
    global linear index so a misplaced element is identifiable (reference analogue:
    ``in(i) = i``, src/tests/test_host_kernels.F90:35-37); ``'random'`` mirrors the
    ``random_number`` fill of tests/test_utils.F90:184-196."""
    n = int(np.prod(dims))
    dtype = np.dtype(dtype)
    if kind == "index":
        base = np.arange(n, dtype=np.float64)
        g = (base + 1j * (base + 0.5)).astype(dtype) if dtype.kind == "c" else base.astype(dtype)
    else:
        rng = np.random.default_rng(seed)
        if dtype.kind == "c":
            g = (rng.random(n) + 1j * rng.random(n)).astype(dtype)
        else:
            g = rng.random(n).astype(dtype)
    return np.asfortranarray(g.reshape(dims, order="F"))


def pencil_slice(G, pencil: L.Pencil):
    """Local flat buffer (Fortran order, pencil axis order) of ``pencil`` cut from G."""
    ndims = G.ndim
    dperm, _ = L.permutations(ndims)
    axes = dperm[pencil.aligned_dim - 1]
    Gp = G.transpose(axes)
    sl = tuple(slice(s, s + c) for s, c in zip(pencil.starts, pencil.counts))
    return np.ascontiguousarray(Gp[sl].reshape(-1, order="F"))


def transpose_datatype(G, dims, comm_dims, ttype):
    """Per-rank destination buffers by global-array slicing (the datatype-path result)."""
    n = int(np.prod(comm_dims))
    _, ri = L.transpose_pencil_ids(ttype)
    return [pencil_slice(G, L.make_pencils(dims, comm_dims, r)[ri]) for r in range(n)]


def scatter_input(G, dims, comm_dims, ttype):
    n = int(np.prod(comm_dims))
    si, _ = L.transpose_pencil_ids(ttype)
    return [pencil_slice(G, L.make_pencils(dims, comm_dims, r)[si]) for r in range(n)]


def alloc_size(dims, comm_dims, rank):
    """``get_local_sizes`` alloc_size: max pencil volume (src/dtfft_pencil.F90:462)."""
    return max(p.size for p in L.make_pencils(dims, comm_dims, rank))


def transpose_generic(inputs, dims, comm_dims, ttype, pipelined=False, fused=False, execute=K.execute):
    """Run the generic pack -> exchange -> unpack schedule on every simulated rank.

    ``inputs[r]`` is rank r's source pencil (flat).  Buffers are sized to the plan's
    alloc_size like the reference's.  Returns the list of destination buffers
    (trimmed to the destination pencil volume)."""
    n = int(np.prod(comm_dims))
    pencils, geos = L.plan_geometry(dims, comm_dims, ttype, pipelined=pipelined, fused=fused)
    _, ri = L.transpose_pencil_ids(ttype)
    dtype = inputs[0].dtype
    sizes = [max(p.size for p in pencils[r]) for r in range(n)]
    a = [np.zeros(sizes[r], dtype) for r in range(n)]     # "in"
    b = [np.full(sizes[r], -7, dtype) for r in range(n)]  # "out"
    for r in range(n):
        a[r][: inputs[r].size] = inputs[r]

    # step 1: pack kernel, in -> out (:752) ; fused: one launch per peer (pack_forward/backward)
    for r in range(n):
        g = geos[r]
        if g.pack_kernel in (K.KERNEL_PACK_FORWARD, K.KERNEL_PACK_BACKWARD):
            for i in range(g.comm_size):
                execute(g.pack_kernel, g.send_dims, a[r], b[r], g.send_nd, i + 1)
        else:
            execute(g.pack_kernel, g.send_dims, a[r], b[r], g.send_nd)
    if geos[0].comm_size == 1:
        return [b[r][: pencils[r][ri].size] for r in range(n)]

    # step 2: exchange out -> in (:755); counts/displs in elements here, the product uses
    # the same numbers scaled to 4-byte floats (src/dtfft_abstract_backend.F90:160-183)
    for r in range(n):
        g = geos[r]
        for i, peer in enumerate(g.members):
            gp = geos[peer]
            cnt = g.recv_counts[i]
            assert cnt == gp.send_counts[gp.members.index(r)]
            so = gp.send_displs[gp.members.index(r)]
            ro = g.recv_displs[i]
            a[r][ro: ro + cnt] = b[peer][so: so + cnt]

    # step 3: unpack kernel, in -> out (:758)
    for r in range(n):
        g = geos[r]
        if g.unpack_kernel in K.PER_NEIGHBOR_KERNELS:
            for i in range(g.comm_size):
                execute(g.unpack_kernel, g.recv_dims, a[r], b[r], g.recv_nd, i + 1)
        else:
            execute(g.unpack_kernel, g.recv_dims, a[r], b[r], g.recv_nd)
    return [b[r][: pencils[r][ri].size] for r in range(n)]


# --------------------------------------------------------------------------------------
# Ground truth for ANY exchange (transposition or brick reshape) + emulation of the product's
# fused-store geometry
# --------------------------------------------------------------------------------------
def redistribute(G, pencils_dst):
    """What the host MPI-datatype path delivers for any redistribution
    (src/dtfft_reshape_handle_datatype.F90:436-847): every rank ends up with the global array
    restricted to its destination box, stored in the destination axis order."""
    return [pencil_slice(G, p) for p in pencils_dst]


def apply_boxes(src, dsts, boxes, members):
    """Emulate the NVLINK_FUSED kernel of one rank on the host: ``boxes[i]`` =
    (n0 n1 n2 in_off out_off is1 is2 os0 os1 os2) places the part of ``src`` owned by member i
    straight into ``dsts[members[i]]`` (include/dtfft_b200_api.h: dtfftb_plan_describe_exchange)."""
    for i, b in enumerate(boxes):
        n0, n1, n2, ioff, ooff, is1, is2, os0, os1, os2 = (int(v) for v in b)
        if n0 <= 0 or n1 <= 0 or n2 <= 0:
            continue
        a, bb, c = np.meshgrid(np.arange(n0), np.arange(n1), np.arange(n2), indexing="ij")
        iidx = (ioff + a + bb * is1 + c * is2).ravel()
        oidx = (ooff + a * os0 + bb * os1 + c * os2).ravel()
        dsts[members[i]][oidx] = src[iidx]


def _exchange(geos, send_bufs, recv_bufs):
    """all-to-all(v) on simulated ranks: member i of my communicator sends me
    ``send_counts`` elements from its ``send_displs`` into my ``recv_displs``
    (src/dtfft_backend_nccl.F90:108-119, element units here)."""
    for r, g in enumerate(geos):
        for i, peer in enumerate(g.members):
            gp = geos[peer]
            j = gp.members.index(r)
            cnt = g.recv_counts[i]
            assert cnt == gp.send_counts[j], (r, peer)
            so, ro = gp.send_displs[j], g.recv_displs[i]
            recv_bufs[r][ro: ro + cnt] = send_bufs[peer][so: so + cnt]


def reshape_generic(inputs, geos, out_sizes, buf_sizes, execute=K.execute):
    """The reference's reshape schedule (src/dtfft_reshape_handle_generic.F90:695-759) for a brick
    <-> pencil reshape on simulated ranks, including the pack-free (:711-716, 734-740) and
    unpack-free (:717-722, 742-746) shortcuts.  ``geos[r]`` = :func:`oracle.layout.reshape_geometry`
    of global rank r, ``inputs[r]`` its source array (flat), ``buf_sizes[r]`` the plan's alloc size.
    Returns the destination arrays trimmed to ``out_sizes[r]``."""
    n = len(geos)
    dtype = inputs[0].dtype
    a = [np.zeros(buf_sizes[r], dtype) for r in range(n)]        # "in"
    b = [np.full(buf_sizes[r], -7, dtype) for r in range(n)]     # "out"
    w = [np.full(buf_sizes[r], -9, dtype) for r in range(n)]     # "aux" (kwargs%p1)
    for r in range(n):
        a[r][: inputs[r].size] = inputs[r]
    if geos[0].comm_size == 1:
        return [a[r][: out_sizes[r]].copy() for r in range(n)]

    def run(kernel, dims, src, dst, nd):
        if kernel == K.KERNEL_DUMMY:
            return
        if kernel in K.PER_NEIGHBOR_KERNELS:
            for i in range(nd.shape[0]):
                execute(kernel, dims, src, dst, nd, i + 1)
        else:
            execute(kernel, dims, src, dst, nd)

    g0 = geos[0]
    if g0.is_pipelined:
        if g0.is_pack_free:        # in -> aux exchange, aux -> out unpack
            _exchange(geos, a, w)
            for r, g in enumerate(geos):
                run(g.unpack_kernel, g.recv_dims, w[r], b[r], g.recv_nd)
        elif g0.is_unpack_free:    # in -> aux pack, aux -> out exchange
            for r, g in enumerate(geos):
                run(g.pack_kernel, g.send_dims, a[r], w[r], g.send_nd)
            _exchange(geos, w, b)
        else:                      # in -> aux pack, aux -> in exchange, in -> out unpack
            for r, g in enumerate(geos):
                run(g.pack_kernel, g.send_dims, a[r], w[r], g.send_nd)
            _exchange(geos, w, a)
            for r, g in enumerate(geos):
                run(g.unpack_kernel, g.recv_dims, a[r], b[r], g.recv_nd)
    elif g0.is_pack_free:          # in -> aux exchange, aux -> out unpack
        _exchange(geos, a, w)
        for r, g in enumerate(geos):
            run(g.unpack_kernel, g.recv_dims, w[r], b[r], g.recv_nd)
    elif g0.is_unpack_free:        # in -> aux pack, aux -> out exchange
        for r, g in enumerate(geos):
            run(g.pack_kernel, g.send_dims, a[r], w[r], g.send_nd)
        _exchange(geos, w, b)
    else:                          # in -> out pack, out -> in exchange, in -> out unpack
        for r, g in enumerate(geos):
            run(g.pack_kernel, g.send_dims, a[r], b[r], g.send_nd)
        _exchange(geos, b, a)
        for r, g in enumerate(geos):
            run(g.unpack_kernel, g.recv_dims, a[r], b[r], g.recv_nd)
    return [b[r][: out_sizes[r]] for r in range(n)]


def apply_local_boxes(src, dst, boxes):
    """Emulate one all-peer launch of the product's pack / unpack kernel over explicit boxes
    (dtfftb_plan_describe_reshape): box = (n0 n1 n2 in_off out_off is1 is2 os0 os1 os2)."""
    apply_boxes(src, [dst] * len(boxes), boxes, list(range(len(boxes))))
