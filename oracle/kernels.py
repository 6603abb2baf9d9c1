"""Oracle (test infrastructure): index maps of every dtFFT reshape kernel kind.

Restates, 0-based and on flat column-major buffers, the loops of
``src/include/_dtfft_kernel_host_routines.inc`` (the reference's CPU kernels, which
are also the bodies of its mock-GPU build) and the index strings the reference emits
for the device in ``src/dtfft_nvrtc_module.F90:494-578``.  Data is opaque: no
arithmetic is performed, so every result is bit-exact by construction.

Two independent formulations are kept on purpose:

* :func:`execute` -- literal ``out[oidx] = in[iidx]`` gather/scatter from the
  reference's index formulas (the "*_write" loops);
* :func:`execute_views` -- the same operation written as numpy reshape/transpose of
  sub-boxes (the *meaning* given in the reference's doc comments, e.g.
  "out(x,y,z) = in(z,x,y)").

``tests/test_oracle_pins.py`` checks the two against each other the way
``src/tests/test_host_kernels.F90`` checks the reference's read/write/block variants
against each other.
"""
from __future__ import annotations

import numpy as np

# kernel_type_t values: src/dtfft_abstract_kernel.F90:59-98
KERNEL_COMPRESSION = -2
KERNEL_DUMMY = -1
KERNEL_PACK = 1
KERNEL_COPY_PIPELINED = 2
KERNEL_UNPACK = 3
KERNEL_COPY = 4
KERNEL_UNPACK_PIPELINED = 5
KERNEL_PACK_PIPELINED = 6
KERNEL_PERMUTE_FORWARD = 7
KERNEL_PERMUTE_BACKWARD = 8
KERNEL_PERMUTE_BACKWARD_START = 9
KERNEL_PERMUTE_BACKWARD_END = 10
KERNEL_PERMUTE_BACKWARD_END_PIPELINED = 11
KERNEL_PACK_FORWARD = 12
KERNEL_PACK_BACKWARD = 13
KERNEL_UNPACK_FORWARD = 15
KERNEL_UNPACK_FORWARD_PIPELINED = 16
KERNEL_UNPACK_BACKWARD = 17
KERNEL_UNPACK_BACKWARD_PIPELINED = 18

KERNEL_NAMES = {
    KERNEL_DUMMY: "dummy", KERNEL_PACK: "pack", KERNEL_COPY_PIPELINED: "copy_pipelined",
    KERNEL_UNPACK: "unpack", KERNEL_COPY: "copy", KERNEL_UNPACK_PIPELINED: "unpack_pipelined",
    KERNEL_PACK_PIPELINED: "pack_pipelined", KERNEL_PERMUTE_FORWARD: "forward",
    KERNEL_PERMUTE_BACKWARD: "backward", KERNEL_PERMUTE_BACKWARD_START: "backward_start",
    KERNEL_PERMUTE_BACKWARD_END: "backward_end",
    KERNEL_PERMUTE_BACKWARD_END_PIPELINED: "backward_end_pipelined",
    KERNEL_PACK_FORWARD: "pack_forward", KERNEL_PACK_BACKWARD: "pack_backward",
    KERNEL_UNPACK_FORWARD: "unpack_forward", KERNEL_UNPACK_FORWARD_PIPELINED: "unpack_forward_pipelined",
    KERNEL_UNPACK_BACKWARD: "unpack_backward", KERNEL_UNPACK_BACKWARD_PIPELINED: "unpack_backward_pipelined",
}

# src/dtfft_abstract_kernel.F90:100-104
TRANSPOSE_KERNELS = (KERNEL_PERMUTE_FORWARD, KERNEL_PERMUTE_BACKWARD, KERNEL_PERMUTE_BACKWARD_START,
                     KERNEL_PACK_FORWARD, KERNEL_PACK_BACKWARD,
                     KERNEL_UNPACK_FORWARD_PIPELINED, KERNEL_UNPACK_BACKWARD_PIPELINED)
UNPACK_KERNELS = (KERNEL_PERMUTE_BACKWARD_END, KERNEL_PERMUTE_BACKWARD_END_PIPELINED, KERNEL_UNPACK,
                  KERNEL_UNPACK_PIPELINED, KERNEL_UNPACK_FORWARD, KERNEL_UNPACK_FORWARD_PIPELINED,
                  KERNEL_UNPACK_BACKWARD, KERNEL_UNPACK_BACKWARD_PIPELINED)
PACK_KERNELS = (KERNEL_PACK, KERNEL_COPY_PIPELINED, KERNEL_PACK_PIPELINED, KERNEL_PACK_FORWARD,
                KERNEL_PACK_BACKWARD)
# kinds that are executed for ONE neighbour per call (need ``neighbor``):
PER_NEIGHBOR_KERNELS = (KERNEL_UNPACK_PIPELINED, KERNEL_PERMUTE_BACKWARD_END_PIPELINED,
                        KERNEL_COPY_PIPELINED, KERNEL_PACK_BACKWARD, KERNEL_PACK_FORWARD,
                        KERNEL_PACK_PIPELINED, KERNEL_UNPACK_FORWARD_PIPELINED,
                        KERNEL_UNPACK_BACKWARD_PIPELINED)
# all-neighbour kinds and the per-neighbour kind they loop over
# (src/include/_dtfft_kernel_host_routines.inc:487-502, 641-652, 726-737, 875-902, 1005-1021)
LOOPED = {
    KERNEL_PACK: KERNEL_PACK_PIPELINED,
    KERNEL_UNPACK: KERNEL_UNPACK_PIPELINED,
    KERNEL_PERMUTE_BACKWARD_END: KERNEL_PERMUTE_BACKWARD_END_PIPELINED,
    KERNEL_UNPACK_FORWARD: KERNEL_UNPACK_FORWARD_PIPELINED,
    KERNEL_UNPACK_BACKWARD: KERNEL_UNPACK_BACKWARD_PIPELINED,
}


def effective_kernel_type(kernel_type: int, ndims: int) -> int:
    """2-D remap of backward kinds to forward ones: src/dtfft_abstract_kernel.F90:271-283."""
    if ndims == 2:
        return {
            KERNEL_PACK_BACKWARD: KERNEL_PACK_FORWARD,
            KERNEL_PERMUTE_BACKWARD: KERNEL_PERMUTE_FORWARD,
            KERNEL_UNPACK_BACKWARD: KERNEL_UNPACK_FORWARD,
            KERNEL_UNPACK_BACKWARD_PIPELINED: KERNEL_UNPACK_FORWARD_PIPELINED,
        }.get(kernel_type, kernel_type)
    return kernel_type


def _dims3(dims):
    dims = [int(d) for d in dims]
    if len(dims) == 2:
        return dims[0], dims[1], 1
    return dims[0], dims[1], dims[2]


def _grid(n1, n2, n3):
    """Index grids x,y,z of a (n1,n2,n3) box, int64, broadcastable."""
    x = np.arange(n1, dtype=np.int64)[:, None, None]
    y = np.arange(n2, dtype=np.int64)[None, :, None]
    z = np.arange(n3, dtype=np.int64)[None, None, :]
    return x, y, z


def index_maps(kernel_type: int, dims, locals5=None):
    """Return ``(iidx, oidx)`` flat 0-based index arrays of one launch of a
    per-neighbour (or whole-buffer permute) kernel.

    ``locals5`` = ``neighbor_data(:, n)`` = (nxx, nyy, nzz, din, dout), elements.
    Formulas: see the per-kind citations below (file
    ``src/include/_dtfft_kernel_host_routines.inc`` unless noted).
    """
    ndims = len(dims)
    nx, ny, nz = _dims3(dims)
    kt = effective_kernel_type(kernel_type, ndims)
    if locals5 is not None:
        nxx, nyy, nzz, din, dout = (int(v) for v in locals5)
        if ndims == 2:
            nzz = 1

    if kt == KERNEL_PERMUTE_FORWARD:
        # :140-196  out(y,z,x) <- in(x,y,z); 2-D: out(y,x) <- in(x,y)
        x, y, z = _grid(nx, ny, nz)
        iidx = x + y * nx + z * nx * ny
        oidx = y + z * ny + x * ny * nz
    elif kt == KERNEL_PERMUTE_BACKWARD:
        # :262-302  out(z,x,y) <- in(x,y,z)
        x, y, z = _grid(nx, ny, nz)
        iidx = x + y * nx + z * nx * ny
        oidx = z + x * nz + y * nz * nx
    elif kt == KERNEL_PERMUTE_BACKWARD_START:
        # :351-390  out(z,y,x) <- in(x,y,z)
        x, y, z = _grid(nx, ny, nz)
        iidx = x + y * nx + z * nx * ny
        oidx = z + y * nz + x * nz * ny
    elif kt == KERNEL_PERMUTE_BACKWARD_END_PIPELINED:
        # :439-482  swap dims 2,3 of the received block while scattering
        x, y, z = _grid(nxx, nyy, nzz)
        iidx = din + x + z * nxx + y * nxx * nzz
        oidx = dout + x + y * nx + z * nx * ny
    elif kt == KERNEL_UNPACK_PIPELINED:
        # :575-636
        x, y, z = _grid(nxx, nyy, nzz)
        iidx = din + x + y * nxx + z * nxx * nyy
        oidx = dout + x + y * nx + z * nx * ny
    elif kt == KERNEL_PACK_PIPELINED:
        # :657-721
        x, y, z = _grid(nxx, nyy, nzz)
        iidx = din + x + y * nx + z * nx * ny
        oidx = dout + x + y * nxx + z * nxx * nyy
    elif kt == KERNEL_PACK_FORWARD:
        # :1026-1086 forward permute of the peer's sub-box
        x, y, z = _grid(nxx, nyy, nzz)
        iidx = din + x + y * nx + z * nx * ny
        oidx = dout + y + z * nyy + x * nyy * nzz
    elif kt == KERNEL_PACK_BACKWARD:
        # :1155-1198 backward permute of the peer's sub-box
        x, y, z = _grid(nxx, nyy, nzz)
        iidx = din + x + y * nx + z * nx * ny
        oidx = dout + z + x * nzz + y * nzz * nxx
    elif kt == KERNEL_UNPACK_FORWARD_PIPELINED:
        # :757-811 (host only; device rejects it: src/dtfft_kernel_device.F90:72-74)
        x, y, z = _grid(nxx, nyy, nzz)
        if ndims == 2:
            iidx = din + y + x * nyy
            oidx = dout + x + y * nx + 0 * z
        else:
            iidx = din + z + x * nzz + y * nzz * nxx
            oidx = dout + x + y * nx + z * nx * ny
    elif kt == KERNEL_UNPACK_BACKWARD_PIPELINED:
        # :907-946 (host only)
        x, y, z = _grid(nxx, nyy, nzz)
        iidx = din + y + z * nyy + x * nzz * nyy
        oidx = dout + x + y * nx + z * nx * ny
    elif kt == KERNEL_COPY_PIPELINED:
        # :742-752 with starts = neighbor_data(4:5)
        n = nxx * nyy * nzz
        i = np.arange(n, dtype=np.int64)
        return din + i, dout + i
    elif kt == KERNEL_COPY:
        n = nx * ny * nz
        i = np.arange(n, dtype=np.int64)
        return i, i
    else:
        raise ValueError(f"index_maps: unsupported kernel type {kernel_type}")
    shape = np.broadcast_shapes(iidx.shape, oidx.shape)
    return (np.broadcast_to(iidx, shape).reshape(-1), np.broadcast_to(oidx, shape).reshape(-1))


def execute(kernel_type: int, dims, inbuf: np.ndarray, outbuf: np.ndarray,
            neighbor_data=None, neighbor: int | None = None) -> None:
    """Run one kernel of kind ``kernel_type`` on flat buffers (dispatch mirrors
    ``execute`` at ``_dtfft_kernel_host_routines.inc:16-96``).

    ``neighbor_data``: int array (P, 5) -- row ``n`` is the reference's
    ``neighbor_data(:, n+1)``.  ``neighbor`` is **1-based** like the reference.
    Zero-volume ``dims`` make the kernel a no-op (``abstract_kernel.F90:236-240``).
    """
    if any(int(d) == 0 for d in dims) or kernel_type == KERNEL_DUMMY:
        return
    kt = effective_kernel_type(kernel_type, len(dims))
    if kt in LOOPED:
        nd = np.asarray(neighbor_data).reshape(-1, 5)
        for n in range(nd.shape[0]):
            execute(LOOPED[kt], dims, inbuf, outbuf, nd, n + 1)
        return
    locals5 = None
    if kt in PER_NEIGHBOR_KERNELS:
        if neighbor is None:
            raise ValueError("neighbor required")
        nd = np.asarray(neighbor_data).reshape(-1, 5)
        if not (1 <= neighbor <= nd.shape[0]):
            raise ValueError("neighbor out of bounds")
        locals5 = nd[neighbor - 1]
        if int(np.prod(locals5[: 3 if len(dims) == 3 else 2])) == 0:
            return
    iidx, oidx = index_maps(kt, dims, locals5)
    outbuf[oidx] = inbuf[iidx]


def execute_views(kernel_type: int, dims, inbuf: np.ndarray, outbuf: np.ndarray,
                  neighbor_data=None, neighbor: int | None = None) -> None:
    """Second, independent formulation through strided views (no index arithmetic on
    elements): each kind is "take this sub-box of ``in`` seen as a Fortran array,
    permute its axes, drop it into that sub-box of ``out``"."""
    from numpy.lib.stride_tricks import as_strided

    if any(int(d) == 0 for d in dims) or kernel_type == KERNEL_DUMMY:
        return
    ndims = len(dims)
    kt = effective_kernel_type(kernel_type, ndims)
    if kt in LOOPED:
        nd = np.asarray(neighbor_data).reshape(-1, 5)
        for n in range(nd.shape[0]):
            execute_views(LOOPED[kt], dims, inbuf, outbuf, nd, n + 1)
        return
    nx, ny, nz = _dims3(dims)
    es = inbuf.itemsize

    def box(buf, off, shape, strides_elems):
        if int(np.prod(shape)) == 0:
            return None
        return as_strided(buf[off:], shape=shape, strides=tuple(s * es for s in strides_elems))

    if kt in (KERNEL_PERMUTE_FORWARD, KERNEL_PERMUTE_BACKWARD, KERNEL_PERMUTE_BACKWARD_START, KERNEL_COPY):
        src = box(inbuf, 0, (nx, ny, nz), (1, nx, nx * ny))          # src[x,y,z]
        if kt == KERNEL_PERMUTE_FORWARD:
            dst = box(outbuf, 0, (ny, nz, nx), (1, ny, ny * nz))     # dst[y,z,x]
            dst[...] = src.transpose(1, 2, 0)
        elif kt == KERNEL_PERMUTE_BACKWARD:
            dst = box(outbuf, 0, (nz, nx, ny), (1, nz, nz * nx))     # dst[z,x,y]
            dst[...] = src.transpose(2, 0, 1)
        elif kt == KERNEL_PERMUTE_BACKWARD_START:
            dst = box(outbuf, 0, (nz, ny, nx), (1, nz, nz * ny))     # dst[z,y,x]
            dst[...] = src.transpose(2, 1, 0)
        else:
            outbuf[: nx * ny * nz] = inbuf[: nx * ny * nz]
        return

    nd = np.asarray(neighbor_data).reshape(-1, 5)
    nxx, nyy, nzz, din, dout = (int(v) for v in nd[neighbor - 1])
    if ndims == 2:
        nzz = 1
    if nxx * nyy * nzz == 0:
        return
    if kt == KERNEL_COPY_PIPELINED:
        n = nxx * nyy * nzz
        outbuf[dout:dout + n] = inbuf[din:din + n]
        return
    full_in = lambda: box(inbuf, din, (nxx, nyy, nzz), (1, nx, nx * ny))      # sub-box of the full local array
    full_out = lambda: box(outbuf, dout, (nxx, nyy, nzz), (1, nx, nx * ny))
    if kt == KERNEL_PACK_PIPELINED:
        box(outbuf, dout, (nxx, nyy, nzz), (1, nxx, nxx * nyy))[...] = full_in()
    elif kt == KERNEL_UNPACK_PIPELINED:
        full_out()[...] = box(inbuf, din, (nxx, nyy, nzz), (1, nxx, nxx * nyy))
    elif kt == KERNEL_PERMUTE_BACKWARD_END_PIPELINED:
        # contiguous block is stored (x, z, y)
        blk = box(inbuf, din, (nxx, nzz, nyy), (1, nxx, nxx * nzz))
        full_out()[...] = blk.transpose(0, 2, 1)
    elif kt == KERNEL_PACK_FORWARD:
        box(outbuf, dout, (nyy, nzz, nxx), (1, nyy, nyy * nzz))[...] = full_in().transpose(1, 2, 0)
    elif kt == KERNEL_PACK_BACKWARD:
        box(outbuf, dout, (nzz, nxx, nyy), (1, nzz, nzz * nxx))[...] = full_in().transpose(2, 0, 1)
    elif kt == KERNEL_UNPACK_FORWARD_PIPELINED:
        if ndims == 2:
            blk = box(inbuf, din, (nyy, nxx, 1), (1, nyy, nyy * nxx))        # stored (y, x)
            full_out()[...] = blk.transpose(1, 0, 2)
        else:
            blk = box(inbuf, din, (nzz, nxx, nyy), (1, nzz, nzz * nxx))      # stored (z, x, y)
            full_out()[...] = blk.transpose(1, 2, 0)
    elif kt == KERNEL_UNPACK_BACKWARD_PIPELINED:
        blk = box(inbuf, din, (nyy, nzz, nxx), (1, nyy, nyy * nzz))          # stored (y, z, x)
        full_out()[...] = blk.transpose(2, 0, 1)
    else:
        raise ValueError(f"execute_views: unsupported kernel type {kernel_type}")
