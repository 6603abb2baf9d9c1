"""Python mirror of the reference's kernel plugin (``abstract_kernel`` / ``kernel_device``,
src/dtfft_abstract_kernel.F90:106-163, src/dtfft_kernel_device.F90:45-57) over the C ABI.

Same vocabulary as the reference: ``create(dims, effort, base_storage, kernel_type,
neighbor_data)``, ``execute(in, out, stream, neighbor)``, ``destroy()``.  Buffers are CUDA
device pointers (ints) or ``torch`` CUDA tensors; ``stream`` is a ``cudaStream_t`` value or a
``torch.cuda.Stream``.  PyTorch is plumbing only (device memory, streams).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

# kernel_type_t (src/dtfft_abstract_kernel.F90:59-98)
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

DTFFT_ESTIMATE, DTFFT_MEASURE, DTFFT_PATIENT, DTFFT_EXHAUSTIVE = 0, 1, 2, 3

FAMILY_NAMES = {0: "none", 1: "copy", 2: "transpose", 3: "rows"}


def _ptr(buf) -> int:
    if buf is None:
        return 0
    if isinstance(buf, int):
        return buf
    if hasattr(buf, "data_ptr"):
        if not buf.is_cuda:
            raise TypeError("dtfft_b200 kernels take device buffers only (no CPU fallback)")
        return int(buf.data_ptr())
    iface = getattr(buf, "__cuda_array_interface__", None)  # cupy / numba arrays, like the reference's cupy inputs
    if iface is not None:
        return int(iface["data"][0])
    if hasattr(buf, "__array_interface__"):
        raise TypeError("dtfft_b200 takes device buffers only (no CPU fallback): got a host array")
    raise TypeError(f"unsupported buffer type {type(buf)}")


def _stream(stream) -> int:
    if stream is None:
        return 0
    if isinstance(stream, int):
        return stream
    return int(stream.cuda_stream)


class Kernel:
    """One pack / unpack / permute kernel object (reference: ``class(abstract_kernel)``)."""

    def __init__(self):
        self._h = C.c_void_p(0)
        self.dims = None
        self.kernel_type = None
        self.base_storage = None
        self.neighbor_data = None

    def create(self, dims, effort, base_storage, kernel_type, neighbor_data=None, force_effort=False):
        """``abstract_kernel%create`` (src/dtfft_abstract_kernel.F90:219-288).

        ``neighbor_data``: array-like (P, 5) -- row n is the reference's ``neighbor_data(:, n+1)``."""
        self.destroy()
        L = _lib.lib()
        dims_arr = (C.c_int32 * len(dims))(*[int(d) for d in dims])
        nd_ptr, n_nb = None, 0
        if neighbor_data is not None:
            nd = np.ascontiguousarray(np.asarray(neighbor_data, dtype=np.int32).reshape(-1, 5))
            n_nb = nd.shape[0]
            self._nd_keep = nd
            nd_ptr = nd.ctypes.data_as(C.POINTER(C.c_int32))  # (P,5) C-order == Fortran (5,P)
        _lib.check(L.dtfftb_kernel_create(C.byref(self._h), len(dims), dims_arr, int(kernel_type), int(base_storage),
                                          nd_ptr, n_nb, int(effort), int(bool(force_effort))), "dtfftb_kernel_create")
        self.dims, self.kernel_type, self.base_storage = list(dims), int(kernel_type), int(base_storage)
        self.neighbor_data = None if neighbor_data is None else np.asarray(neighbor_data, dtype=np.int32).reshape(-1, 5)
        return self

    def create_dry(self, dims, base_storage, kernel_type, neighbor_data=None):
        """Host-only twin of :meth:`create` (``dtfftb_kernel_create_dry``): same geometry and work-item
        tables, no device; for CPU tests of the table builder (:meth:`dump_table`)."""
        self.destroy()
        L = _lib.lib()
        dims_arr = (C.c_int32 * len(dims))(*[int(d) for d in dims])
        nd_ptr, n_nb = None, 0
        if neighbor_data is not None:
            nd = np.ascontiguousarray(np.asarray(neighbor_data, dtype=np.int32).reshape(-1, 5))
            n_nb = nd.shape[0]
            self._nd_keep = nd
            nd_ptr = nd.ctypes.data_as(C.POINTER(C.c_int32))
        _lib.check(L.dtfftb_kernel_create_dry(C.byref(self._h), len(dims), dims_arr, int(kernel_type), int(base_storage),
                                              nd_ptr, n_nb), "dtfftb_kernel_create_dry")
        self.dims, self.kernel_type, self.base_storage = list(dims), int(kernel_type), int(base_storage)
        return self

    def create_boxes_dry(self, family: int, base_storage, boxes, remote_peers=False):
        """Host-only kernel over explicit boxes (rows of n0 n1 n2 in_off out_off is1 is2 os0 os1 os2), as
        the plan layer builds for the fused NVLink path (family 2 or 3) and the brick reshapes (family 3)."""
        self.destroy()
        bx = np.ascontiguousarray(np.asarray(boxes, dtype=np.int64).reshape(-1, 10))
        _lib.check(_lib.lib().dtfftb_kernel_create_boxes_dry(C.byref(self._h), int(family), int(base_storage), bx.shape[0],
                                                             bx.ctypes.data_as(C.POINTER(C.c_int64)),
                                                             int(bool(remote_peers))), "dtfftb_kernel_create_boxes_dry")
        self.base_storage = int(base_storage)
        return self

    def create_boxes(self, family: int, base_storage, boxes, out_bases=None):
        """Kernel over explicit boxes on the current device (``dtfftb_kernel_create_boxes``): what the plan's
        direct-store transpositions run.  ``out_bases``: one destination base per box (device pointers or
        tensors, may live on a peer GPU with peer access enabled), or None for the launch's ``out``."""
        self.destroy()
        bx = np.ascontiguousarray(np.asarray(boxes, dtype=np.int64).reshape(-1, 10))
        bases = None
        if out_bases is not None:
            assert len(out_bases) == bx.shape[0]
            bases = (C.c_void_p * bx.shape[0])(*[C.c_void_p(_ptr(b) or None) for b in out_bases])
        _lib.check(_lib.lib().dtfftb_kernel_create_boxes(C.byref(self._h), int(family), int(base_storage), bx.shape[0],
                                                         bx.ctypes.data_as(C.POINTER(C.c_int64)), bases),
                   "dtfftb_kernel_create_boxes")
        self.base_storage = int(base_storage)
        return self

    def dump_table(self, unit=0, neighbor=0) -> dict:
        """Work-item table of one launch of a dry kernel (``dtfftb_kernel_dump_table``)."""
        L = _lib.lib()
        nb, total = C.c_int32(0), C.c_int64(0)
        launch = (C.c_int32 * 3)()
        _lib.check(L.dtfftb_kernel_dump_table(self._h, int(unit), int(neighbor), 0, None, C.byref(nb), C.byref(total),
                                              launch), "dtfftb_kernel_dump_table")
        rows = np.zeros((nb.value, 20), np.int64)
        if nb.value:
            _lib.check(L.dtfftb_kernel_dump_table(self._h, int(unit), int(neighbor), nb.value,
                                                  rows.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(nb), C.byref(total),
                                                  launch), "dtfftb_kernel_dump_table")
        return {"blocks": rows, "total_items": int(total.value), "launch": [int(v) for v in launch]}

    def execute(self, inbuf, outbuf, stream=None, neighbor=None, sync=False):
        """``abstract_kernel%execute`` (src/dtfft_abstract_kernel.F90:290-403). ``neighbor`` is 1-based."""
        _lib.check(_lib.lib().dtfftb_kernel_execute(self._h, _ptr(inbuf), _ptr(outbuf), _stream(stream),
                                                    int(neighbor or 0), int(bool(sync))), "dtfftb_kernel_execute")

    def execute_all(self, inbuf, outbuf, stream=None):
        _lib.check(_lib.lib().dtfftb_kernel_execute_all(self._h, _ptr(inbuf), _ptr(outbuf), _stream(stream)),
                   "dtfftb_kernel_execute_all")

    def set_tile(self, ka, kb, rows):
        _lib.check(_lib.lib().dtfftb_kernel_set_tile(self._h, ka, kb, rows), "dtfftb_kernel_set_tile")

    def autotune(self, inbuf, outbuf, stream=None, n_warmup=2, n_iters=5) -> float:
        ms = C.c_float(0)
        _lib.check(_lib.lib().dtfftb_kernel_autotune(self._h, _ptr(inbuf), _ptr(outbuf), _stream(stream), n_warmup,
                                                     n_iters, C.byref(ms)), "dtfftb_kernel_autotune")
        return float(ms.value)

    def autotune_report(self, inbuf, outbuf, stream=None, n_warmup=2, n_iters=5) -> list:
        """Timed tile autotune with the per-candidate log of the reference
        (src/dtfft_kernel_device.F90:385-389): [{"tile_a", "tile_b", "threads", "ms", "gbs"}, ...]; the
        fastest candidate is kept."""
        cap = 32
        n = C.c_int(0)
        tiles = (C.c_int32 * (3 * cap))()
        ms = (C.c_float * cap)()
        gbs = (C.c_double * cap)()
        _lib.check(_lib.lib().dtfftb_kernel_autotune_report(self._h, _ptr(inbuf), _ptr(outbuf), _stream(stream), n_warmup,
                                                            n_iters, cap, C.byref(n), tiles, ms, gbs),
                   "dtfftb_kernel_autotune_report")
        return [{"tile_a": tiles[3 * i], "tile_b": tiles[3 * i + 1], "threads": tiles[3 * i + 2], "ms": float(ms[i]),
                 "gbs": float(gbs[i])} for i in range(min(n.value, cap))]

    def info(self) -> dict:
        fam, unit, ta, tb, thr = (C.c_int(0) for _ in range(5))
        items = C.c_int64(0)
        _lib.check(_lib.lib().dtfftb_kernel_get_info(self._h, C.byref(fam), C.byref(unit), C.byref(ta), C.byref(tb),
                                                     C.byref(thr), C.byref(items)), "dtfftb_kernel_get_info")
        return {"family": FAMILY_NAMES[fam.value], "unit_bytes": unit.value, "tile_a": ta.value, "tile_b": tb.value,
                "threads": thr.value, "items": items.value}

    def destroy(self):
        if self._h:
            _lib.lib().dtfftb_kernel_destroy(C.byref(self._h))
        self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
