"""Python mirrors of the reference's exchange-backend and FFT-executor plugin interfaces over the
C ABI (include/dtfft_b200.h): ``abstract_backend`` / ``backend_nccl``
(src/dtfft_abstract_backend.F90:114-343, src/dtfft_backend_nccl.F90:38-134), ``backend_helper``'s
NCCL communicator (src/dtfft_abstract_backend.F90:395-457) and ``abstract_executor`` /
``cufft_executor`` (src/dtfft_abstract_executor.F90:67-112,
src/interfaces/fft/cufft/dtfft_executor_cufft_m.F90:52-125).  Same vocabulary as the reference:
``create`` / ``execute`` / ``destroy``.  Device buffers only; torch is plumbing."""
from __future__ import annotations

import ctypes as C

from . import _lib
from .comm import as_comm_pointer
from .kernel import Kernel, _ptr, _stream

FFT_C2C, FFT_R2C, FFT_R2R = 0, 1, 2            # src/dtfft_abstract_executor.F90:35-39
FFT_FORWARD, FFT_BACKWARD = -1, 1              # src/include/_dtfft_private.h:27-28
BACKEND_NCCL, BACKEND_NCCL_PIPELINED = 24, 27  # include/dtfft_config.h.in:153-169


def _declare(L):
    if getattr(L, "_dtfftb_plugins_declared", False):
        return L
    vp, pvp = C.c_void_p, C.POINTER(C.c_void_p)
    i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    sig = {
        "dtfftb_nccl_comm_create": [vp, pvp],
        "dtfftb_nccl_comm_destroy": [pvp],
        "dtfftb_backend_create": [pvp, C.c_int, vp, C.c_int, C.c_int, i32p, i64p, i64p, i64p, i64p, C.c_int64],
        "dtfftb_backend_set_unpack_kernel": [vp, vp],
        "dtfftb_backend_get_aux_bytes": [vp, i64p],
        "dtfftb_backend_execute": [vp, vp, vp, vp, vp],
        "dtfftb_backend_destroy": [pvp],
        "dtfftb_executor_create": [pvp, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_int32, i32p, i32p, i32p, vp],
        "dtfftb_executor_execute": [vp, vp, vp, C.c_int],
        "dtfftb_executor_destroy": [pvp],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes, fn.restype = argtypes, C.c_int
    L._dtfftb_plugins_declared = True
    return L


class NcclComm:
    """One NCCL communicator over a process group (``backend_helper%create``); ``comm=None`` = 1 rank."""

    def __init__(self, comm=None):
        self._L = _declare(_lib.lib())
        self._h = C.c_void_p(0)
        ptr, self._keep = as_comm_pointer(comm)
        _lib.check(self._L.dtfftb_nccl_comm_create(ptr, C.byref(self._h)), "dtfftb_nccl_comm_create")

    @property
    def handle(self) -> int:
        return int(self._h.value or 0)

    def destroy(self):
        if self._h:
            self._L.dtfftb_nccl_comm_destroy(C.byref(self._h))
        self._h = C.c_void_p(0)


class ExchangeBackend:
    """``class(abstract_backend)`` with the NCCL implementations of this path."""

    def __init__(self):
        self._L = _declare(_lib.lib())
        self._h = C.c_void_p(0)
        self._unpack = None

    def create(self, backend_type, nccl: NcclComm, comm_rank, send_displs, send_counts, recv_displs, recv_counts,
               base_storage, comm_mapping=None):
        """Counts / displacements per member of the 1-D communicator, in elements, 0-based."""
        self.destroy()
        P = len(send_counts)
        arr = lambda v: (C.c_int64 * P)(*[int(x) for x in v])
        mapping = (C.c_int32 * P)(*[int(x) for x in comm_mapping]) if comm_mapping is not None else None
        _lib.check(self._L.dtfftb_backend_create(C.byref(self._h), int(backend_type), nccl.handle, int(comm_rank), P, mapping,
                                                 arr(send_displs), arr(send_counts), arr(recv_displs), arr(recv_counts),
                                                 int(base_storage)), "dtfftb_backend_create")
        return self

    def set_unpack_kernel(self, kernel: Kernel):
        self._unpack = kernel  # keep alive: the backend only borrows it
        _lib.check(self._L.dtfftb_backend_set_unpack_kernel(self._h, kernel._h), "dtfftb_backend_set_unpack_kernel")

    @property
    def aux_bytes(self) -> int:
        n = C.c_int64(0)
        _lib.check(self._L.dtfftb_backend_get_aux_bytes(self._h, C.byref(n)), "dtfftb_backend_get_aux_bytes")
        return n.value

    def execute(self, inbuf, outbuf, stream=None, aux=None):
        _lib.check(self._L.dtfftb_backend_execute(self._h, _ptr(inbuf), _ptr(outbuf), _stream(stream), _ptr(aux) or None),
                   "dtfftb_backend_execute")

    def destroy(self):
        if self._h:
            self._L.dtfftb_backend_destroy(C.byref(self._h))
        self._h = C.c_void_p(0)


class FftExecutor:
    """``class(abstract_executor)`` backed by cuFFT."""

    def __init__(self):
        self._L = _declare(_lib.lib())
        self._h = C.c_void_p(0)

    def create(self, fft_rank, fft_type, precision, idist, odist, how_many, fft_sizes, inembed, onembed, stream=None):
        """``create_private`` argument for argument (sizes slowest first, unit stride)."""
        self.destroy()
        n = len(fft_sizes)
        arr = lambda v: (C.c_int32 * n)(*[int(x) for x in v])
        _lib.check(self._L.dtfftb_executor_create(C.byref(self._h), int(fft_rank), int(fft_type), int(precision), int(idist),
                                                  int(odist), int(how_many), arr(fft_sizes), arr(inembed), arr(onembed),
                                                  _stream(stream)), "dtfftb_executor_create")
        return self

    def execute(self, a, b, sign):
        _lib.check(self._L.dtfftb_executor_execute(self._h, _ptr(a), _ptr(b), int(sign)), "dtfftb_executor_execute")

    def destroy(self):
        if self._h:
            self._L.dtfftb_executor_destroy(C.byref(self._h))
        self._h = C.c_void_p(0)
