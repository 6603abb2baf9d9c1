"""Python mirror of the reference's plan API (src/interfaces/python/__init__.py:138-848:
``Pencil``, ``Config``, ``PlanC2C`` / ``PlanR2C`` / ``PlanR2R`` with ``execute``, ``transpose``,
``reshape``, ``get_pencil``, ``mem_alloc`` ...) over the C ABI of ``libdtfft_b200.so``
(include/dtfft_b200_api.h).  Same names, argument meaning and error behaviour; the differences
are those of the C header: the communicator is a :class:`dtfft_b200.comm.TorchComm` (or
``None`` for one rank) and buffers are CUDA device memory (``torch`` CUDA tensors or raw device
pointers).  There is no CPU path: host buffers raise.
"""
from __future__ import annotations

import ctypes as C
from enum import IntEnum

from . import _lib
from .comm import as_comm_pointer
from .kernel import _ptr


class Execute(IntEnum):
    FORWARD = 11
    BACKWARD = 12


class Transpose(IntEnum):
    X_TO_Y = 1
    Y_TO_X = -1
    Y_TO_Z = 2
    Z_TO_Y = -2
    X_TO_Z = 3
    Z_TO_X = -3


class Reshape(IntEnum):
    X_BRICKS_TO_PENCILS = 11
    X_PENCILS_TO_BRICKS = 12
    Z_PENCILS_TO_BRICKS = 13
    Z_BRICKS_TO_PENCILS = 14


class Precision(IntEnum):
    SINGLE = 0
    DOUBLE = 1


class Effort(IntEnum):
    ESTIMATE = 0
    MEASURE = 1
    PATIENT = 2
    EXHAUSTIVE = 3


class Executor(IntEnum):
    NONE = 0
    FFTW3 = 1
    MKL = 2
    CUFFT = 3
    VKFFT = 4


class R2RKind(IntEnum):
    DCT_1 = 3
    DCT_2 = 5
    DCT_3 = 4
    DCT_4 = 6
    DST_1 = 7
    DST_2 = 9
    DST_3 = 8
    DST_4 = 10


class Layout(IntEnum):
    X_BRICKS = 1
    X_PENCILS = 2
    X_PENCILS_FOURIER = 3
    Y_PENCILS = 4
    Z_PENCILS = 5
    Z_BRICKS = 6


class Backend(IntEnum):
    MPI_DATATYPE = 21
    MPI_P2P = 22
    MPI_A2A = 23
    NCCL = 24
    CUFFTMP = 25
    MPI_P2P_PIPELINED = 26
    NCCL_PIPELINED = 27
    CUFFTMP_PIPELINED = 28
    MPI_RMA = 29
    MPI_RMA_PIPELINED = 30
    MPI_P2P_SCHEDULED = 31
    MPI_P2P_FUSED = 32
    MPI_RMA_FUSED = 33
    MPI_P2P_COMPRESSED = 34
    MPI_RMA_COMPRESSED = 35
    ADAPTIVE = 36
    NCCL_COMPRESSED = 37
    NVLINK_FUSED = 38
    NONE = -111


class Platform(IntEnum):
    HOST = 1
    CUDA = 2


class TransposeMode(IntEnum):
    PACK = 15
    UNPACK = 16


class AccessMode(IntEnum):
    WRITE = -1
    READ = 1


# Build queries of the reference module (src/interfaces/python/__init__.py:25-72).  This library is the
# CUDA + NCCL + cuFFT build of the reshape path and nothing else.
def is_fftw_enabled() -> bool:
    return False


def is_mkl_enabled() -> bool:
    return False


def is_cufft_enabled() -> bool:
    return True


def is_vkfft_enabled() -> bool:
    return False


def is_cuda_enabled() -> bool:
    return True


def is_transpose_only_enabled() -> bool:
    return False


def is_nccl_enabled() -> bool:
    return True


def is_nvshmem_enabled() -> bool:
    return False


def is_compression_enabled() -> bool:
    return False


def get_backend_string(backend) -> str:
    """String representation of a ``Backend`` value (``dtfft_get_backend_string``)."""
    s = _lib.lib().dtfft_get_backend_string(int(backend))
    return s.decode() if s else ""


class Version:
    """dtFFT version this library mirrors (``DTFFT_VERSION_*``, include/dtfft_b200_api.h);
    ``Version.get()`` asks the loaded library (``dtfft_get_version``: major * 100000 + minor * 1000 + patch)."""

    MAJOR, MINOR, PATCH = 3, 2, 0

    @staticmethod
    def get() -> int:
        return int(_lib.lib().dtfft_get_version())


class Request:
    """Async request handle (``dtfft_request_t``) returned by ``transpose_start`` / ``reshape_start``
    (src/interfaces/python/__init__.py:360-381)."""

    def __init__(self, handle, kind: str):
        self._handle = int(handle or 0)
        self._kind = kind

    @property
    def handle(self) -> int:
        return self._handle

    @property
    def kind(self) -> str:
        return self._kind

    def __int__(self) -> int:
        return self._handle

    def __repr__(self) -> str:
        return f"Request(kind={self._kind!r}, handle=0x{self._handle:x})"


class DtfftError(RuntimeError):
    """Non-zero ``dtfft_error_t`` (same codes as the reference, include/dtfft_config.h.in:82-151)."""

    def __init__(self, code: int, where: str = ""):
        self.code = int(code)
        L = _lib.lib()
        msg = L.dtfft_get_error_string(self.code)
        super().__init__(f"dtFFT error {self.code}{' in ' + where if where else ''}: {msg.decode() if msg else '?'}")


def _check(code: int, where: str = ""):
    if code != 0:
        raise DtfftError(code, where)


class PencilStruct(C.Structure):
    """``dtfft_pencil_t`` (include/dtfft.h:364-380)."""

    _fields_ = [("dim", C.c_uint8), ("ndims", C.c_uint8), ("starts", C.c_int32 * 3), ("counts", C.c_int32 * 3),
                ("size", C.c_size_t)]


class ConfigStruct(C.Structure):
    """``dtfft_config_t`` (include/dtfft.h:1159-1389, CUDA build)."""

    _fields_ = [("enable_log", C.c_bool), ("enable_z_slab", C.c_bool), ("enable_y_slab", C.c_bool),
                ("n_measure_warmup_iters", C.c_int32), ("n_measure_iters", C.c_int32), ("platform", C.c_int),
                ("stream", C.c_void_p), ("backend", C.c_int), ("reshape_backend", C.c_int),
                ("enable_datatype_backend", C.c_bool), ("enable_mpi_backends", C.c_bool),
                ("enable_pipelined_backends", C.c_bool), ("enable_rma_backends", C.c_bool),
                ("enable_fused_backends", C.c_bool), ("enable_nccl_backends", C.c_bool),
                ("enable_nvshmem_backends", C.c_bool), ("enable_kernel_autotune", C.c_bool),
                ("enable_fourier_reshape", C.c_bool), ("transpose_mode", C.c_int), ("access_mode", C.c_int)]


class Pencil:
    """User-described local box (``dtfft_pencil_t``): ``starts`` / ``counts`` in natural
    Fortran order (x fastest).  Reference: src/interfaces/python/__init__.py:138-183."""

    def __init__(self, starts, counts):
        if len(starts) != len(counts):
            raise DtfftError(25, "Pencil")
        self.starts, self.counts = [int(s) for s in starts], [int(c) for c in counts]
        self.ndims = len(self.starts)
        self.dim = 0
        self.size = 1
        for c in self.counts:
            self.size *= c

    def _struct(self) -> PencilStruct:
        p = PencilStruct()
        p.dim, p.ndims = 0, self.ndims
        for i in range(min(3, self.ndims)):
            p.starts[i], p.counts[i] = self.starts[i], self.counts[i]
        p.size = self.size
        return p

    @classmethod
    def _from_struct(cls, s: PencilStruct) -> "Pencil":
        p = cls(list(s.starts[: s.ndims]), list(s.counts[: s.ndims]))
        p.dim, p.size = int(s.dim), int(s.size)
        return p

    def __repr__(self):
        return f"Pencil(dim={self.dim}, starts={self.starts}, counts={self.counts}, size={self.size})"


class Config:
    """``dtfft_config_t`` with the reference's field names; applied with ``dtfft_set_config``
    when passed to a plan constructor (src/interfaces/python/__init__.py:185-301)."""

    def __init__(self, **kwargs):
        self._s = ConfigStruct()
        _check(_lib.lib().dtfft_create_config(C.byref(self._s)), "dtfft_create_config")
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if name == "_s":
            return object.__setattr__(self, name, value)
        if name not in dict(ConfigStruct._fields_):
            raise AttributeError(f"dtfft_config_t has no field '{name}'")
        if name == "stream" and value is not None and not isinstance(value, int):
            value = int(value.cuda_stream)
        setattr(self._s, name, int(value) if isinstance(value, IntEnum) else value)

    def __getattr__(self, name):
        return getattr(self._s, name)

    def _commit(self):
        _check(_lib.lib().dtfft_set_config(C.byref(self._s)), "dtfft_set_config")


class _DeviceBuffer:
    """Device memory from ``dtfft_mem_alloc``; exposes ``__cuda_array_interface__`` so that
    ``torch.as_tensor(buf, device='cuda')`` wraps it without a copy."""

    def __init__(self, plan: "Plan", ptr: int, nbytes: int, dtype_str: str, itemsize: int):
        self._plan, self.ptr, self.nbytes = plan, ptr, nbytes
        self._typestr, self._itemsize = dtype_str, itemsize

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.nbytes // self._itemsize,), "typestr": self._typestr, "data": (self.ptr, False),
                "version": 3, "strides": None}

    def data_ptr(self):
        return self.ptr

    @property
    def is_cuda(self):
        return True

    def free(self):
        if self.ptr and not self._plan._destroyed:
            self._plan.mem_free(self)
        self.ptr = 0

    _owned = False  # set by Plan.get_ndarray: the array owns the allocation

    def __del__(self):
        if self._owned:
            try:
                self.free()
            except Exception:
                pass


def _shaped_view(flat, shape, order):
    """View of the first prod(shape) elements of a flat tensor as ``shape`` in C or Fortran order."""
    n = 1
    for v in shape:
        n *= v
    flat = flat[:n]
    if order == "C":
        return flat.view(shape)
    return flat.view(shape[::-1]).permute(*range(len(shape) - 1, -1, -1))


class Plan:
    """Base plan wrapper (reference: ``class Plan``, src/interfaces/python/__init__.py:396-720)."""

    _KIND = None

    def __init__(self, dims_or_pencil, comm=None, precision=Precision.DOUBLE, effort=Effort.ESTIMATE,
                 executor=Executor.NONE, config: Config | None = None, kinds=None, dry: bool = False):
        L = _lib.lib()
        if config is not None:
            config._commit()
        self._h = C.c_void_p(0)
        self._destroyed = True
        comm_ptr, self._comm_keep = as_comm_pointer(comm)
        is_pencil = isinstance(dims_or_pencil, Pencil)
        if dry:  # host metadata only (dtfftb_plan_create_dry): for CPU-side tests of the plan logic
            kind = {"c2c": 0, "r2c": 1, "r2r": 2}[self._KIND]
            if is_pencil:
                ps = dims_or_pencil._struct()
                rc = L.dtfftb_plan_create_dry(kind, 0, None, C.byref(ps), comm_ptr, int(precision), int(executor),
                                              C.byref(self._h))
            else:
                dims = [int(d) for d in dims_or_pencil]
                rc = L.dtfftb_plan_create_dry(kind, len(dims), (C.c_int32 * len(dims))(*dims), None, comm_ptr,
                                              int(precision), int(executor), C.byref(self._h))
            _check(rc, "dtfftb_plan_create_dry")
            self._destroyed = False
            return
        kinds_arr = None
        if kinds:
            kinds_arr = (C.c_int * len(kinds))(*[int(k) for k in kinds])
        if is_pencil:
            ps = dims_or_pencil._struct()
            args = [C.byref(ps)]
        else:
            dims = [int(d) for d in dims_or_pencil]
            args = [C.c_int8(len(dims)), (C.c_int32 * len(dims))(*dims)]
        if self._KIND == "r2r":
            args.append(kinds_arr)
        args += [comm_ptr, int(precision), int(effort), int(executor), C.byref(self._h)]
        fn = getattr(L, f"dtfft_create_plan_{self._KIND}{'_pencil' if is_pencil else ''}")
        _check(fn(*args), fn.__name__)
        self._destroyed = False

    # ---- execution ----------------------------------------------------------------------
    def execute(self, inbuf, outbuf, execute_type: Execute, aux=None):
        _check(_lib.lib().dtfft_execute(self._h, _ptr(inbuf), _ptr(outbuf), int(execute_type), _ptr(aux) or None),
               "dtfft_execute")

    def transpose(self, inbuf, outbuf, transpose_type: Transpose, aux=None):
        _check(_lib.lib().dtfft_transpose(self._h, _ptr(inbuf), _ptr(outbuf), int(transpose_type), _ptr(aux) or None),
               "dtfft_transpose")

    def reshape(self, inbuf, outbuf, reshape_type: Reshape, aux=None):
        _check(_lib.lib().dtfft_reshape(self._h, _ptr(inbuf), _ptr(outbuf), int(reshape_type), _ptr(aux) or None),
               "dtfft_reshape")

    def transpose_start(self, inbuf, outbuf, transpose_type: Transpose, aux=None):
        req = C.c_void_p(0)
        _check(_lib.lib().dtfft_transpose_start(self._h, _ptr(inbuf), _ptr(outbuf), int(transpose_type),
                                                _ptr(aux) or None, C.byref(req)), "dtfft_transpose_start")
        return Request(req.value, "Transpose." + Transpose(int(transpose_type)).name)

    def transpose_end(self, request):
        handle = request.handle if isinstance(request, Request) else int(getattr(request, "value", request) or 0)
        _check(_lib.lib().dtfft_transpose_end(self._h, C.c_void_p(handle)), "dtfft_transpose_end")

    def reshape_start(self, inbuf, outbuf, reshape_type: Reshape, aux=None):
        req = C.c_void_p(0)
        _check(_lib.lib().dtfft_reshape_start(self._h, _ptr(inbuf), _ptr(outbuf), int(reshape_type), _ptr(aux) or None,
                                              C.byref(req)), "dtfft_reshape_start")
        return Request(req.value, "Reshape." + Reshape(int(reshape_type)).name)

    def reshape_end(self, request):
        handle = request.handle if isinstance(request, Request) else int(getattr(request, "value", request) or 0)
        _check(_lib.lib().dtfft_reshape_end(self._h, C.c_void_p(handle)), "dtfft_reshape_end")

    # ---- metadata -------------------------------------------------------------------------
    def _size_t(self, name) -> int:
        v = C.c_size_t(0)
        _check(getattr(_lib.lib(), name)(self._h, C.byref(v)), name)
        return int(v.value)

    alloc_size = property(lambda s: s._size_t("dtfft_get_alloc_size"))
    alloc_bytes = property(lambda s: s._size_t("dtfft_get_alloc_bytes"))
    element_size = property(lambda s: s._size_t("dtfft_get_element_size"))
    aux_size = property(lambda s: s._size_t("dtfft_get_aux_size"))
    aux_bytes = property(lambda s: s._size_t("dtfft_get_aux_bytes"))
    aux_size_transpose = property(lambda s: s._size_t("dtfft_get_aux_size_transpose"))
    aux_bytes_transpose = property(lambda s: s._size_t("dtfft_get_aux_bytes_transpose"))
    aux_size_reshape = property(lambda s: s._size_t("dtfft_get_aux_size_reshape"))
    aux_bytes_reshape = property(lambda s: s._size_t("dtfft_get_aux_bytes_reshape"))

    @property
    def local_sizes(self):
        """``(in_starts, in_counts, out_starts, out_counts, alloc_size)``."""
        a = [(C.c_int32 * 3)() for _ in range(4)]
        n = C.c_size_t(0)
        _check(_lib.lib().dtfft_get_local_sizes(self._h, a[0], a[1], a[2], a[3], C.byref(n)), "dtfft_get_local_sizes")
        nd = len(self.dims)
        return tuple(list(x[:nd]) for x in a) + (int(n.value),)

    def _int_array(self, name):
        nd = C.c_int8(0)
        p = C.POINTER(C.c_int32)()
        _check(getattr(_lib.lib(), name)(self._h, C.byref(nd), C.byref(p)), name)
        return [int(p[i]) for i in range(nd.value)]

    dims = property(lambda s: s._int_array("dtfft_get_dims"))
    grid_dims = property(lambda s: s._int_array("dtfft_get_grid_dims"))

    def _bool(self, name):
        v = C.c_bool(False)
        _check(getattr(_lib.lib(), name)(self._h, C.byref(v)), name)
        return bool(v.value)

    z_slab_enabled = property(lambda s: s._bool("dtfft_get_z_slab_enabled"))
    y_slab_enabled = property(lambda s: s._bool("dtfft_get_y_slab_enabled"))

    def _enum(self, name, cls):
        v = C.c_int(0)
        _check(getattr(_lib.lib(), name)(self._h, C.byref(v)), name)
        return cls(v.value)

    executor = property(lambda s: s._enum("dtfft_get_executor", Executor))
    precision = property(lambda s: s._enum("dtfft_get_precision", Precision))
    backend = property(lambda s: s._enum("dtfft_get_backend", Backend))
    reshape_backend = property(lambda s: s._enum("dtfft_get_reshape_backend", Backend))
    platform = property(lambda s: s._enum("dtfft_get_platform", Platform))

    @property
    def stream(self) -> int:
        """Raw ``cudaStream_t`` the plan enqueues on (wrap with ``torch.cuda.ExternalStream``)."""
        v = C.c_void_p(0)
        _check(_lib.lib().dtfft_get_stream(self._h, C.byref(v)), "dtfft_get_stream")
        return int(v.value or 0)

    def get_pencil(self, layout: Layout) -> Pencil:
        ps = PencilStruct()
        _check(_lib.lib().dtfft_get_pencil(self._h, int(layout), C.byref(ps)), "dtfft_get_pencil")
        return Pencil._from_struct(ps)

    def report(self):
        _check(_lib.lib().dtfft_report(self._h), "dtfft_report")

    # ---- memory -----------------------------------------------------------------------------
    def mem_alloc(self, alloc_bytes: int, typestr: str = "|u1", itemsize: int = 1) -> _DeviceBuffer:
        p = C.c_void_p(0)
        _check(_lib.lib().dtfft_mem_alloc(self._h, C.c_size_t(int(alloc_bytes)), C.byref(p)), "dtfft_mem_alloc")
        return _DeviceBuffer(self, int(p.value), int(alloc_bytes), typestr, itemsize)

    def mem_free(self, buf):
        ptr = buf.ptr if isinstance(buf, _DeviceBuffer) else _ptr(buf)
        _check(_lib.lib().dtfft_mem_free(self._h, ptr), "dtfft_mem_free")
        if isinstance(buf, _DeviceBuffer):
            buf.ptr = 0

    @property
    def dtype(self):
        """Plan-native element type (src/interfaces/python/__init__.py:472-476): complex for C2C plans,
        real otherwise, width from ``element_size``."""
        import numpy as np

        es = self.element_size
        if self._KIND == "c2c":
            return np.dtype(np.complex128 if es == 16 else np.complex64)
        return np.dtype(np.float64 if es == 8 else np.float32)

    def get_ndarray(self, size: int, shape=None, dtype=None, order: str = "C"):
        """Array backed by plan-managed device memory (``dtfft_mem_alloc``), freed when the last view of it
        dies -- the reference's ``Plan.get_ndarray`` (src/interfaces/python/__init__.py:656-706) with a CUDA
        ``torch.Tensor`` in place of the ``cupy.ndarray`` (there is no host platform here).  ``size`` elements
        are allocated; ``shape`` may cover fewer.  ``order='F'`` gives the column-major view dtFFT's layouts
        are described in (``shape[0]`` fastest)."""
        import math

        import numpy as np
        import torch

        if shape is None:
            shape = (int(size),)
        elif isinstance(shape, int):
            shape = (shape,)
        shape = tuple(int(v) for v in shape)
        if math.prod(shape) > size:
            raise ValueError(f"Shape {shape} is too large for requested size {size}")
        if order not in ("C", "F"):
            raise ValueError("order must be 'C' or 'F'")
        np_dtype = np.dtype(self.dtype if dtype is None else dtype)
        tdtype = {"float32": torch.float32, "float64": torch.float64, "complex64": torch.complex64,
                  "complex128": torch.complex128, "uint8": torch.uint8, "int32": torch.int32,
                  "int64": torch.int64}[np_dtype.name]
        buf = self.mem_alloc(int(size) * np_dtype.itemsize)
        buf._owned = True  # freed by __del__ of the buffer object, which the tensor's storage keeps alive
        return _shaped_view(torch.as_tensor(buf, device="cuda").view(tdtype), shape, order)

    def register_buffer(self, buf, nbytes: int | None = None):
        """NVLINK_FUSED backend: make a user-allocated device buffer reachable by the peers (collective)."""
        if nbytes is None:
            nbytes = buf.numel() * buf.element_size()
        _check(_lib.lib().dtfftb_plan_register_buffer(self._h, _ptr(buf), C.c_size_t(int(nbytes))),
               "dtfftb_plan_register_buffer")

    def unregister_buffer(self, buf):
        _check(_lib.lib().dtfftb_plan_unregister_buffer(self._h, _ptr(buf)), "dtfftb_plan_unregister_buffer")

    def stats(self) -> dict:
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        _check(_lib.lib().dtfftb_plan_get_stats(self._h, C.byref(a), C.byref(b), C.byref(c)), "dtfftb_plan_get_stats")
        return {"kernel_launches": a.value, "local_bytes": b.value, "remote_bytes": c.value}

    def set_overlap(self, nchunks: int, exchange_ctas: int = 0):
        """Stage overlap of cuFFT with the NVLINK_FUSED exchange inside execute (extension, see
        dtfftb_plan_set_overlap): nchunks <= 1 disables.  Same value on every rank."""
        _check(_lib.lib().dtfftb_plan_set_overlap(self._h, int(nchunks), int(exchange_ctas)), "dtfftb_plan_set_overlap")

    @property
    def overlap_chunks(self) -> int:
        n = C.c_int(0)
        _check(_lib.lib().dtfftb_plan_get_overlap(self._h, C.byref(n)), "dtfftb_plan_get_overlap")
        return n.value

    def set_graphs(self, enable: bool):
        """CUDA-graph replay of execute (extension, see dtfftb_plan_set_graphs)."""
        _check(_lib.lib().dtfftb_plan_set_graphs(self._h, int(bool(enable))), "dtfftb_plan_set_graphs")

    def exchange_form(self, transpose_type) -> dict:
        """How one transposition moves its data on this rank (dtfftb_plan_get_exchange_form)."""
        f, n = C.c_int(0), C.c_int(0)
        _check(_lib.lib().dtfftb_plan_get_exchange_form(self._h, int(transpose_type), C.byref(f), C.byref(n)),
               "dtfftb_plan_get_exchange_form")
        names = {0: "local kernel", 1: "nccl", 2: "direct-store kernel", 3: "copy engines",
                 4: "direct-store kernel alone, copy engines in pair pipelines"}
        return {"form": names.get(f.value, str(f.value)), "copies_per_execute": n.value}

    @property
    def fallbacks(self) -> int:
        """NVLINK_FUSED: transpositions / reshapes that ran on the NCCL stand-in (buffer not shareable over cudaIpc)."""
        n = C.c_int64(0)
        _check(_lib.lib().dtfftb_plan_get_fallbacks(self._h, C.byref(n)), "dtfftb_plan_get_fallbacks")
        return n.value

    @property
    def graph_replays(self) -> int:
        n = C.c_int64(0)
        _check(_lib.lib().dtfftb_plan_get_graph_replays(self._h, C.byref(n)), "dtfftb_plan_get_graph_replays")
        return n.value

    @property
    def overlapped_stages(self) -> int:
        n = C.c_int64(0)
        _check(_lib.lib().dtfftb_plan_get_overlapped_stages(self._h, C.byref(n)), "dtfftb_plan_get_overlapped_stages")
        return n.value

    def describe_exchange(self, type_) -> dict:
        """Exchange geometry of one transposition / reshape on this rank (dtfftb_plan_describe_exchange)."""
        import numpy as np

        L = _lib.lib()
        n, me = C.c_int32(0), C.c_int32(0)
        _check(L.dtfftb_plan_describe_exchange(self._h, int(type_), 0, C.byref(n), C.byref(me), None, None, None, None,
                                               None, None, None), "dtfftb_plan_describe_exchange")
        P = n.value
        members = np.zeros(P, np.int32)
        kernels = np.zeros(2, np.int32)
        send_nd, recv_nd = np.zeros((P, 5), np.int32), np.zeros((P, 5), np.int32)
        cd = np.zeros((4, P), np.int64)
        boxes = np.zeros((P, 10), np.int64)
        tr = C.c_int32(0)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        _check(L.dtfftb_plan_describe_exchange(self._h, int(type_), P, C.byref(n), C.byref(me), ip(members), ip(kernels),
                                               ip(send_nd), ip(recv_nd), lp(cd), lp(boxes), C.byref(tr)),
               "dtfftb_plan_describe_exchange")
        return {"members": members.tolist(), "me": me.value, "pack_kernel": int(kernels[0]),
                "unpack_kernel": int(kernels[1]), "send_nd": send_nd, "recv_nd": recv_nd, "send_counts": cd[0],
                "send_displs": cd[1], "recv_counts": cd[2], "recv_displs": cd[3], "fused_boxes": boxes,
                "fused_transposing": bool(tr.value)}

    def describe_chunk(self, transpose_type, k: int, nchunks: int) -> dict:
        """Boxes of chunk ``k`` of ``nchunks`` of a stage-overlapped fused transposition (dtfftb_plan_describe_chunk)."""
        import numpy as np

        L = _lib.lib()
        n, off = C.c_int32(0), C.c_int64(0)
        _check(L.dtfftb_plan_describe_chunk(self._h, int(transpose_type), int(k), int(nchunks), 0, C.byref(n), None,
                                            C.byref(off)), "dtfftb_plan_describe_chunk")
        boxes = np.zeros((n.value, 10), np.int64)
        _check(L.dtfftb_plan_describe_chunk(self._h, int(transpose_type), int(k), int(nchunks), n.value, C.byref(n),
                                            boxes.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(off)),
               "dtfftb_plan_describe_chunk")
        return {"boxes": boxes, "chunk_offset": int(off.value)}

    def describe_dma(self, ttype) -> dict:
        """Copy-engine form of one transposition on this rank (dtfftb_plan_describe_dma): one entry per (member, slice)
        with the pack box, the strided 3-D copy and the direct-store box of the same slice."""
        import numpy as np

        L = _lib.lib()
        n, me, ne = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        _check(L.dtfftb_plan_describe_dma(self._h, int(ttype), 0, 0, C.byref(n), C.byref(me), None, C.byref(ne), None),
               "dtfftb_plan_describe_dma")
        members = (C.c_int32 * n.value)()
        rows = np.zeros((ne.value, 30), np.int64)
        _check(L.dtfftb_plan_describe_dma(self._h, int(ttype), n.value, ne.value, C.byref(n), C.byref(me), members, C.byref(ne),
                                          rows.ctypes.data_as(C.POINTER(C.c_int64))), "dtfftb_plan_describe_dma")
        keys = ("run", "rows", "planes", "dst_off", "dst_pitch", "dst_plane_rows", "ok")
        entries = [{"member": int(r[27]), "sub": int(r[28]), "nsub": int(r[29]), "pack": r[:10].copy(), "fused": r[17:27].copy(),
                    "copy": dict(zip(keys, (int(v) for v in r[10:17])))} for r in rows]
        return {"members": list(members), "me": me.value, "entries": entries}

    def describe_peer_piece(self, t_local, t_exchange, side: int, peer: int, sub: int = 0):
        """The piece of a local transposition cut by slice ``sub`` of the block exchanged with member ``peer`` of the
        transposition next to it: (box[1, 10], slices of that block)."""
        import numpy as np

        box = np.zeros((1, 10), np.int64)
        nsub = C.c_int32(1)
        _check(_lib.lib().dtfftb_plan_describe_peer_piece(self._h, int(t_local), int(t_exchange), int(side), int(peer), int(sub),
                                                          C.byref(nsub), box.ctypes.data_as(C.POINTER(C.c_int64))),
               "dtfftb_plan_describe_peer_piece")
        return box, nsub.value

    def describe_reshape(self, type_) -> dict:
        """NCCL-path geometry of one brick <-> pencil reshape on this rank (dtfftb_plan_describe_reshape)."""
        import numpy as np

        L = _lib.lib()
        n, me = C.c_int32(0), C.c_int32(0)
        _check(L.dtfftb_plan_describe_reshape(self._h, int(type_), 0, C.byref(n), C.byref(me), None, None, None, None,
                                              None), "dtfftb_plan_describe_reshape")
        P = n.value
        members, flags = np.zeros(P, np.int32), np.zeros(3, np.int32)
        pack, unpack = np.zeros((P, 10), np.int64), np.zeros((P, 10), np.int64)
        cd = np.zeros((4, P), np.int64)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        _check(L.dtfftb_plan_describe_reshape(self._h, int(type_), P, C.byref(n), C.byref(me), ip(members), lp(pack),
                                              lp(unpack), lp(cd), ip(flags)), "dtfftb_plan_describe_reshape")
        return {"members": members.tolist(), "me": me.value, "pack_boxes": pack, "unpack_boxes": unpack,
                "send_counts": cd[0], "send_displs": cd[1], "recv_counts": cd[2], "recv_displs": cd[3],
                "is_pack_free": bool(flags[0]), "is_unpack_free": bool(flags[1]), "reshape_strat": int(flags[2])}

    def peer_error(self) -> int:
        return int(_lib.lib().dtfftb_plan_peer_error(self._h))

    def destroy(self):
        if not self._destroyed and self._h:
            _lib.lib().dtfft_destroy(C.byref(self._h))
        self._destroyed = True

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class PlanC2C(Plan):
    """Complex-to-complex plan (src/interfaces/python/__init__.py:722-761)."""

    _KIND = "c2c"


class PlanR2C(Plan):
    """Real-to-complex plan (src/interfaces/python/__init__.py:763-802)."""

    _KIND = "r2c"


class PlanR2R(Plan):
    """Real-to-real plan (src/interfaces/python/__init__.py:804-848)."""

    _KIND = "r2r"

    def __init__(self, dims_or_pencil, kinds=None, comm=None, precision=Precision.DOUBLE, effort=Effort.ESTIMATE,
                 executor=Executor.NONE, config: Config | None = None, dry: bool = False):
        super().__init__(dims_or_pencil, comm, precision, effort, executor, config, kinds=kinds, dry=dry)
