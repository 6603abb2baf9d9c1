"""ctypes loader of ``libdtfft_b200.so`` (the C ABI declared in ``include/dtfft_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C dtfft_b200/csrc``.
There is NO CPU fallback: if the shared object is missing every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdtfft_b200.so")

_lib = None


class DtfftB200Error(RuntimeError):
    """Non-zero return code of the C ABI (dtfft_error_t or DTFFTB_ERROR_*)."""

    def __init__(self, code: int, where: str):
        self.code = int(code)
        super().__init__(f"{where} failed with code {self.code} ({describe_error(self.code)})")


def describe_error(code: int) -> str:
    named = {-30001: "NVLINK_FUSED: `out` is not a registered buffer", -30002: "host allgather callback failed",
             -30003: "NVLINK_FUSED: a peer missed a device barrier (time-out); the plan is dead"}
    if code in named:
        return named[code]
    if code <= -30000:
        return "internal invariant violated"
    if code <= -20000:
        return f"NCCL error {-(code + 20000)}"
    if code <= -10000:
        return f"CUDA error {-(code + 10000)}"
    return "dtfft_error_t"


def _preload_nccl() -> None:
    """Load the NCCL that torch ships (site-packages/nvidia/nccl) BEFORE our library, so that the
    `libnccl.so.2` our .so needs resolves to the very copy torch's own NCCL process group uses.
    Otherwise the dynamic loader would pick the older system libnccl first and a later
    `import torch` in the same process fails with an undefined NCCL symbol."""
    import importlib.util
    import sys

    if "torch" in sys.modules:
        return  # torch already mapped its NCCL
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for d in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(d, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            C.CDLL(cand, mode=C.RTLD_GLOBAL)
            return


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C dtfft_b200/csrc`. dtfft_b200 has no CPU fallback.")
    _preload_nccl()
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32p, i64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    L.dtfftb_version.restype = C.c_char_p
    L.dtfftb_device_available.restype = C.c_int
    L.dtfftb_kernel_create.argtypes = [C.POINTER(vp), C.c_int, i32p, C.c_int, C.c_int64, i32p, C.c_int, C.c_int, C.c_int]
    L.dtfftb_kernel_execute.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int]
    L.dtfftb_kernel_execute_all.argtypes = [vp, vp, vp, vp]
    L.dtfftb_kernel_set_peer_out.argtypes = [vp, C.POINTER(vp), i64p]
    L.dtfftb_kernel_destroy.argtypes = [C.POINTER(vp)]
    L.dtfftb_kernel_get_info.argtypes = [vp] + [C.POINTER(C.c_int)] * 5 + [i64p]
    L.dtfftb_kernel_set_tile.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.dtfftb_kernel_autotune.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.dtfftb_kernel_autotune_report.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), i32p,
                                                C.POINTER(C.c_float), C.POINTER(C.c_double)]
    L.dtfftb_kernel_autotune_report.restype = C.c_int
    L.dtfftb_kernel_create_dry.argtypes = [C.POINTER(vp), C.c_int, i32p, C.c_int, C.c_int64, i32p, C.c_int]
    L.dtfftb_kernel_create_boxes.argtypes = [C.POINTER(vp), C.c_int, C.c_int64, C.c_int, i64p, C.POINTER(vp)]
    L.dtfftb_kernel_create_boxes.restype = C.c_int
    L.dtfftb_kernel_create_boxes_dry.argtypes = [C.POINTER(vp), C.c_int, C.c_int64, C.c_int, i64p, C.c_int]
    L.dtfftb_kernel_dump_table.argtypes = [vp, C.c_int, C.c_int, C.c_int32, i64p, i32p, i64p, i32p]
    for name in ("dtfftb_kernel_create_dry", "dtfftb_kernel_create_boxes_dry", "dtfftb_kernel_dump_table"):
        getattr(L, name).restype = C.c_int
    for name in ("dtfftb_kernel_create", "dtfftb_kernel_execute", "dtfftb_kernel_execute_all",
                 "dtfftb_kernel_set_peer_out", "dtfftb_kernel_destroy", "dtfftb_kernel_get_info",
                 "dtfftb_kernel_set_tile", "dtfftb_kernel_autotune"):
        getattr(L, name).restype = C.c_int
    _declare_plan_api(L)
    _lib = L
    return L


def _declare_plan_api(L):
    """argtypes / restypes of include/dtfft_b200_api.h (pointers must not be truncated to int)."""
    vp, i32p = C.c_void_p, C.POINTER(C.c_int32)
    szp = C.POINTER(C.c_size_t)
    pvp = C.POINTER(vp)
    intp = C.POINTER(C.c_int)
    sig = {
        "dtfft_create_plan_c2c": [C.c_int8, i32p, vp, C.c_int, C.c_int, C.c_int, pvp],
        "dtfft_create_plan_r2c": [C.c_int8, i32p, vp, C.c_int, C.c_int, C.c_int, pvp],
        "dtfft_create_plan_r2r": [C.c_int8, i32p, intp, vp, C.c_int, C.c_int, C.c_int, pvp],
        "dtfft_create_plan_c2c_pencil": [vp, vp, C.c_int, C.c_int, C.c_int, pvp],
        "dtfft_create_plan_r2c_pencil": [vp, vp, C.c_int, C.c_int, C.c_int, pvp],
        "dtfft_create_plan_r2r_pencil": [vp, intp, vp, C.c_int, C.c_int, C.c_int, pvp],
        "dtfft_execute": [vp, vp, vp, C.c_int, vp],
        "dtfft_transpose": [vp, vp, vp, C.c_int, vp],
        "dtfft_reshape": [vp, vp, vp, C.c_int, vp],
        "dtfft_transpose_start": [vp, vp, vp, C.c_int, vp, pvp],
        "dtfft_reshape_start": [vp, vp, vp, C.c_int, vp, pvp],
        "dtfft_transpose_end": [vp, vp],
        "dtfft_reshape_end": [vp, vp],
        "dtfft_destroy": [pvp],
        "dtfft_get_local_sizes": [vp, i32p, i32p, i32p, i32p, szp],
        "dtfft_get_pencil": [vp, C.c_int, vp],
        "dtfft_mem_alloc": [vp, C.c_size_t, pvp],
        "dtfft_mem_free": [vp, vp],
        "dtfft_report": [vp],
        "dtfft_get_dims": [vp, C.POINTER(C.c_int8), C.POINTER(i32p)],
        "dtfft_get_grid_dims": [vp, C.POINTER(C.c_int8), C.POINTER(i32p)],
        "dtfft_get_stream": [vp, pvp],
        "dtfft_get_backend_pipelined": [C.c_int, C.POINTER(C.c_bool)],
        "dtfft_create_config": [vp],
        "dtfft_set_config": [vp],
        "dtfftb_plan_register_buffer": [vp, vp, C.c_size_t],
        "dtfftb_plan_unregister_buffer": [vp, vp],
        "dtfftb_plan_get_stats": [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
        "dtfftb_plan_peer_error": [vp],
        "dtfftb_plan_get_fallbacks": [vp, C.POINTER(C.c_int64)],
        "dtfftb_plan_get_exchange_form": [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "dtfftb_plan_set_overlap": [vp, C.c_int, C.c_int],
        "dtfftb_plan_set_graphs": [vp, C.c_int],
        "dtfftb_plan_get_overlap": [vp, C.POINTER(C.c_int)],
        "dtfftb_plan_get_graph_replays": [vp, C.POINTER(C.c_int64)],
        "dtfftb_plan_get_overlapped_stages": [vp, C.POINTER(C.c_int64)],
        "dtfftb_plan_create_dry": [C.c_int, C.c_int8, i32p, vp, vp, C.c_int, C.c_int, pvp],
        "dtfftb_plan_describe_exchange": [vp, C.c_int, C.c_int32, i32p, i32p, i32p, i32p, i32p, i32p,
                                          C.POINTER(C.c_int64), C.POINTER(C.c_int64), i32p],
        "dtfftb_plan_describe_chunk": [vp, C.c_int, C.c_int32, C.c_int32, C.c_int32, i32p, C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64)],
        "dtfftb_plan_describe_dma": [vp, C.c_int, C.c_int32, C.c_int32, i32p, i32p, i32p, i32p, C.POINTER(C.c_int64)],
        "dtfftb_plan_describe_peer_piece": [vp, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int32, i32p, C.POINTER(C.c_int64)],
        "dtfftb_plan_describe_reshape": [vp, C.c_int, C.c_int32, i32p, i32p, i32p, C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64), i32p],
    }
    for name in ("alloc_size", "alloc_bytes", "element_size", "aux_size", "aux_bytes", "aux_size_transpose",
                 "aux_bytes_transpose", "aux_size_reshape", "aux_bytes_reshape"):
        sig[f"dtfft_get_{name}"] = [vp, szp]
    for name in ("z_slab_enabled", "y_slab_enabled"):
        sig[f"dtfft_get_{name}"] = [vp, C.POINTER(C.c_bool)]
    for name in ("executor", "precision", "backend", "reshape_backend", "platform"):
        sig[f"dtfft_get_{name}"] = [vp, intp]
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name in ("dtfft_get_error_string", "dtfft_get_precision_string", "dtfft_get_executor_string",
                 "dtfft_get_backend_string"):
        fn = getattr(L, name)
        fn.argtypes = [C.c_int]
        fn.restype = C.c_char_p
    L.dtfft_get_version.restype = C.c_int32


def check(code: int, where: str) -> None:
    if code != 0:
        raise DtfftB200Error(code, where)
