"""Oracle (test infrastructure): ctypes access to the REFERENCE's regenerated device kernels
(``oracle/_ref/libref_kernels.so``, built by ``make -C oracle`` from the reference's own
``get_code()``; see gen_ref_kernels.py).  Used as a GPU-side checker and as the A/B
baseline "kernel to beat"; never imported by the product."""
from __future__ import annotations

import ctypes as C
import os

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_kernels.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
        _lib.ref_kernel_launch.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]
        _lib.ref_kernel_launch.restype = C.c_int
    return _lib


def configs():
    buf = (C.c_int * 64)()
    n = lib().ref_kernel_configs(buf, 32)
    return [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]


def valid_config(es: int, tile: int, rows: int) -> bool:
    """The reference's candidate filters (src/dtfft_nvrtc_block_optimizer.F90:600-639)."""
    return tile * (tile + 1) * es < 0.9 * 48 * 1024 and 64 <= tile * rows <= 1024 and tile >= rows


def launch(kernel_type, dims, es, tile, rows, out_ptr, in_ptr, nd5=None, stream=0):
    """One launch of the reference kernel ``kernel_type`` (per-neighbour kinds take ``nd5``)."""
    ndims = len(dims)
    nx, ny = int(dims[0]), int(dims[1])
    nz = int(dims[2]) if ndims == 3 else 1
    l = [0, 0, 0, 0, 0] if nd5 is None else [int(v) for v in nd5]
    if ndims == 2:
        l[2] = 1
    rc = lib().ref_kernel_launch(int(kernel_type), ndims, int(es), tile, rows, out_ptr, in_ptr, nx, ny, nz, *l, stream)
    if rc != 0:
        raise RuntimeError(f"ref_kernel_launch rc={rc} (kernel {kernel_type}, es {es}, tile {tile}x{rows})")
