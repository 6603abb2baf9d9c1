"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        if os.path.basename(h) == "dtfft_b200_mpi.h":
            continue  # header-only adapter for MPI programs: static inline functions, nothing exported
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"\b(dtfftb?_[a-z0-9_]+)\s*\(", src):
            names.add(m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import dtfft_b200

    lib = dtfft_b200.lib()
    names = declared_functions()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_version_and_no_cpu_fallback():
    import dtfft_b200
    from dtfft_b200.kernel import Kernel

    lib = dtfft_b200.lib()
    lib.dtfftb_version.restype = ctypes.c_char_p
    assert b"dtfft_b200" in lib.dtfftb_version()
    import numpy as np
    import pytest

    # host (numpy) buffers are rejected: the product never computes on the CPU
    with pytest.raises(TypeError):
        Kernel().execute(np.zeros(4), np.zeros(4))


def test_library_then_torch_share_one_nccl():
    """Loading libdtfft_b200.so BEFORE torch must not bind the older system libnccl: the loader
    preloads the NCCL that torch ships, so a later `import torch` finds its own symbols."""
    import subprocess
    import sys

    code = ("import dtfft_b200; dtfft_b200.lib(); import torch; "
            "n=[l.split()[-1] for l in open('/proc/self/maps') if 'libnccl' in l]; "
            "assert len(set(n)) == 1, set(n); print('one nccl')")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0 and "one nccl" in out.stdout, out.stderr[-2000:]


def test_plugin_abi_argument_validation():
    """Error behaviour of the backend / executor plugin entry points that needs no device: NULL handles,
    backend types outside the NCCL family, element sizes, the R2R refusal of the cuFFT executor
    (src/interfaces/fft/cufft/dtfft_executor_cufft_m.F90:94-98), bad precision / rank."""
    import ctypes as C

    import dtfft_b200

    L = dtfft_b200.lib()
    vp, i32p, i64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    L.dtfftb_backend_create.argtypes = [C.POINTER(vp), C.c_int, vp, C.c_int, C.c_int, i32p, i64p, i64p, i64p, i64p, C.c_int64]
    L.dtfftb_backend_create.restype = C.c_int
    L.dtfftb_executor_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_int32, i32p,
                                         i32p, i32p, vp]
    L.dtfftb_executor_create.restype = C.c_int
    h = vp(0)
    one = (C.c_int64 * 1)(0)
    fake_comm = vp(0x1000)  # never dereferenced before the checks below fail
    usage = L.dtfftb_backend_create(None, 24, fake_comm, 0, 1, None, one, one, one, one, 8)
    assert usage != 0
    assert L.dtfftb_backend_create(C.byref(h), 23, fake_comm, 0, 1, None, one, one, one, one, 8) == 202   # MPI_A2A: not here
    assert L.dtfftb_backend_create(C.byref(h), 24, None, 0, 1, None, one, one, one, one, 8) == usage      # no communicator
    assert L.dtfftb_backend_create(C.byref(h), 24, fake_comm, 1, 1, None, one, one, one, one, 8) == usage  # rank >= size
    assert L.dtfftb_backend_create(C.byref(h), 24, fake_comm, 0, 1, None, one, one, one, one, 12) == usage  # element size
    assert h.value in (None, 0)
    n = (C.c_int32 * 2)(8, 8)
    e = vp(0)
    assert L.dtfftb_executor_create(None, 1, 0, 1, 8, 8, 4, n, n, n, None) == usage
    assert L.dtfftb_executor_create(C.byref(e), 3, 0, 1, 8, 8, 4, n, n, n, None) == usage       # fft_rank
    assert L.dtfftb_executor_create(C.byref(e), 1, 2, 1, 8, 8, 4, n, n, n, None) == 101         # R2R: DTFFT_ERROR_R2R_FFT_NOT_SUPPORTED
    assert L.dtfftb_executor_create(C.byref(e), 1, 0, 7, 8, 8, 4, n, n, n, None) == 6           # precision
    assert L.dtfftb_executor_create(C.byref(e), 1, 0, 1, 8, 8, 4, None, n, n, None) == usage
    assert e.value in (None, 0)
