"""The C/OpenMP port of the reference host kernels (oracle/host_kernels.c, the CPU baseline that
bench.py times) against the numpy oracle: unblocked loops of
src/include/_dtfft_kernel_host_routines.inc and the blocked permutes of
src/include/_dtfft_kernel_host_block_routines.inc (BLOCK_SIZE 4..64), on the dims the reference's own
host-kernel test uses (src/tests/test_host_kernels.F90:8-22: [33,77,21], [90,57]) -- "all host
variants agree", bit for bit."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import kernels as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def port():
    path = os.path.join(ROOT, "oracle", "_build", "liboracle_host.so")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_build/liboracle_host.so"], check=True,
                   stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(path)
    i32p = ctypes.POINTER(ctypes.c_int32)
    lib.oracle_kernel_execute.argtypes = [ctypes.c_int, ctypes.c_int, i32p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          i32p, ctypes.c_int, ctypes.c_int]
    lib.oracle_kernel_execute_blocked.argtypes = [ctypes.c_int, ctypes.c_int, i32p, ctypes.c_int, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_int]
    lib.oracle_set_num_threads(4)
    return lib


DTYPES = {4: np.float32, 8: np.float64, 16: np.complex128}
PERMUTES = (K.KERNEL_PERMUTE_FORWARD, K.KERNEL_PERMUTE_BACKWARD, K.KERNEL_PERMUTE_BACKWARD_START)


def _src(n, es):
    v = np.arange(1, n + 1)  # in(i) = i, test_host_kernels.F90:35-37
    return (v + 1j * (v + 0.5)).astype(np.complex128) if es == 16 else v.astype(DTYPES[es])


@pytest.mark.parametrize("dims", [[33, 77, 21], [90, 57], [64, 64, 64], [5, 3, 70]])
@pytest.mark.parametrize("es", [4, 8, 16])
def test_port_permutes_match_numpy_oracle(port, dims, es):
    n = int(np.prod(dims))
    src = _src(n, es)
    cdims = (ctypes.c_int32 * len(dims))(*dims)
    for kt in PERMUTES:
        if len(dims) == 2 and kt == K.KERNEL_PERMUTE_BACKWARD_START:
            continue
        gold = np.zeros(n, src.dtype)
        K.execute(kt, dims, src, gold)
        got = np.zeros(n, src.dtype)
        assert port.oracle_kernel_execute(kt, len(dims), cdims, es, src.ctypes.data, got.ctypes.data, None, 0, 0) == 0
        assert np.array_equal(got.view(np.uint8), gold.view(np.uint8)), (kt, "unblocked")
        for block in (4, 8, 16, 32, 64):
            got = np.zeros(n, src.dtype)
            assert port.oracle_kernel_execute_blocked(kt, len(dims), cdims, es, src.ctypes.data, got.ctypes.data, block) == 0
            assert np.array_equal(got.view(np.uint8), gold.view(np.uint8)), (kt, block)


def test_port_rejects_unknown_block(port):
    dims = (ctypes.c_int32 * 3)(4, 4, 4)
    a = np.zeros(64, np.float64)
    assert port.oracle_kernel_execute_blocked(K.KERNEL_PERMUTE_FORWARD, 3, dims, 8, a.ctypes.data, a.ctypes.data, 5) == -3
    assert port.oracle_kernel_execute_blocked(K.KERNEL_UNPACK, 3, dims, 8, a.ctypes.data, a.ctypes.data, 8) == -1


def test_port_unpack_matches_numpy_oracle(port):
    """A multi-peer unpack with real plan geometry (uneven split), as in smoke()."""
    from oracle import layout as L

    gdims, grid = (48, 21, 36), (1, 3, 2)
    pencils, geos = L.plan_geometry(gdims, grid, L.Y_TO_X)
    g = geos[1]
    alloc = max(p.size for p in pencils[1])
    buf = np.random.default_rng(7).random(alloc)
    gold = np.zeros(alloc, np.float64)
    K.execute(g.unpack_kernel, g.recv_dims, buf, gold, g.recv_nd)
    got = np.zeros(alloc, np.float64)
    nd = np.ascontiguousarray(np.asarray(g.recv_nd, dtype=np.int32))
    cdims = (ctypes.c_int32 * 3)(*g.recv_dims)
    rc = port.oracle_kernel_execute(g.unpack_kernel, 3, cdims, 8, buf.ctypes.data, got.ctypes.data,
                                    nd.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), nd.shape[0], 0)
    assert rc == 0
    assert np.array_equal(got.view(np.uint8), gold.view(np.uint8))
