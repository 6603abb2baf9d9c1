"""Checks of what was written after the round-1 GPU budget was spent and therefore runs on a GPU for the
first time in this file (kept last in collection order): the two wide tile variants, and the
Python-side helpers of the reference module on the GPU (src/interfaces/python/__init__.py):
``Plan.get_ndarray`` (array backed by ``dtfft_mem_alloc`` memory, here a CUDA torch tensor instead of a
cupy array) and the ``Request`` objects of ``transpose_start`` / ``transpose_end``."""
import gc

import numpy as np
import pytest

from dtfft_b200.plan import Config, DtfftError, PlanC2C, Request, Transpose
from oracle import pipeline as P

pytestmark = pytest.mark.gpu


def test_get_ndarray_and_async_requests(cuda):
    torch = cuda
    dims = [24, 20, 16]
    plan = PlanC2C(dims, config=Config(enable_z_slab=False))
    n = plan.alloc_size
    a = plan.get_ndarray(n, shape=dims, order="F")                            # a[x, y, z], x fastest
    b = plan.get_ndarray(n, shape=(dims[1], dims[2], dims[0]), order="F")     # b[y, z, x], y fastest
    assert a.is_cuda and a.dtype == torch.complex128 and tuple(a.shape) == tuple(dims)
    assert a.stride() == (1, dims[0], dims[0] * dims[1])
    flat = plan.get_ndarray(n)
    assert tuple(flat.shape) == (n,) and flat.dtype == torch.complex128
    G = P.global_array(dims, np.complex128)
    a.copy_(torch.from_numpy(np.ascontiguousarray(G)).cuda())
    b.zero_()
    torch.cuda.synchronize()
    req = plan.transpose_start(a, b, Transpose.X_TO_Y)
    assert isinstance(req, Request) and req.handle != 0 and "X_TO_Y" in req.kind
    with pytest.raises(DtfftError) as e:
        plan.reshape_end(req)            # a transposition is not a reshape
    assert e.value.code == 35
    plan.transpose_end(req)
    with pytest.raises(DtfftError) as e:
        plan.transpose_end(req)          # retired
    assert e.value.code == 35
    torch.cuda.ExternalStream(plan.stream).synchronize()
    assert np.array_equal(b.cpu().numpy(), G.transpose(1, 2, 0))   # Y pencil: b[y, z, x] = G[x, y, z]
    # the arrays own their allocations: dropping them frees the memory through the plan
    del a, b, flat
    gc.collect()
    plan.destroy()
    Config()._commit()


def test_wide_tile_variants_agree(cuda):
    """32 x 128 and 128 x 32 tiles (candidates of the DTFFT_EXHAUSTIVE kernel autotune since the end of
    round 1) give the oracle's bytes for 4-, 8- and 16-byte elements on shapes that do not divide the tile."""
    from dtfft_b200.kernel import Kernel
    from oracle import kernels as K
    from tests.gpu_utils import device_filled, to_device, to_host

    torch = cuda
    for dims in ([70, 45, 19], [130, 33, 5], [200, 150]):
        n = int(np.prod(dims))
        for dtype in (np.float32, np.float64, np.complex128):
            src = np.random.default_rng(2).random(n).astype(dtype)
            for kt in (K.KERNEL_PERMUTE_FORWARD, K.KERNEL_PERMUTE_BACKWARD):
                gold = np.zeros(n, dtype)
                K.execute(kt, dims, src, gold)
                d_in = to_device(torch, src)
                k = Kernel().create(dims, 0, np.dtype(dtype).itemsize, kt)
                for cfg in [(1, 4, 16), (4, 1, 16)]:
                    k.set_tile(*cfg)
                    d_out = device_filled(torch, n, dtype)
                    k.execute(d_in, d_out, sync=True)
                    assert np.array_equal(to_host(d_out, dtype).view(np.uint8), gold.view(np.uint8)), (dims, dtype, kt, cfg)
                k.destroy()
