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


@pytest.mark.parametrize("dims", [[40, 36, 34], [64, 33, 50]])
def test_y_slab_fft_matches_numpy(cuda, dims):
    """Y-slab plans (enable_y_slab with the Z-slab off): 1-D transforms along x, X -> Y, then ONE 2-D transform
    over (y, z) of the Y pencil (dtfft_plan.F90:2555-2556); the result stays in Y pencils (y, z, x)."""
    from dtfft_b200.plan import Execute, Executor
    from oracle import layout as L
    from tests.gpu_utils import to_device, to_host

    torch = cuda
    plan = PlanC2C(dims, executor=Executor.CUFFT, config=Config(enable_z_slab=False, enable_y_slab=True))
    assert plan.y_slab_enabled and not plan.z_slab_enabled
    G = P.global_array(dims, np.complex128)
    pencils = L.make_pencils(dims, [1, 1, 1], 0)
    x = P.pencil_slice(G, pencils[0])
    ins, inc, outs, outc, alloc = plan.local_sizes
    assert outc == pencils[1].counts
    a = to_device(torch, x)
    b = torch.full((plan.alloc_bytes,), 0xAB, dtype=torch.uint8, device="cuda")
    c = torch.full((plan.alloc_bytes,), 0xAB, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    plan.execute(a, b, Execute.FORWARD)
    torch.cuda.ExternalStream(plan.stream).synchronize()
    want = P.pencil_slice(np.asfortranarray(np.fft.fftn(G)), pencils[1])
    got = to_host(b, np.complex128)[: want.size]
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-12
    plan.execute(b, c, Execute.BACKWARD)
    torch.cuda.ExternalStream(plan.stream).synchronize()
    back = to_host(c, np.complex128)[: x.size] / np.prod(dims)
    assert np.max(np.abs(back - x)) <= 5 * np.log2(float(np.prod(dims))) * 2 * np.finfo(np.float64).eps
    plan.destroy()
    Config()._commit()
