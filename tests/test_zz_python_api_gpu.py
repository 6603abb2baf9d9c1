"""Python-side helpers of the reference module on the GPU (src/interfaces/python/__init__.py):
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
