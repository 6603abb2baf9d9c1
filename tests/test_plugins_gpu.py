"""Exchange-backend and FFT-executor plugin surfaces of the C ABI (include/dtfft_b200.h), driven
like the Fortran classes they replace: ``abstract_executor`` (create / execute(a, b, sign) /
destroy, src/dtfft_abstract_executor.F90:67-112) and ``abstract_backend`` + ``backend_nccl``
(src/dtfft_abstract_backend.F90:143-343, src/dtfft_backend_nccl.F90:65-134) on a one-rank
communicator, where the all-to-all degenerates to the self block.  The multi-rank exchange is
covered through the plan in tests/test_multi_gpu.py."""
import numpy as np
import pytest

from oracle import kernels as K
from tests.gpu_utils import to_device

pytestmark = pytest.mark.gpu


def rel_l2(got, want):
    return float(np.linalg.norm(got - want) / np.linalg.norm(want))


def test_executor_c2c_r2c_against_numpy(cuda):
    torch = cuda
    from dtfft_b200.plugins import FFT_BACKWARD, FFT_C2C, FFT_FORWARD, FFT_R2C, FftExecutor

    rng = np.random.default_rng(1234)
    # 1-D c2c, double: 50 transforms of 96 points, batched along the fastest axis
    n, batch = 96, 50
    x = (rng.random((batch, n)) + 1j * rng.random((batch, n))).astype(np.complex128)
    a, b = to_device(torch, x), torch.zeros(x.nbytes, dtype=torch.uint8, device="cuda")
    ex = FftExecutor().create(1, FFT_C2C, 1, n, n, batch, [n], [n], [n])
    ex.execute(a, b, FFT_FORWARD)
    torch.cuda.synchronize()
    got = b.cpu().numpy().view(np.complex128).reshape(batch, n)
    assert rel_l2(got, np.fft.fft(x, axis=1)) <= 1e-12
    ex.execute(b, a, FFT_BACKWARD)  # unnormalised, like the reference
    torch.cuda.synchronize()
    back = a.cpu().numpy().view(np.complex128).reshape(batch, n) / n
    assert np.max(np.abs(back - x)) <= 5 * np.log2(n) * 2 * np.finfo(np.float64).eps
    ex.destroy()

    # 1-D r2c / c2r, single: 66 reals -> 34 complex per transform
    n, batch = 66, 40
    r = rng.random((batch, n)).astype(np.float32)
    a = to_device(torch, r)
    b = torch.zeros(batch * (n // 2 + 1) * 8, dtype=torch.uint8, device="cuda")
    c = torch.zeros(r.nbytes, dtype=torch.uint8, device="cuda")
    ex = FftExecutor().create(1, FFT_R2C, 0, n, n // 2 + 1, batch, [n], [n], [n // 2 + 1])
    ex.execute(a, b, FFT_FORWARD)
    torch.cuda.synchronize()
    got = b.cpu().numpy().view(np.complex64).reshape(batch, n // 2 + 1).astype(np.complex128)
    assert rel_l2(got, np.fft.rfft(r.astype(np.float64), axis=1)) <= 1e-5
    ex.execute(b, c, FFT_BACKWARD)
    torch.cuda.synchronize()
    back = c.cpu().numpy().view(np.float32).reshape(batch, n) / n
    assert np.max(np.abs(back - r)) <= 5 * np.log2(n) * 2 * np.finfo(np.float32).eps
    ex.destroy()

    # 2-D c2c (the Z-slab executor): 6 planes of 24 x 40, sizes slowest first
    ny, nx, batch = 24, 40, 6
    x = (rng.random((batch, ny, nx)) + 1j * rng.random((batch, ny, nx))).astype(np.complex128)
    a, b = to_device(torch, x), torch.zeros(x.nbytes, dtype=torch.uint8, device="cuda")
    ex = FftExecutor().create(2, FFT_C2C, 1, nx * ny, nx * ny, batch, [ny, nx], [ny, nx], [ny, nx])
    ex.execute(a, b, FFT_FORWARD)
    torch.cuda.synchronize()
    got = b.cpu().numpy().view(np.complex128).reshape(batch, ny, nx)
    assert rel_l2(got, np.fft.fft2(x, axes=(1, 2))) <= 1e-12
    ex.destroy()

    # a rank without data gets a valid no-op handle; r2r is refused like the reference's cuFFT executor
    ex = FftExecutor().create(1, FFT_C2C, 1, n, n, 0, [n], [n], [n])
    ex.execute(a, b, FFT_FORWARD)
    ex.destroy()
    from dtfft_b200 import DtfftB200Error

    with pytest.raises(DtfftB200Error) as err:
        FftExecutor().create(1, 2, 1, n, n, 1, [n], [n], [n])
    assert err.value.code == 101  # DTFFT_ERROR_R2R_FFT_NOT_SUPPORTED


def test_nccl_backend_plugin_one_rank(cuda):
    torch = cuda
    from dtfft_b200.kernel import Kernel
    from dtfft_b200.plugins import BACKEND_NCCL, BACKEND_NCCL_PIPELINED, ExchangeBackend, NcclComm

    nccl = NcclComm()  # one-rank communicator
    rng = np.random.default_rng(7)
    dims = [18, 33, 20]
    n = int(np.prod(dims))
    src = (rng.random(n) + 1j * rng.random(n)).astype(np.complex128)
    stream = torch.cuda.Stream()

    # plain flavour: grouped send / recv of the only block = a copy in -> out
    a, b = to_device(torch, src), torch.zeros(src.nbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    be = ExchangeBackend().create(BACKEND_NCCL, nccl, 0, [0], [n], [0], [n], 16)
    assert be.aux_bytes == 0
    be.execute(a, b, stream)
    stream.synchronize()
    assert np.array_equal(b.cpu().numpy(), src.view(np.uint8))
    be.destroy()

    # pipelined flavour: self block copied into aux and unpacked by the borrowed per-peer kernel
    nd = np.array([[dims[0], dims[1], dims[2], 0, 0]], dtype=np.int32)
    gold = np.zeros(n, np.complex128)
    K.execute(K.KERNEL_UNPACK_PIPELINED, dims, src, gold, nd, 1)
    unpack = Kernel().create(dims, 0, 16, K.KERNEL_UNPACK_PIPELINED, nd)
    be = ExchangeBackend().create(BACKEND_NCCL_PIPELINED, nccl, 0, [0], [n], [0], [n], 16)
    be.set_unpack_kernel(unpack)
    assert be.aux_bytes == n * 16
    aux = torch.zeros(be.aux_bytes, dtype=torch.uint8, device="cuda")
    out = torch.zeros(src.nbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    from dtfft_b200 import DtfftB200Error

    with pytest.raises(DtfftB200Error):
        be.execute(a, out, stream)  # pipelined needs aux (DTFFT_ERROR_INVALID_AUX)
    be.execute(a, out, stream, aux)
    stream.synchronize()
    assert np.array_equal(out.cpu().numpy(), gold.view(np.uint8))
    be.destroy()
    unpack.destroy()
    nccl.destroy()
