"""GPU parity of the plan layer through the public C ABI (dtfft_create_plan_* / dtfft_transpose /
dtfft_execute / dtfft_mem_alloc ...), single GPU.  Mirrors the reference's integration tests
(tests/c/test_c2c_3d_c.c, tests/fortran/test_c2c_3d_f.F90: forward -> backward round trip within
5*log2(N)*2*eps, tests/test_utils.F90:96,107) and adds what they never pin: the CONTENT of every
intermediate layout (bit-exact vs the oracle) and forward spectra vs numpy.fft (<= 1e-12 fp64,
<= 1e-5 fp32 relative L2, BASELINE.json north_star)."""
import numpy as np
import pytest

from dtfft_b200.plan import (Config, DtfftError, Execute, Executor, Layout, PlanC2C, PlanR2C, PlanR2R, Precision,
                             Transpose)
from oracle import layout as L
from oracle import pipeline as P
from tests.gpu_utils import to_device, to_host

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def dev_empty(torch, nbytes):
    """Poisoned device buffer; synchronised because the plan runs on its OWN non-blocking stream."""
    t = torch.full((int(nbytes),), 0xAB, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    return t


def sync(torch, plan):
    torch.cuda.ExternalStream(plan.stream).synchronize() if plan.stream else torch.cuda.synchronize()


@pytest.mark.parametrize("dims", [[64, 64, 64], [129, 99, 33], [18, 33, 155], [33, 77, 21], [90, 57], [512, 64, 32]])
@pytest.mark.parametrize("dtype,cls,prec", [(np.complex128, PlanC2C, Precision.DOUBLE),
                                            (np.complex64, PlanC2C, Precision.SINGLE),
                                            (np.float64, PlanR2R, Precision.DOUBLE),
                                            (np.float32, PlanR2R, Precision.SINGLE)])
def test_transposes_bit_exact_one_rank(cuda, dims, dtype, cls, prec):
    """Every dtfft_transpose_t of a 1-rank plan lands exactly where the datatype path puts it."""
    torch = cuda
    plan = cls(dims, precision=prec, config=Config(enable_z_slab=True))
    assert plan.element_size == np.dtype(dtype).itemsize
    nd = len(dims)
    G = P.global_array(dims, dtype, kind="random")
    ttypes = [1, -1] if nd == 2 else [1, -1, 2, -2] + ([3, -3] if plan.z_slab_enabled else [])
    comm_dims = [1] * nd
    for t in ttypes:
        src = P.scatter_input(G, dims, comm_dims, t)[0]
        want = P.transpose_datatype(G, dims, comm_dims, t)[0]
        d_in, d_out = to_device(torch, src), dev_empty(torch, plan.alloc_bytes)
        plan.transpose(d_in, d_out, t)
        sync(torch, plan)
        got = to_host(d_out, dtype)[: want.size]
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), L.TRANSPOSE_NAMES[t]
    plan.destroy()
    Config()._commit()


@pytest.mark.parametrize("dims", [[64, 64, 64], [129, 99, 33], [40, 64], [512, 64, 32]])
@pytest.mark.parametrize("z_slab", [True, False])
def test_transpose_only_execute_round_trip(cuda, dims, z_slab):
    """dtfft_execute of a transpose-only plan: forward = X->Y->Z (or X->Z), backward returns the
    input bit for bit (dtfft_plan.F90:996-1003, 1072-1086)."""
    torch = cuda
    plan = PlanC2C(dims, config=Config(enable_z_slab=z_slab))
    nd = len(dims)
    G = P.global_array(dims, np.complex128)
    x = P.pencil_slice(G, L.make_pencils(dims, [1] * nd, 0)[0])
    want = P.pencil_slice(G, L.make_pencils(dims, [1] * nd, 0)[nd - 1])
    a, b = to_device(torch, x), dev_empty(torch, plan.alloc_bytes)
    c = dev_empty(torch, plan.alloc_bytes)
    aux = dev_empty(torch, plan.aux_bytes)
    plan.execute(a, b, Execute.FORWARD, aux)
    sync(torch, plan)
    assert np.array_equal(to_host(b, np.complex128)[: want.size].view(np.uint8), want.view(np.uint8))
    plan.execute(b, c, Execute.BACKWARD)  # aux allocated internally on first use (check_aux)
    sync(torch, plan)
    assert np.array_equal(to_host(c, np.complex128)[: x.size].view(np.uint8), x.view(np.uint8))
    st = plan.stats()
    assert st["kernel_launches"] >= 1 and st["local_bytes"] > 0 and st["remote_bytes"] == 0
    plan.destroy()
    Config()._commit()


@pytest.mark.parametrize("dims", [[129, 99, 33], [90, 57]])
def test_execute_graph_replay_is_bit_identical(cuda, dims):
    """The second dtfft_execute with the same buffers is captured into a CUDA graph and later calls
    replay it: results must equal the eager first call bit for bit (transposes and cuFFT alike),
    and freeing a buffer must drop the graphs."""
    torch = cuda
    from dtfft_b200.plan import Executor

    for executor in (Executor.NONE, Executor.CUFFT):
        plan = PlanC2C(dims, executor=executor, config=Config(enable_z_slab=False))
        nd = len(dims)
        G = P.global_array(dims, np.complex128, kind="random")
        x = P.pencil_slice(G, L.make_pencils(dims, [1] * nd, 0)[0])
        hx = torch.from_numpy(x.view(np.uint8).copy())
        ba, bb = plan.mem_alloc(plan.alloc_bytes), plan.mem_alloc(plan.alloc_bytes)
        a, b = torch.as_tensor(ba, device="cuda"), torch.as_tensor(bb, device="cuda")
        outs = []
        for i in range(4):
            a[: hx.numel()] = hx.cuda()
            b.fill_(0xAB)
            torch.cuda.synchronize()
            plan.execute(a, b, Execute.FORWARD)
            sync(torch, plan)
            outs.append(b.cpu().numpy().copy())
        assert plan.graph_replays >= 2
        for o in outs[1:]:
            assert np.array_equal(o, outs[0])
        plan.set_graphs(False)
        a[: hx.numel()] = hx.cuda()
        torch.cuda.synchronize()
        n0 = plan.graph_replays
        plan.execute(a, b, Execute.FORWARD)
        sync(torch, plan)
        assert plan.graph_replays == n0 and np.array_equal(b.cpu().numpy(), outs[0])
        plan.mem_free(ba)
        plan.mem_free(bb)
        plan.destroy()
    Config()._commit()


@pytest.mark.parametrize("dims", [[64, 48, 40], [129, 99, 33], [64, 96], [512, 64, 32]])
@pytest.mark.parametrize("prec,cdtype,tol", [(Precision.DOUBLE, np.complex128, 1e-12), (Precision.SINGLE, np.complex64, 1e-5)])
@pytest.mark.parametrize("z_slab", [True, False])
def test_c2c_fft_matches_numpy(cuda, dims, prec, cdtype, tol, z_slab):
    """Forward spectrum vs numpy.fft.fftn (pocketfft, fp64) and the reference's round-trip check."""
    torch = cuda
    plan = PlanC2C(dims, precision=prec, executor=Executor.CUFFT, config=Config(enable_z_slab=z_slab))
    nd = len(dims)
    G = P.global_array(dims, cdtype)
    x = P.pencil_slice(G, L.make_pencils(dims, [1] * nd, 0)[0])
    a, b = to_device(torch, x), dev_empty(torch, plan.alloc_bytes)
    plan.execute(a, b, Execute.FORWARD)
    sync(torch, plan)
    spec = np.fft.fftn(G.astype(np.complex128))
    want = P.pencil_slice(np.asfortranarray(spec), L.make_pencils(dims, [1] * nd, 0)[nd - 1])
    got = to_host(b, cdtype)[: want.size].astype(np.complex128)
    assert rel_l2(got, want) <= tol
    c = dev_empty(torch, plan.alloc_bytes)
    plan.execute(b, c, Execute.BACKWARD)
    sync(torch, plan)
    back = to_host(c, cdtype)[: x.size].astype(np.complex128) / np.prod(dims)
    n = float(np.prod(dims))
    eps = np.finfo(np.float64 if prec == Precision.DOUBLE else np.float32).eps
    assert np.max(np.abs(back - x.astype(np.complex128))) <= 5 * np.log2(n) * 2 * eps  # tests/test_utils.F90:96,107
    plan.destroy()
    Config()._commit()


@pytest.mark.parametrize("dims", [[64, 48, 40], [129, 99, 33], [64, 96], [33, 20, 18]])
@pytest.mark.parametrize("prec,rdtype,cdtype,tol", [(Precision.DOUBLE, np.float64, np.complex128, 1e-12),
                                                    (Precision.SINGLE, np.float32, np.complex64, 1e-5)])
@pytest.mark.parametrize("z_slab", [True, False])
def test_r2c_fft_matches_numpy(cuda, dims, prec, rdtype, cdtype, tol, z_slab):
    """R2C: real X pencil in, complex last pencil (nx/2+1 along x) out; C2R round trip."""
    torch = cuda
    plan = PlanR2C(dims, precision=prec, executor=Executor.CUFFT, config=Config(enable_z_slab=z_slab))
    nd = len(dims)
    cdims = [dims[0] // 2 + 1] + dims[1:]
    G = P.global_array(dims, rdtype)
    x = np.ascontiguousarray(G.reshape(-1, order="F"))
    ins, inc, outs, outc, alloc = plan.local_sizes
    assert inc == dims and plan.element_size == np.dtype(rdtype).itemsize
    a = dev_empty(torch, plan.alloc_bytes)
    a[: x.nbytes] = to_device(torch, x)
    b = dev_empty(torch, plan.alloc_bytes)
    torch.cuda.synchronize()
    plan.execute(a, b, Execute.FORWARD)
    sync(torch, plan)
    spec = np.fft.rfftn(G.astype(np.float64).transpose(tuple(range(nd - 1, -1, -1)))).transpose(tuple(range(nd - 1, -1, -1)))
    want = P.pencil_slice(np.asfortranarray(spec), L.make_pencils(cdims, [1] * nd, 0)[nd - 1])
    assert outc == L.make_pencils(cdims, [1] * nd, 0)[nd - 1].counts
    got = to_host(b, cdtype)[: want.size].astype(np.complex128)
    assert rel_l2(got, want) <= tol
    c = dev_empty(torch, plan.alloc_bytes)
    plan.execute(b, c, Execute.BACKWARD)
    sync(torch, plan)
    back = to_host(c, rdtype)[: x.size].astype(np.float64) / np.prod(dims)
    eps = np.finfo(rdtype).eps
    assert np.max(np.abs(back - x.astype(np.float64))) <= 5 * np.log2(float(np.prod(dims))) * 2 * eps
    plan.destroy()
    Config()._commit()


def test_error_codes_and_memory(cuda):
    """Input validation of execute / transpose (dtfft_plan.F90:722-742, 794-830) and mem_alloc."""
    torch = cuda

    def code(fn):
        with pytest.raises(DtfftError) as e:
            fn()
        return e.value.code

    plan = PlanC2C([32, 16, 8], config=Config(enable_z_slab=False))
    a, b = dev_empty(torch, plan.alloc_bytes), dev_empty(torch, plan.alloc_bytes)
    assert code(lambda: plan.transpose(a, a, Transpose.X_TO_Y)) == 14       # in-place transpose
    assert code(lambda: plan.transpose(a, b, 5)) == 2                        # invalid type
    assert code(lambda: plan.transpose(a, b, Transpose.X_TO_Z)) == 2         # X<->Z needs a Z-slab plan
    assert code(lambda: plan.transpose(a, b, Transpose.X_TO_Y, aux=a)) == 15  # aux aliases in
    assert code(lambda: plan.execute(a, b, 99)) == 43
    host = torch.zeros(plan.alloc_bytes, dtype=torch.uint8).pin_memory()
    assert code(lambda: plan.execute(host.data_ptr(), b, Execute.FORWARD)) == 300  # not a device pointer
    buf = plan.mem_alloc(plan.alloc_bytes)
    t = torch.as_tensor(buf, device="cuda")
    assert t.data_ptr() == buf.ptr and t.numel() == plan.alloc_bytes
    t.fill_(7)
    torch.cuda.synchronize()
    plan.transpose(t, b, Transpose.X_TO_Y)
    sync(torch, plan)
    assert int(b[: 32 * 16 * 8 * 16].min()) == 7
    plan.mem_free(buf)
    assert code(lambda: plan.mem_free(a)) == 20
    assert code(lambda: plan.mem_alloc(0)) == 21
    p2 = PlanC2C([32, 16], config=Config())
    assert code(lambda: p2.execute(a, a, Execute.FORWARD)) == 14              # 2-D in-place transpose-only
    plan.destroy()
    p2.destroy()
    assert code(lambda: plan.transpose(a, b, Transpose.X_TO_Y)) == 1         # destroyed plan
    Config()._commit()


def test_full_size_cycle_properties(cuda):
    """BASELINE config 2 at full size (512^3 c128): the X->Y->Z->Y->X cycle is the identity and the
    forward half equals an independent torch permute -- size-independent properties, no oracle."""
    torch = cuda
    n = 512
    plan = PlanC2C([n, n, n], config=Config(enable_z_slab=False))
    assert plan.alloc_bytes == n ** 3 * 16
    a = torch.rand(2 * n ** 3, dtype=torch.float64, device="cuda")
    ref = a.clone()
    b, c = torch.empty_like(a), torch.empty_like(a)
    torch.cuda.synchronize()
    plan.transpose(a, b, Transpose.X_TO_Y)
    plan.transpose(b, c, Transpose.Y_TO_Z)
    sync(torch, plan)
    # Z pencil (z,x,y) from X pencil (x,y,z): torch view [z][y][x][2] -> [y][x][z][2]
    want = ref.view(n, n, n, 2).permute(1, 2, 0, 3).contiguous().view(-1)
    assert torch.equal(c, want)
    del want
    plan.transpose(c, b, Transpose.Z_TO_Y)
    plan.transpose(b, a, Transpose.Y_TO_X)
    sync(torch, plan)
    assert torch.equal(a, ref)
    plan.destroy()
    Config()._commit()
