"""GPU parity: every CUDA kernel kind, through the C ABI, bit-exact against the oracle.

Restates src/tests/test_device_kernels.F90:27-40 (device kernel == host kernel on dims
[18,155] and [18,33,155], one neighbour, zero displacements, fp32) and widens it to what the
reference never pins: 8- and 16-byte elements, several neighbours with
non-zero displacements taken from real plan geometry, uneven and zero-size peers.
"""
import numpy as np
import pytest

from dtfft_b200.kernel import Kernel
from oracle import kernels as K
from oracle import layout as L
from oracle import pipeline as P
from tests.gpu_utils import device_filled, host_filled, to_device, to_host

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64, np.complex128]


def run_both(torch, kt, dims, dtype, nd=None, neighbor=None, n_out=None, seed=0):
    n_in = int(np.prod(dims))
    n_out = n_out or n_in
    rng = np.random.default_rng(seed)
    host_in = (rng.random(n_in) * np.arange(1, n_in + 1)).astype(dtype)  # rnd * i like the reference test
    if np.dtype(dtype).kind == "c":
        host_in = host_in + 1j * rng.random(n_in)
    gold = host_filled(n_out, dtype)
    K.execute(kt, dims, host_in, gold, nd, neighbor)
    d_in = to_device(torch, host_in)
    d_out = device_filled(torch, n_out, dtype)
    k = Kernel().create(dims, 0, np.dtype(dtype).itemsize, kt, nd)
    k.execute(d_in, d_out, None, neighbor, sync=True)
    got = to_host(d_out, dtype)
    k.destroy()
    return got, gold


def whole_nd(dims):
    nd = np.zeros((1, 5), dtype=np.int32)
    nd[0, : len(dims)] = dims
    if len(dims) == 2:
        nd[0, 2] = 1
    return nd


REF_CASES = ([([18, 155], kt) for kt in (K.KERNEL_PERMUTE_FORWARD, K.KERNEL_PACK, K.KERNEL_UNPACK, K.KERNEL_PACK_FORWARD)]
             + [([18, 33, 155], kt) for kt in (K.KERNEL_PERMUTE_FORWARD, K.KERNEL_PERMUTE_BACKWARD,
                                                K.KERNEL_PERMUTE_BACKWARD_START, K.KERNEL_PERMUTE_BACKWARD_END,
                                                K.KERNEL_PACK, K.KERNEL_UNPACK, K.KERNEL_PACK_FORWARD,
                                                K.KERNEL_PACK_BACKWARD)])


@pytest.mark.parametrize("dims,kt", REF_CASES)
@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_device_kernel_cases(cuda, dims, kt, dtype):
    """The 12 cases of src/tests/test_device_kernels.F90 (there fp32 only; here 4/8/16 B)."""
    nb = 1 if K.effective_kernel_type(kt, len(dims)) in K.PER_NEIGHBOR_KERNELS else None
    got, gold = run_both(cuda, kt, dims, dtype, whole_nd(dims), nb)
    assert np.array_equal(got.view(np.uint8), gold.view(np.uint8))


@pytest.mark.parametrize("dims", [[33, 77, 21], [90, 57], [1, 5, 7], [64, 64, 64], [129, 3, 65], [7, 1, 9]])
@pytest.mark.parametrize("dtype", DTYPES)
def test_permutes_and_extra_kinds(cuda, dims, dtype):
    """dims of src/tests/test_host_kernels.F90 + degenerate extents; includes the UNPACK_FORWARD /
    UNPACK_BACKWARD kinds the reference only has on the host."""
    kinds = [K.KERNEL_PERMUTE_FORWARD, K.KERNEL_PERMUTE_BACKWARD, K.KERNEL_UNPACK_FORWARD, K.KERNEL_COPY]
    if len(dims) == 3:
        kinds += [K.KERNEL_PERMUTE_BACKWARD_START, K.KERNEL_UNPACK_BACKWARD]
    for kt in kinds:
        got, gold = run_both(cuda, kt, dims, dtype, whole_nd(dims), None)
        assert np.array_equal(got.view(np.uint8), gold.view(np.uint8)), K.KERNEL_NAMES[kt]


GEOM_CASES = [((48, 20, 36), (1, 2, 2)), ((13, 7, 9), (1, 3, 2)), ((9, 3, 5), (1, 4, 2)), ((40, 33, 28), (1, 1, 4)),
              ((66, 10, 12), (1, 4, 1)), ((37, 22), (1, 3)), ((128, 64, 96), (1, 2, 4))]


@pytest.mark.parametrize("dims,grid", GEOM_CASES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", ["plain", "pipelined", "fused"])
def test_multi_neighbor_kernels_from_plan_geometry(cuda, dims, grid, dtype, mode):
    """Pack / unpack kernels with the neighbor_data a real plan produces (non-zero din/dout,
    uneven and empty peers): device == oracle for every rank of the grid."""
    torch = cuda
    ndims = len(dims)
    tts = (L.X_TO_Y, L.Y_TO_X) if ndims == 2 else (L.X_TO_Y, L.Y_TO_X, L.Y_TO_Z, L.Z_TO_Y, L.X_TO_Z, L.Z_TO_X)
    es = np.dtype(dtype).itemsize
    for tt in tts:
        if abs(tt) == 3 and grid[1] != 1:
            continue
        pencils, geos = L.plan_geometry(dims, grid, tt, pipelined=mode != "plain", fused=mode == "fused")
        for r, g in enumerate(geos):
            if g.comm_size == 1:
                continue
            alloc = max(p.size for p in pencils[r])
            if alloc == 0:
                continue
            rng = np.random.default_rng(r)
            src = rng.random(alloc).astype(dtype)
            for (kt, kdims, nd) in ((g.pack_kernel, g.send_dims, g.send_nd), (g.unpack_kernel, g.recv_dims, g.recv_nd)):
                gold = host_filled(alloc, dtype)
                d_in = to_device(torch, src)
                d_out = device_filled(torch, alloc, dtype)
                k = Kernel().create(kdims, 0, es, kt, nd)
                if K.effective_kernel_type(kt, ndims) in K.PER_NEIGHBOR_KERNELS:
                    for n in range(1, g.comm_size + 1):
                        K.execute(kt, kdims, src, gold, nd, n)
                        k.execute(d_in, d_out, None, n)
                else:
                    K.execute(kt, kdims, src, gold, nd)
                    k.execute(d_in, d_out, None, None)
                torch.cuda.synchronize()
                got = to_host(d_out, dtype)
                assert np.array_equal(got.view(np.uint8), gold.view(np.uint8)), (L.TRANSPOSE_NAMES[tt], r, K.KERNEL_NAMES[kt])
                # one-launch-for-all-peers extension must equal the per-peer sequence
                d_out2 = device_filled(torch, alloc, dtype)
                k.execute_all(d_in, d_out2, None)
                torch.cuda.synchronize()
                assert torch.equal(d_out, d_out2)
                k.destroy()


@pytest.mark.parametrize("dims,grid", [((48, 20, 36), (1, 2, 2)), ((13, 7, 9), (1, 3, 2)), ((40, 33, 28), (1, 1, 4)), ((37, 22), (1, 3))])
@pytest.mark.parametrize("dtype", [np.float32, np.complex128])
def test_simulated_transposition_on_one_gpu(cuda, dims, grid, dtype):
    """Whole pack -> exchange -> unpack on simulated ranks with the CUDA kernels doing steps 1
    and 3 and device-to-device slices standing in for the exchange: equals the datatype path."""
    torch = cuda
    es = np.dtype(dtype).itemsize
    G = P.global_array(dims, dtype, kind="index")
    tts = (L.X_TO_Y, L.Y_TO_X) if len(dims) == 2 else (L.X_TO_Y, L.Y_TO_X, L.Y_TO_Z, L.Z_TO_Y, L.X_TO_Z, L.Z_TO_X)
    for tt in tts:
        if abs(tt) == 3 and grid[1] != 1:
            continue
        pencils, geos = L.plan_geometry(dims, grid, tt)
        ins = P.scatter_input(G, dims, grid, tt)
        ref = P.transpose_datatype(G, dims, grid, tt)
        n = len(geos)
        alloc = [max(1, max(p.size for p in pencils[r])) for r in range(n)]
        a = [device_filled(torch, alloc[r], dtype, 0) for r in range(n)]
        b = [device_filled(torch, alloc[r], dtype) for r in range(n)]
        for r in range(n):
            a[r][: ins[r].size * es] = to_device(torch, ins[r])
        for r, g in enumerate(geos):
            Kernel().create(g.send_dims, 0, es, g.pack_kernel, g.send_nd).execute(a[r], b[r])
        if geos[0].comm_size > 1:
            for r, g in enumerate(geos):
                for i, peer in enumerate(g.members):
                    gp = geos[peer]
                    j = gp.members.index(r)
                    cnt, so, ro = g.recv_counts[i] * es, gp.send_displs[j] * es, g.recv_displs[i] * es
                    a[r][ro: ro + cnt] = b[peer][so: so + cnt]
            for r, g in enumerate(geos):
                Kernel().create(g.recv_dims, 0, es, g.unpack_kernel, g.recv_nd).execute(a[r], b[r])
        torch.cuda.synchronize()
        for r in range(n):
            got = to_host(b[r], dtype)[: ref[r].size]
            assert np.array_equal(got.view(np.uint8), ref[r].view(np.uint8)), (L.TRANSPOSE_NAMES[tt], r)


def test_tile_configs_agree(cuda):
    """Every family-T tile configuration gives the same bytes (the reference checks its block
    variants against the base kernel the same way, test_host_kernels.F90:275-287)."""
    torch = cuda
    dims = [70, 45, 19]
    for dtype in DTYPES:
        n = int(np.prod(dims))
        src = np.random.default_rng(1).random(n).astype(dtype)
        gold = np.zeros(n, dtype)
        K.execute(K.KERNEL_PERMUTE_BACKWARD, dims, src, gold)
        d_in = to_device(torch, src)
        k = Kernel().create(dims, 0, np.dtype(dtype).itemsize, K.KERNEL_PERMUTE_BACKWARD)
        for cfg in [(1, 1, 4), (1, 1, 8), (1, 1, 16), (2, 1, 8), (1, 2, 8), (2, 2, 8), (2, 2, 16)]:
            k.set_tile(*cfg)
            d_out = device_filled(torch, n, dtype)
            k.execute(d_in, d_out, sync=True)
            assert np.array_equal(to_host(d_out, dtype).view(np.uint8), gold.view(np.uint8)), cfg


def test_zero_volume_and_dummy_are_noops(cuda):
    torch = cuda
    d_in = device_filled(torch, 16, np.float32, 1)
    d_out = device_filled(torch, 16, np.float32, 2)
    Kernel().create([0, 4, 4], 0, 4, K.KERNEL_PERMUTE_FORWARD).execute(d_in, d_out, sync=True)
    Kernel().create([2, 2, 4], 0, 4, K.KERNEL_DUMMY).execute(d_in, d_out, sync=True)
    assert int(d_out.min()) == 2 and int(d_out.max()) == 2


def test_full_size_permute_properties(cuda):
    """BASELINE config 2 size (512^3 complex128, 2 GiB per buffer): size-independent checks --
    the transpose-only cycle X->Y->Z->Y->X is the identity, every stage is a permutation
    (checksum of 64-bit words preserved) and sampled elements land where the index map says."""
    torch = cuda
    n = 512
    dims = [n, n, n]
    N = n ** 3
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randint(-2 ** 62, 2 ** 62, (2 * N,), dtype=torch.int64, device="cuda", generator=g)  # 16-byte elements
    y = torch.empty_like(x)
    fwd = Kernel().create(dims, 0, 16, K.KERNEL_PERMUTE_FORWARD)
    bwd = Kernel().create(dims, 0, 16, K.KERNEL_PERMUTE_BACKWARD)
    ref_sum = int(x.sum())
    fwd.execute(x, y)                       # X -> Y
    assert int(y.sum()) == ref_sum
    idx = torch.randint(0, N, (4096,), device="cuda", generator=g)
    xi, yi, zi = idx % n, (idx // n) % n, idx // (n * n)
    oidx = yi + zi * n + xi * n * n        # out[y + z*ny + x*ny*nz] = in[x + y*nx + z*nx*ny]
    assert torch.equal(y.view(-1, 2)[oidx], x.view(-1, 2)[idx])
    z = torch.empty_like(x)
    fwd.execute(y, z)                       # Y -> Z
    bwd.execute(z, y)                       # Z -> Y
    w = torch.empty_like(x)
    bwd.execute(y, w)                       # Y -> X
    torch.cuda.synchronize()
    assert torch.equal(w, x)
