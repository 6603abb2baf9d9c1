"""The timed kernel autotune (src/dtfft_kernel_device.F90:338-397: every tile candidate is timed on
scratch buffers -- 2 warm-up + 5 iterations by default -- and the fastest kept; :385-389 logs time and
bandwidth per candidate).  Here: dtfftb_kernel_autotune[_report] and effort = DTFFT_EXHAUSTIVE /
enable_kernel_autotune through the plan.  The chosen tile must be a compiled instantiation and the
kernel must stay bit-exact whatever was picked."""
import numpy as np
import pytest

from dtfft_b200.kernel import Kernel
from oracle import kernels as K
from tests.gpu_utils import device_filled, host_filled, to_device, to_host

pytestmark = pytest.mark.gpu

SUPPORTED = {(32, 32, 128), (32, 32, 256), (32, 32, 512), (64, 32, 256), (32, 64, 256), (64, 64, 256), (64, 64, 512),
             (32, 128, 512), (128, 32, 512)}


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex128])
@pytest.mark.parametrize("kt", [K.KERNEL_PERMUTE_FORWARD, K.KERNEL_PERMUTE_BACKWARD, K.KERNEL_PERMUTE_BACKWARD_START])
def test_autotune_report_keeps_a_supported_tile_and_stays_bit_exact(cuda, kt, dtype):
    torch = cuda
    dims = [200, 136, 72]  # not multiples of any tile: edge tiles on every candidate
    n = int(np.prod(dims))
    es = np.dtype(dtype).itemsize
    rng = np.random.default_rng(3)
    src = rng.random(n).astype(dtype)
    if np.dtype(dtype).kind == "c":
        src = src + 1j * rng.random(n)
    gold = host_filled(n, dtype)
    K.execute(kt, dims, src, gold)
    d_in, d_out = to_device(torch, src), device_filled(torch, n, dtype)
    k = Kernel().create(dims, 0, es, kt)
    log = k.autotune_report(d_in, d_out, None, 2, 5)
    assert len(log) == len(SUPPORTED)  # every compiled candidate was timed
    for e in log:
        assert (e["tile_a"], e["tile_b"], e["threads"]) in SUPPORTED
        assert e["ms"] > 0 and e["gbs"] > 0
        # the bandwidth the log reports is 2 x bytes / time, like the reference's
        assert e["gbs"] == pytest.approx(2 * n * es / (e["ms"] * 1e-3) / 1e9, rel=1e-3)
    info = k.info()
    best = min(log, key=lambda e: e["ms"])
    assert (info["tile_a"], info["tile_b"], info["threads"]) == (best["tile_a"], best["tile_b"], best["threads"])
    d_out.fill_(0)
    k.execute(d_in, d_out, sync=True)
    assert np.array_equal(to_host(d_out, dtype).view(np.uint8), gold.view(np.uint8))
    k.destroy()


def test_exhaustive_effort_autotunes_the_plan_kernels(cuda, capfd):
    """DTFFT_EXHAUSTIVE runs the timed autotune at plan creation (kernel_device.F90:328-337); the plan
    stays bit-exact against the oracle's datatype-path truth."""
    import os

    torch = cuda
    from dtfft_b200.plan import Config, Effort, Execute, PlanC2C
    from oracle import layout as L
    from oracle import pipeline as P

    os.environ["DTFFTB_LOG"] = "1"
    try:
        pdims = [80, 48, 56]
        plan = PlanC2C(pdims, effort=Effort.EXHAUSTIVE, config=Config(enable_z_slab=False))
    finally:
        os.environ.pop("DTFFTB_LOG", None)
    err = capfd.readouterr().err
    assert "autotune es=16" in err and "GB/s" in err
    G = P.global_array(pdims, np.complex128)
    x = P.pencil_slice(G, L.make_pencils(pdims, [1, 1, 1], 0)[0])
    want = P.pencil_slice(G, L.make_pencils(pdims, [1, 1, 1], 0)[2])
    a = to_device(torch, x)
    b = torch.zeros(x.nbytes, dtype=torch.uint8, device="cuda")
    c = torch.zeros(x.nbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    plan.execute(a, b, Execute.FORWARD)
    plan.execute(b, c, Execute.BACKWARD)
    torch.cuda.ExternalStream(plan.stream).synchronize()
    assert np.array_equal(b.cpu().numpy().view(np.uint8), want.view(np.uint8))
    assert np.array_equal(c.cpu().numpy().view(np.uint8), x.view(np.uint8))
    plan.destroy()
    Config()._commit()
