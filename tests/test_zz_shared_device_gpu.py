"""Multi-RANK parity on a ONE-GPU box: 2 and 4 processes time-slice cuda:0 (DTFFTB_ALLOW_SHARED_DEVICE=1) and
run tests/_gpu_worker.py on the NVLINK_FUSED backend -- the P > 1 path of handle.cu (global-index intersection
boxes, peer-mapped destinations through cudaIpc, device barriers, chunked stage overlap, CUDA-graph replay,
brick reshapes, any-pointer publication) against the oracle's datatype-path truth, bit for bit.  NCCL refuses
several ranks per device, so the NCCL backends keep needing >= 2 GPUs (tests/test_multi_gpu.py, bench.py's
parity block at N > 1).  The reference registers every integration test at nproc = 1..N
(tests/CMakeLists.txt:27-35, 83-89); this is the same idea on the hardware the driver has."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,mode", [(2, "store"), (4, "store"), (2, "dma"), (3, "dma")])
def test_fused_backend_ranks_sharing_one_gpu(cuda, world, mode):
    """mode: the direct-store kernel, or every exchange in its copy-engine form (pack + strided 3-D copy per peer,
    peer-by-peer pair pipelines with pairwise landed flags)."""
    env = dict(os.environ)
    env.update({"DTFFTB_ALLOW_SHARED_DEVICE": "1", "DTFFTB_PEER_TIMEOUT_MS": "20000", "DTFFTB_TEST_EXPERIMENTAL": "1",
                "DTFFTB_FUSED_MODE": mode, "DTFFTB_DMA_SUB_BYTES": "4096", "OMP_NUM_THREADS": "1"})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-6000:]
    assert out.stdout.count("multi-GPU plan checks OK") == world
