"""N > 1 GPUs: every backend of the reshape path (NCCL, NCCL_PIPELINED, NVLINK_FUSED) through the
public plan API under torchrun, one rank per GPU (tests/_gpu_worker.py).  Skipped on 1-GPU boxes;
run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
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


def run_worker(torch, extra_env):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "_gpu_worker.py")]
    env = dict(os.environ)
    env.update(extra_env)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-6000:]
    assert out.stdout.count("multi-GPU plan checks OK") == world


def test_all_backends_multi_gpu(cuda):
    run_worker(cuda, {})


def test_fused_backend_copy_engine_form_multi_gpu(cuda):
    """NVLINK_FUSED with every exchange in its copy-engine form (pack -> one strided 3-D copy per peer; the
    automatic rule keeps the direct-store kernel below 1 MiB per peer, i.e. for every test-sized case) and the
    peer-by-peer pipelines of Plan::run_transpose_pair."""
    run_worker(cuda, {"DTFFTB_FUSED_MODE": "dma", "DTFFTB_TEST_BACKENDS": "NVLINK_FUSED", "DTFFTB_TEST_EXPERIMENTAL": "1",
                      "DTFFTB_DMA_SUB_BYTES": "4096"})  # blocks cut into slices even at test sizes
