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


def test_all_backends_multi_gpu(cuda):
    torch = cuda
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-6000:]
    assert out.stdout.count("multi-GPU plan checks OK") == world
