#!/usr/bin/env python
"""Smallest multi-rank run of the NVLINK_FUSED backend, for compute-sanitizer (memcheck / racecheck follow the
torchrun children with --target-processes all):

    compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 \\
        python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/sanitize_fused.py

Both forms of the exchange (DTFFTB_FUSED_MODE=store | dma with sliced blocks) on a slab-shaped grid: every
transposition, then dtfft_execute forward + backward (peer-by-peer pair pipeline in the dma form), checked against
the index-encoded analytic pencils.  With DTFFTB_ALLOW_SHARED_DEVICE=1 the ranks share cuda:0 (1-GPU boxes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from bench import encode, expected_pencil, holds
    from dtfft_b200.comm import TorchComm
    from dtfft_b200.plan import Backend, Config, Execute, Layout, PlanC2C, Transpose

    shared = os.environ.get("DTFFTB_ALLOW_SHARED_DEVICE", "0") == "1"
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if shared:
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = TorchComm()
    dims = [64, 24, 16 * world]
    ok_all = True
    for mode in ("store", "dma"):
        os.environ["DTFFTB_FUSED_MODE"] = mode
        os.environ["DTFFTB_DMA_SUB_BYTES"] = "4096"
        stream = torch.cuda.Stream()
        plan = PlanC2C(dims, comm=TorchComm(cart_dims=[1, 1, world]),
                       config=Config(backend=Backend.NVLINK_FUSED, enable_z_slab=False, stream=stream))
        n = plan.alloc_bytes // 8
        a, b = (torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(2))
        pen = [plan.get_pencil(l) for l in (Layout.X_PENCILS, Layout.Y_PENCILS, Layout.Z_PENCILS)]
        with torch.cuda.stream(stream):
            want = [expected_pencil(torch, pen[d], d, dims, a.device) for d in range(3)]
            encode(torch, a, want[0])
        stream.synchronize()
        seq = [(Transpose.X_TO_Y, a, b, 1), (Transpose.Y_TO_Z, b, a, 2), (Transpose.Z_TO_Y, a, b, 1), (Transpose.Y_TO_X, b, a, 0)]
        for t, x, y, d in seq:
            dist.barrier()
            plan.transpose(x, y, t)
            stream.synchronize()
            with torch.cuda.stream(stream):
                ok = holds(torch, y, want[d])
            ok_all &= ok
            print(f"rank {rank} {mode} {t.name}: {'bit-exact' if ok else 'MISMATCH'}", flush=True)
        for rep in range(2):
            dist.barrier()
            plan.execute(a, b, Execute.FORWARD)
            stream.synchronize()
            with torch.cuda.stream(stream):
                okf = holds(torch, b, want[2])
            plan.execute(b, a, Execute.BACKWARD)
            stream.synchronize()
            with torch.cuda.stream(stream):
                okb = holds(torch, a, want[0])
            ok_all &= okf and okb
            print(f"rank {rank} {mode} execute #{rep}: fwd {okf} bwd {okb} pipelined stages {plan.overlapped_stages}", flush=True)
        assert plan.peer_error() == 0
        dist.barrier()
        plan.destroy()
    Config()._commit()
    dist.barrier()
    print(f"rank {rank}: sanitize_fused {'OK' if ok_all else 'FAILED'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
