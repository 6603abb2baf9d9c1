#!/usr/bin/env python
"""First run of the DTFFT_MEASURE / DTFFT_PATIENT process-grid search on real GPUs (torchrun, one rank
per GPU): creates a default 512^3 C2C transpose-only plan at each effort level, prints the grid and
backend the search picked and the fwd+bwd cycle time of the resulting plan, beside the ESTIMATE plan.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/grid_search_probe.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from dtfft_b200.comm import TorchComm
    from dtfft_b200.plan import Backend, Config, Effort, Execute, PlanC2C

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank = dist.get_rank()
    comm = TorchComm()
    n = int(os.environ.get("PROBE_N", "512"))
    for effort in (Effort.ESTIMATE, Effort.MEASURE, Effort.PATIENT):
        cfg = Config(enable_z_slab=False, backend=Backend.NVLINK_FUSED)
        plan = PlanC2C([n, n, n], comm=comm, effort=effort, config=cfg)
        a, b, c = (plan.mem_alloc(plan.alloc_bytes) for _ in range(3))
        at = torch.as_tensor(a, device="cuda")
        at.zero_()
        torch.cuda.synchronize()
        stream = torch.cuda.ExternalStream(plan.stream)
        for _ in range(5):
            plan.execute(a, b, Execute.FORWARD)
            plan.execute(b, c, Execute.BACKWARD)
        stream.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record(stream)
        for _ in range(iters):
            plan.execute(a, b, Execute.FORWARD)
            plan.execute(b, c, Execute.BACKWARD)
        e1.record(stream)
        stream.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"effort": effort.name, "grid": plan.grid_dims, "backend": plan.backend.name,
                              "ms_per_cycle": float(ms)}), flush=True)
        for buf in (a, b, c):
            plan.mem_free(buf)
        plan.destroy()
    Config()._commit()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
