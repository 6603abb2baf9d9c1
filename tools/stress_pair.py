#!/usr/bin/env python
"""Race hunt for the copy-engine pair pipeline: dtfft_execute forward + backward of the headline plan, `--iters` times,
with the destination, the intermediate pencil AND the staging workspace poisoned before every call, every element of
every rank checked against the analytic pencils after every call.  (A pipeline that reads a slice before it has landed
would otherwise go unnoticed from the second call on: the stale bytes are the previous call's identical values.)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/stress_pair.py [--size 512]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from bench import encode, expected_pencil, holds
    from dtfft_b200.comm import TorchComm
    from dtfft_b200.plan import Backend, Config, Execute, Layout, PlanC2C

    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=512)
    ap.add_argument("--iters", type=int, default=50)
    args = ap.parse_args()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    shared = os.environ.get("DTFFTB_ALLOW_SHARED_DEVICE", "0") == "1"
    if shared:
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    dims = [args.n] * 3
    stream = torch.cuda.Stream()
    plan = PlanC2C(dims, comm=TorchComm(cart_dims=[1, 1, world]),
                   config=Config(backend=Backend.NVLINK_FUSED, enable_z_slab=False, stream=stream))
    bufs = [plan.mem_alloc(plan.alloc_bytes) for _ in range(2)]
    aux_buf = plan.mem_alloc(plan.aux_bytes)
    a, b = (torch.as_tensor(x, device="cuda").view(torch.float64) for x in bufs)
    aux = torch.as_tensor(aux_buf, device="cuda")
    pen = [plan.get_pencil(l) for l in (Layout.X_PENCILS, Layout.Y_PENCILS, Layout.Z_PENCILS)]
    with torch.cuda.stream(stream):
        want = [expected_pencil(torch, pen[d], d, dims, a.device) for d in range(3)]
    bad = {"forward": 0, "backward": 0}
    for it in range(args.iters):
        with torch.cuda.stream(stream):
            encode(torch, a, want[0])
            b.fill_(-7.0)
            aux.fill_(0x5B)
        stream.synchronize()
        dist.barrier()
        plan.execute(a, b, Execute.FORWARD, aux)
        stream.synchronize()
        dist.barrier()
        with torch.cuda.stream(stream):
            okf = holds(torch, b, want[2])
            a.fill_(-7.0)
            aux.fill_(0x5B)
        stream.synchronize()
        dist.barrier()
        plan.execute(b, a, Execute.BACKWARD, aux)
        stream.synchronize()
        dist.barrier()
        with torch.cuda.stream(stream):
            okb = holds(torch, a, want[0])
        bad["forward"] += 0 if okf else 1
        bad["backward"] += 0 if okb else 1
    t = torch.tensor([bad["forward"], bad["backward"]], dtype=torch.float64)
    if not shared:
        t = t.cuda()
    dist.all_reduce(t)
    if rank == 0:
        print(json.dumps({"what": "pair pipeline stress", "n_gpus": world, "n": args.n, "iters": args.iters,
                          "exchange_form": plan.exchange_form(2), "pipelined_stages": plan.overlapped_stages,
                          "mismatching_rank_calls": {"forward": int(t[0]), "backward": int(t[1])},
                          "switches": {k: v for k, v in os.environ.items() if k.startswith("DTFFTB_")},
                          "peer_error": plan.peer_error()}), flush=True)
    for x in bufs + [aux_buf]:
        plan.mem_free(x)
    plan.destroy()
    Config()._commit()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if (t[0] + t[1]) > 0 else 0)


if __name__ == "__main__":
    main()
