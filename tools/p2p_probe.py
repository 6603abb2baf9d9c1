#!/usr/bin/env python
"""Measured NVLink ceiling of the box: copy-engine all-to-all between all visible GPUs from ONE
process (cudaMemcpyPeerAsync through torch), as the denominator next to the nominal
900 GB/s/direction when judging the NVLINK_FUSED / NCCL exchange of bench.py.

    python tools/p2p_probe.py [--mib 256]   -> one JSON line

Every GPU sends `mib/(N-1)`-sized chunks to every other GPU at the same time (what one
transposition's exchange does); reported = bytes leaving one GPU / time, max time over GPUs."""
import argparse
import json

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256, help="payload leaving each GPU per round")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    n = torch.cuda.device_count()
    if n < 2:
        print(json.dumps({"error": "needs >= 2 GPUs"}))
        return
    chunk = (args.mib << 20) // (n - 1)
    src = [torch.empty(chunk * (n - 1), dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
    dst = [torch.empty(chunk * (n - 1), dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
    for i in range(n):
        for j in range(n):
            if i != j:
                assert torch.cuda.can_device_access_peer(i, j)
    streams = [[torch.cuda.Stream(device=i) for _ in range(n)] for i in range(n)]

    def round_():
        for i in range(n):
            k = 0
            for j in range(n):
                if i == j:
                    continue
                # slot of sender i in receiver j
                slot = i if i < j else i - 1
                with torch.cuda.device(i), torch.cuda.stream(streams[i][j]):
                    dst[j][slot * chunk:(slot + 1) * chunk].copy_(src[i][k * chunk:(k + 1) * chunk], non_blocking=True)
                k += 1

    def sync():
        for i in range(n):
            torch.cuda.synchronize(i)

    for _ in range(3):
        round_()
    sync()
    import time

    t0 = time.perf_counter()
    for _ in range(args.iters):
        round_()
    sync()
    dt = (time.perf_counter() - t0) / args.iters
    out_bytes = chunk * (n - 1)
    # one direction only, one pair: the unidirectional link peak
    with torch.cuda.device(0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        big_s = torch.empty(args.mib << 20, dtype=torch.uint8, device="cuda:0")
        big_d = torch.empty(args.mib << 20, dtype=torch.uint8, device="cuda:1")
        for _ in range(3):
            big_d.copy_(big_s, non_blocking=True)
        torch.cuda.synchronize(0)
        e0.record()
        for _ in range(args.iters):
            big_d.copy_(big_s, non_blocking=True)
        e1.record()
        torch.cuda.synchronize(0)
        uni = (args.mib << 20) / (e0.elapsed_time(e1) / args.iters * 1e-3) / 1e9
    print(json.dumps({"n_gpus": n, "bytes_out_per_gpu": out_bytes, "all_to_all_ms": dt * 1e3,
                      "all_to_all_GBps_out_per_gpu": out_bytes / dt / 1e9,
                      "pair_0_to_1_unidirectional_GBps": uni,
                      "how": "cudaMemcpyPeerAsync (copy engines), all pairs concurrently; wall clock over "
                             f"{args.iters} rounds after 3 warm-ups"}))


if __name__ == "__main__":
    main()
