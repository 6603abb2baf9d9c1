"""Kernel micro-benchmark / tile sweep (run on the GPU box).  Prints one JSON line per
configuration: achieved GB/s = 2 * elements * element_bytes / CUDA-event time."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_kernels as R  # noqa: E402  (A/B baseline only)
from dtfft_b200.kernel import (KERNEL_PERMUTE_BACKWARD, KERNEL_PERMUTE_BACKWARD_START, KERNEL_PERMUTE_FORWARD,  # noqa: E402
                               KERNEL_UNPACK, KERNEL_PERMUTE_BACKWARD_END, Kernel)

TILES = [(1, 1, 8), (2, 1, 8), (1, 2, 8), (2, 2, 8), (2, 2, 16), (1, 4, 16), (4, 1, 16)]


def time_ms(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def quick(args):
    n = args.n
    N = n ** 3
    for es in (16, 8, 4):
        a = torch.empty(N * es, dtype=torch.uint8, device="cuda")
        a.random_(0, 255)
        b = torch.empty_like(a)
        gb = 2 * N * es / 1e9
        for kt, name in ((KERNEL_PERMUTE_FORWARD, "forward"), (KERNEL_PERMUTE_BACKWARD, "backward"),
                         (KERNEL_PERMUTE_BACKWARD_START, "backward_start")):
            k = Kernel().create([n, n, n], 0, es, kt)
            ms = time_ms(lambda: k.execute(a, b), warmup=5, iters=20)
            print(json.dumps({"what": name, "es": es, "n": n, "ms": ms, "gbs": gb / ms * 1e3}), flush=True)
            k.destroy()
        P = 8
        nxx = n // P
        nd = np.zeros((P, 5), dtype=np.int32)
        for i in range(P):
            nd[i] = (nxx, n, n, i * nxx * n * n, i * nxx)
        for kt, name in ((KERNEL_UNPACK, "unpack8"), (KERNEL_PERMUTE_BACKWARD_END, "backward_end8")):
            k = Kernel().create([n, n, n], 0, es, kt, nd)
            ms = time_ms(lambda: k.execute(a, b), warmup=5, iters=20)
            print(json.dumps({"what": name, "es": es, "n": n, "ms": ms, "gbs": gb / ms * 1e3}), flush=True)
            k.destroy()
        del a, b
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--out", default="gpurun_out/kbench.jsonl")
    ap.add_argument("--quick", action="store_true",
                    help="default tile / grid of every kernel only (A/B of process-wide switches such as DTFFTB_TILE)")
    args = ap.parse_args()
    if args.quick:
        return quick(args)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    f = open(args.out, "a")

    def emit(rec):
        line = json.dumps(rec)
        print(line, flush=True)
        f.write(line + "\n")
        f.flush()

    n = args.n
    N = n ** 3
    for es in (16, 8, 4):
        a = torch.empty(N * es, dtype=torch.uint8, device="cuda")
        a.random_(0, 255)
        b = torch.empty_like(a)
        gb = 2 * N * es / 1e9
        ms = time_ms(lambda: b.copy_(a))
        emit({"what": "torch_copy", "es": es, "n": n, "ms": ms, "gbs": gb / ms * 1e3})
        for kt, name in ((KERNEL_PERMUTE_FORWARD, "forward"), (KERNEL_PERMUTE_BACKWARD, "backward"),
                         (KERNEL_PERMUTE_BACKWARD_START, "backward_start")):
            k = Kernel().create([n, n, n], 0, es, kt)
            for tile in TILES:
                for gm in (4, 16, 64):
                    os.environ["DTFFTB_GRID_MULT"] = str(gm)
                    k.set_tile(*tile)
                    ms = time_ms(lambda: k.execute(a, b))
                    emit({"what": name, "es": es, "n": n, "tile": tile, "grid_mult": gm, "ms": ms, "gbs": gb / ms * 1e3})
            k.destroy()
            # the kernel to beat: the reference's generated kernel, best of its candidate configs
            if R.available():
                best = None
                for (t, r) in R.configs():
                    if not R.valid_config(es, t, r):
                        continue
                    ms = time_ms(lambda: R.launch(kt, [n, n, n], es, t, r, b.data_ptr(), a.data_ptr()), iters=5)
                    rec = {"what": "REF_" + name, "es": es, "n": n, "tile": [t, r], "ms": ms, "gbs": gb / ms * 1e3}
                    emit(rec)
                    best = rec if best is None or ms < best["ms"] else best
                emit(dict(best, what="REF_BEST_" + name))
        # multi-peer unpack / backward_end as produced by an 8-rank slab transposition
        P = 8
        for kt, name in ((KERNEL_UNPACK, "unpack8"), (KERNEL_PERMUTE_BACKWARD_END, "backward_end8")):
            nxx = n // P
            nd = np.zeros((P, 5), dtype=np.int32)
            for i in range(P):
                nd[i] = (nxx, n, n, i * nxx * n * n, i * nxx)
            if name == "backward_end8":
                pass
            if R.available():
                per = {KERNEL_UNPACK: 5, KERNEL_PERMUTE_BACKWARD_END: 11}[kt]
                for (t, r) in ((32, 8), (32, 4), (64, 8), (16, 16)):
                    if not R.valid_config(es, t, r):
                        continue
                    def ref_all():
                        for i in range(P):
                            R.launch(per, [n, n, n], es, t, r, b.data_ptr(), a.data_ptr(), nd[i])
                    ms = time_ms(ref_all, iters=5)
                    emit({"what": "REF_" + name, "es": es, "n": n, "tile": [t, r], "ms": ms, "gbs": gb / ms * 1e3})
            for gm in (1, 4, 16):
                os.environ["DTFFTB_GRID_MULT"] = str(gm)
                k = Kernel().create([n, n, n], 0, es, kt, nd)
                ms = time_ms(lambda: k.execute(a, b))
                emit({"what": name, "es": es, "n": n, "grid_mult": gm, "ms": ms, "gbs": gb / ms * 1e3, "info": k.info()})
                k.destroy()
        del a, b
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
