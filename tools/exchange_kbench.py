#!/usr/bin/env python
"""The exchanging transposition of the headline cycle as ONE process over G devices: device d runs the product's
family-T kernel over the G boxes of a Y -> Z transposition on a 1 x 1 x G grid (n^3 elements of 16 bytes), box p
stored straight into device p's Z pencil through peer access.  No barriers, no ranks: this is the kernel alone,
every device sending and receiving at once -- the harness for tile A/Bs and for `ncu` (kernel replay is safe here,
nothing waits on another GPU), e.g.

    ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
        --clock-control none -k regex:transpose_tiles -c 4 python tools/exchange_kbench.py --devices 2 --iters 1 --tiles 1,2,8

Prints one JSON line per tile: time (max over devices), GB/s per direction per GPU of the remote payload."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtfft_b200.kernel import Kernel  # noqa: E402

ES = 16


def boxes_for(d, G, n):
    """Boxes of device d: Y pencil (y: n, z: n/G, x: n), y fastest -> Z pencil of device p (z: n, x: n, y: n/G)."""
    ny, nzl, nx, nyl = n, n // G, n, n // G
    rows = []
    for p in range(G):
        rows.append([nyl, nzl, nx,            # n0 = my y's owned by p, n1 = my z's, n2 = x
                     p * nyl, d * nzl,        # in_off (y start), out_off (z start inside p's pencil)
                     ny, ny * nzl,            # is1 (z stride), is2 (x stride) in the Y pencil
                     n * nx, 1, n])           # os0 (y stride), os1 (z contiguous), os2 (x stride) in the Z pencil
    return rows


def uneven_case(G, zcounts, ny, nx):
    """Y -> Z like config 5 (768 x 512 x 1024 f64 bricks: Y pencils 512 x {250|262} x 384, Z pencils 1024 x 384 x 128):
    the z ranges the ranks own are not multiples of a 128-byte line, so every destination run starts misaligned.
    Returns (boxes per device, local elements per device, Z-pencil elements)."""
    nz = sum(zcounts)
    zstart = [sum(zcounts[:i]) for i in range(G)]
    nyl = ny // G
    out = []
    for d in range(G):
        rows = []
        for p in range(G):
            rows.append([nyl, zcounts[d], nx, p * nyl, zstart[d], ny, ny * zcounts[d], nz * nx, 1, nz])
        out.append(rows)
    return out, [ny * zcounts[d] * nx for d in range(G)], nz * nx * nyl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, default=torch.cuda.device_count())
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tiles", default="1,2,8;2,2,8;2,1,8;1,4,16;1,1,8", help="ka,kb,rows;...")
    ap.add_argument("--check", action="store_true", help="verify the landed Z pencils (index-encoded)")
    ap.add_argument("--uneven", default="", help="z counts per device, e.g. 250,262: config-5-like 8-byte case (512 x z x 384)")
    ap.add_argument("--local", action="store_true", help="--uneven on ONE device: the peers' pencils are local buffers (is the "
                    "access pattern itself slow, or only over NVLink?)")
    args = ap.parse_args()
    if args.uneven:
        return uneven(args)
    G, n = args.devices, args.n
    rt = ctypes.CDLL("libcudart.so.12")
    for d in range(G):
        torch.cuda.set_device(d)
        for p in range(G):
            if p != d:
                rt.cudaDeviceEnablePeerAccess(p, 0)
    n_local = n * n * (n // G)
    src, dst, streams = [], [], []
    for d in range(G):
        torch.cuda.set_device(d)
        # element (y, z, x) of device d's Y pencil carries its global index (y, z0 + z, x) twice
        y = torch.arange(n, device="cuda", dtype=torch.int64)[None, None, :]
        z = (torch.arange(n // G, device="cuda", dtype=torch.int64) + d * (n // G))[None, :, None]
        x = torch.arange(n, device="cuda", dtype=torch.int64)[:, None, None]
        g = (x + n * (y + n * z)).reshape(-1)  # memory order of the Y pencil: y fastest, then z, then x
        s = torch.empty(2 * n_local, dtype=torch.int64, device="cuda")
        s.view(-1, 2)[:, 0] = g
        s.view(-1, 2)[:, 1] = ~g
        src.append(s)
        dst.append(torch.zeros(2 * n_local, dtype=torch.int64, device="cuda"))
        streams.append(torch.cuda.Stream(device=d))
    for tile in args.tiles.split(";"):
        ka, kb, rows = (int(v) for v in tile.split(","))
        kernels = []
        for d in range(G):
            torch.cuda.set_device(d)
            k = Kernel().create_boxes(2, ES, boxes_for(d, G, n), out_bases=[dst[p] for p in range(G)])
            k.set_tile(ka, kb, rows)
            kernels.append(k)

        def launch_all():
            for d in range(G):
                torch.cuda.set_device(d)
                kernels[d].execute_all(src[d], dst[d], streams[d].cuda_stream)

        for _ in range(args.warmup):
            launch_all()
        for d in range(G):
            torch.cuda.synchronize(d)
        ev = []
        for d in range(G):
            torch.cuda.set_device(d)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(streams[d])
            ev.append((e0, e1))
        for _ in range(args.iters):
            launch_all()
        for d in range(G):
            torch.cuda.set_device(d)
            ev[d][1].record(streams[d])
        ms = 0.0
        for d in range(G):
            torch.cuda.synchronize(d)
            ms = max(ms, ev[d][0].elapsed_time(ev[d][1]) / args.iters)
        remote = n_local * ES * (G - 1) // G
        rec = {"what": "Y_TO_Z exchange kernel alone", "devices": G, "n": n, "tile": [32 * ka, 32 * kb], "threads": 32 * rows,
               "ms": ms, "remote_bytes_per_gpu": remote, "GBps_per_direction": remote / (ms * 1e-3) / 1e9,
               "frac_of_900": remote / (ms * 1e-3) / 1e9 / 900.0, "local_hbm_GBps": 2 * n_local * ES / (ms * 1e-3) / 1e9}
        if args.check:
            ok = True
            for p in range(G):
                torch.cuda.set_device(p)
                z = torch.arange(n, device="cuda", dtype=torch.int64)[None, None, :]
                x = torch.arange(n, device="cuda", dtype=torch.int64)[None, :, None]
                y = (torch.arange(n // G, device="cuda", dtype=torch.int64) + p * (n // G))[:, None, None]
                want = (x + n * (y + n * z)).reshape(-1)  # Z pencil: z fastest, then x, then y
                v = dst[p].view(-1, 2)
                ok = ok and bool(torch.equal(v[:, 0], want)) and bool(torch.equal(v[:, 1], ~want))
            rec["landed_bit_exact"] = ok
        print(json.dumps(rec), flush=True)
        for k in kernels:
            k.destroy()


def uneven(args):
    zc = [int(v) for v in args.uneven.split(",")]
    G = len(zc)
    es, ny, nx = 8, 512, 384
    boxes, nloc, nz_elems = uneven_case(G, zc, ny, nx)
    if args.local:
        torch.cuda.set_device(0)
        dsts = [torch.zeros(nz_elems, dtype=torch.int64, device="cuda") for _ in range(G)]
        for d in range(G):
            srcd = torch.arange(nloc[d], dtype=torch.int64, device="cuda")
            for tile in args.tiles.split(";"):
                ka, kb, rows = (int(v) for v in tile.split(","))
                k = Kernel().create_boxes(2, es, boxes[d], out_bases=dsts)
                k.set_tile(ka, kb, rows)
                for _ in range(args.warmup):
                    k.execute_all(srcd, dsts[0])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    k.execute_all(srcd, dsts[0])
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                print(json.dumps({"what": "config-5-like Y_TO_Z kernel of virtual rank %d, every destination LOCAL" % d, "z_counts": zc,
                                  "tile": [32 * ka, 32 * kb], "ms": ms, "hbm_GBps": 2 * nloc[d] * es / (ms * 1e-3) / 1e9,
                                  "align_tiles": os.environ.get("DTFFTB_ALIGN_TILES", "1")}), flush=True)
                k.destroy()
        return
    rt = ctypes.CDLL("libcudart.so.12")
    for d in range(G):
        torch.cuda.set_device(d)
        for p in range(G):
            if p != d:
                rt.cudaDeviceEnablePeerAccess(p, 0)
    src, dst, streams = [], [], []
    for d in range(G):
        torch.cuda.set_device(d)
        src.append(torch.arange(nloc[d], dtype=torch.int64, device="cuda") + (d << 40))
        dst.append(torch.zeros(nz_elems, dtype=torch.int64, device="cuda"))
        streams.append(torch.cuda.Stream(device=d))
    for tile in args.tiles.split(";"):
        ka, kb, rows = (int(v) for v in tile.split(","))
        kernels = []
        for d in range(G):
            torch.cuda.set_device(d)
            k = Kernel().create_boxes(2, es, boxes[d], out_bases=[dst[p] for p in range(G)])
            k.set_tile(ka, kb, rows)
            kernels.append(k)

        def launch_all():
            for d in range(G):
                torch.cuda.set_device(d)
                kernels[d].execute_all(src[d], dst[d], streams[d].cuda_stream)

        for _ in range(args.warmup):
            launch_all()
        for d in range(G):
            torch.cuda.synchronize(d)
        ev = []
        for d in range(G):
            torch.cuda.set_device(d)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(streams[d])
            ev.append((e0, e1))
        for _ in range(args.iters):
            launch_all()
        for d in range(G):
            torch.cuda.set_device(d)
            ev[d][1].record(streams[d])
        ms = 0.0
        for d in range(G):
            torch.cuda.synchronize(d)
            ms = max(ms, ev[d][0].elapsed_time(ev[d][1]) / args.iters)
        remote = max(nloc) * es * (G - 1) // G
        print(json.dumps({"what": "config-5-like Y_TO_Z exchange kernel alone (8-byte elements, 512 x z x 384)", "z_counts": zc,
                          "devices": G, "tile": [32 * ka, 32 * kb], "threads": 32 * rows, "ms": ms,
                          "remote_bytes_per_gpu": remote, "GBps_per_direction": remote / (ms * 1e-3) / 1e9,
                          "run_bytes": [z * es for z in zc], "info": kernels[0].info()}), flush=True)
        for k in kernels:
            k.destroy()


if __name__ == "__main__":
    main()
