#!/usr/bin/env python
"""BASELINE.json configs at FULL size on N GPUs (torchrun, one rank per GPU): parity through
size-independent properties + timing of every backend, with and without stage overlap.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/configs_bench.py [--configs c2fft,c3,c4,c5] [--backends nccl,nccl_pipe,nvlink] \
        [--overlap 1,4] [--iters 10] [--scale 1.0]

Configs (BASELINE.json `configs`):
  c2fft  3D C2C fp64 512^3 with the cuFFT executor (the bench.py workload plus FFTs)
  c3     3D R2C fp32 1024^3 pencils, cuFFT
  c4     2D C2C fp64 16384^2 slab, cuFFT (single exchange)
  c5     3D R2R fp64 768x512x1024 bricks -> pencils with uneven non-power-of-two cuts (transpose-only:
         cuFFT has no r2r, like the reference's cuFFT executor)
Parity at full size:
  * transpose-only (c5): every element carries its global linear index; after the forward execute
    every rank checks ON DEVICE that its output box holds exactly the indices of the destination
    layout (bit-exact, every element), and that backward returns the input bit for bit;
  * FFT configs: a plane wave exp(2 pi i k.r / N) must transform to a single spike of height
    prod(N) at k (relative L2 error against the exact spectrum, tolerance 1e-12 fp64 / 1e-5 fp32),
    and backward(forward(x)) / prod(N) == x for uniform random x within 5 log2(N) 2 eps
    (reference: tests/test_utils.F90:96,107).
One JSON line per (config, backend, overlap) on rank 0."""
import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from dtfft_b200.comm import TorchComm
    from dtfft_b200.plan import (Backend, Config, Execute, Executor, Layout, Pencil, PlanC2C, PlanR2C, PlanR2R, Precision)

    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2fft,c3,c4,c5")
    ap.add_argument("--backends", default="nccl,nccl_pipe,nvlink")
    ap.add_argument("--overlap", default="1,4")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every extent (smoke runs)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = TorchComm()
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(v):
        t = torch.as_tensor(v, device=dev, dtype=torch.float64).clone()
        if world > 1:
            dist.all_reduce(t)
        return t

    names = {"nccl": Backend.NCCL, "nccl_pipe": Backend.NCCL_PIPELINED, "nvlink": Backend.NVLINK_FUSED}
    backends = [("single", Backend.NONE)] if world == 1 else [(b, names[b]) for b in args.backends.split(",")]
    overlaps = [int(x) for x in args.overlap.split(",")]
    sc = lambda n: max(8, int(round(n * args.scale)))
    lines = []

    def emit(d):
        lines.append(d)
        if rank == 0:
            print(json.dumps(d), flush=True)

    def timed(plan, stream, fn, iters):
        for _ in range(args.warmup):
            fn()
        stream.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        stream.synchronize()
        barrier()
        return allmax(e0.elapsed_time(e1) / iters)

    def grid_index(starts, counts, dims, device):
        """Global linear index x + Nx (y + Ny z) of every element of a local box given in NATURAL
        (x, y, z) order, returned as a tensor shaped [cz, cy, cx] (x fastest in memory)."""
        nd = len(dims)
        ax = [torch.arange(starts[d], starts[d] + counts[d], device=device, dtype=torch.float64) for d in range(nd)]
        if nd == 2:
            return ax[0][None, :] + dims[0] * ax[1][:, None]
        return ax[0][None, None, :] + dims[0] * (ax[1][None, :, None] + dims[1] * ax[2][:, None, None])

    # ------------------------------------------------------------------ FFT configs
    def run_fft_config(name, cls, dims, prec):
        nd = len(dims)
        cdt = torch.complex128 if prec == Precision.DOUBLE else torch.complex64
        rdt = torch.float64 if prec == Precision.DOUBLE else torch.float32
        eps = torch.finfo(rdt).eps
        tol = 1e-12 if prec == Precision.DOUBLE else 1e-5
        is_r2c = cls is PlanR2C
        for bname, backend in backends:
            for ov in (overlaps if backend == Backend.NVLINK_FUSED else [1]):
                stream = torch.cuda.Stream()
                cfg = Config(backend=backend, stream=stream, enable_z_slab=(name == "c4"))
                pcomm = comm
                if name == "c3" and world > 1:  # pencil grid of SURVEY.md section 8: 1x2x1 / 1x2x2 / 1x4x2
                    p1 = {2: 2, 4: 2, 8: 4}.get(world, world)
                    pcomm = TorchComm(cart_dims=[1, p1, world // p1])
                plan = cls(dims, comm=pcomm, precision=prec, executor=Executor.CUFFT, config=cfg)
                if ov > 0:  # 0 = the plan's own choice (rule at ESTIMATE, timed at MEASURE)
                    plan.set_overlap(ov)
                ins, inc, outs, outc, alloc = plan.local_sizes
                nbytes = plan.alloc_bytes
                bufs = [plan.mem_alloc(nbytes) for _ in range(3)]
                a, b, c = (torch.as_tensor(x, device="cuda") for x in bufs)
                n_in = int(np.prod(inc))
                n_out = int(np.prod(outc))
                in_dt = rdt if is_r2c else cdt
                x = a.view(in_dt)[:n_in]
                # ---- plane wave -> spike -------------------------------------------------------
                k = [3, 5, 7][:nd]
                with torch.cuda.stream(stream):
                    ph = torch.zeros([inc[d] for d in range(nd - 1, -1, -1)], device=dev, dtype=torch.float64)
                    for d in range(nd):
                        shape = [1] * nd
                        shape[nd - 1 - d] = inc[d]
                        idx = torch.arange(ins[d], ins[d] + inc[d], device=dev, dtype=torch.float64).reshape(shape)
                        ph = ph + (2 * math.pi * k[d] / dims[d]) * idx
                    if is_r2c:
                        x.copy_(torch.cos(ph).to(rdt).reshape(-1))
                    else:
                        x.copy_(torch.polar(torch.ones_like(ph), ph).to(cdt).reshape(-1))
                    del ph
                stream.synchronize()
                barrier()
                plan.execute(a, b, Execute.FORWARD)
                stream.synchronize()
                # output box = Z pencil (Y in 2-D) in ITS axis order: local axis j is natural axis perm[j]
                perm = [1, 0] if nd == 2 else [2, 0, 1]
                spec = b.view(cdt)[:n_out].reshape([outc[j] for j in range(nd - 1, -1, -1)])
                N = float(np.prod(dims))
                want_norm2 = N * N if not is_r2c else (N / 2) ** 2  # r2c keeps only the +k spike
                with torch.cuda.stream(stream):
                    tot = (spec.abs().double() ** 2).sum()
                    # remove the expected spike, what is left is the error
                    loc = [k[perm[j]] - outs[j] for j in range(nd)]
                    err2 = tot.clone()
                    inside = all(0 <= loc[j] < outc[j] for j in range(nd))
                    if inside:
                        pos = tuple(loc[j] for j in range(nd - 1, -1, -1))
                        v = spec[pos]
                        amp = N if not is_r2c else N / 2
                        err2 = err2 - v.abs().double() ** 2 + (v.to(torch.complex128) - amp).abs() ** 2
                stream.synchronize()
                rel = float(torch.sqrt(allsum(err2.reshape(1))[0] / want_norm2))
                ok_spike = rel <= tol
                # ---- round trip on random data -------------------------------------------------
                with torch.cuda.stream(stream):
                    g = torch.Generator(device=dev)
                    g.manual_seed(1234 + rank)
                    if is_r2c:
                        x.uniform_(0, 1, generator=g)
                    else:
                        torch.view_as_real(x).uniform_(0, 1, generator=g)
                    keep = x.clone()
                stream.synchronize()
                barrier()
                plan.execute(a, b, Execute.FORWARD)
                plan.execute(b, c, Execute.BACKWARD)
                stream.synchronize()
                with torch.cuda.stream(stream):
                    back = c.view(in_dt)[:n_in] / N
                    maxerr = (back - keep).abs().max()
                stream.synchronize()
                maxerr = allmax(float(maxerr))
                bound = 5 * math.log2(N) * 2 * eps
                ok_rt = maxerr <= bound
                stages = plan.overlapped_stages
                del keep, back, spec
                # ---- timing: forward + backward ------------------------------------------------
                def cyc():
                    plan.execute(a, b, Execute.FORWARD)
                    plan.execute(b, c, Execute.BACKWARD)

                ms = timed(plan, stream, cyc, args.iters)
                emit({"config": name, "dims": dims, "n_gpus": world, "grid": plan.grid_dims, "backend": plan.backend.name,
                      "overlap_chunks": plan.overlap_chunks, "overlap_requested": ov, "overlapped_stages_per_execute": stages, "fwd_bwd_ms": ms,
                      "spike_rel_l2": rel, "spike_ok": ok_spike, "round_trip_max_err": maxerr, "round_trip_bound": bound,
                      "round_trip_ok": ok_rt, "z_slab": plan.z_slab_enabled, "graph_replays": plan.graph_replays, "peer_error": plan.peer_error()})
                assert ok_spike and ok_rt, (name, bname, ov, rel, maxerr)
                for x_ in bufs:
                    plan.mem_free(x_)
                plan.destroy()
                del a, b, c, x
                torch.cuda.empty_cache()

    # ------------------------------------------------------------------ C5: bricks, transpose-only
    def brick_boxes(dims):
        """Brick grid 2 x 2 x 2 (8 ranks), 2 x 2 x 1 (4) or 2 x 1 x 1 (2) with uneven cuts in the
        proportions of SURVEY.md section 8: x {300,468}/768, y {200,312}/512, z {500,524}/1024."""
        gx = 2
        gy = 2 if world >= 4 else 1
        gz = 2 if world >= 8 else 1
        if gx * gy * gz != world:
            return None

        def cuts(n, g, frac):
            if g == 1:
                return [n]
            first = int(round(n * frac))
            return [first, n - first]

        cx, cy, cz = cuts(dims[0], gx, 300 / 768), cuts(dims[1], gy, 200 / 512), cuts(dims[2], gz, 500 / 1024)
        boxes = []
        for kz in range(gz):
            for jy in range(gy):
                for ix in range(gx):
                    boxes.append(([sum(cx[:ix]), sum(cy[:jy]), sum(cz[:kz])], [cx[ix], cy[jy], cz[kz]]))
        return boxes

    def run_c5(dims):
        boxes = brick_boxes(dims)
        if boxes is None:
            if rank == 0:
                print(json.dumps({"config": "c5", "skipped": f"needs 2, 4 or 8 ranks, got {world}"}), flush=True)
            return
        for bname, backend in backends:
            stream = torch.cuda.Stream()
            cfg = Config(backend=backend, reshape_backend=backend, stream=stream, enable_z_slab=False,
                         enable_fourier_reshape=True)
            plan = PlanR2R(Pencil(*boxes[rank]), comm=comm, precision=Precision.DOUBLE, config=cfg)
            gd = plan.dims
            assert list(gd) == list(dims), (gd, dims)
            nbytes = plan.alloc_bytes
            bufs = [plan.mem_alloc(nbytes) for _ in range(3)]
            a, b, c = (torch.as_tensor(x, device="cuda").view(torch.float64) for x in bufs)
            aux = plan.mem_alloc(plan.aux_bytes) if plan.aux_bytes else None
            b1 = plan.get_pencil(Layout.X_BRICKS)
            b2 = plan.get_pencil(Layout.Z_BRICKS)
            n_in, n_out = b1.size, b2.size
            with torch.cuda.stream(stream):
                a.fill_(-1.0)
                a[:n_in].copy_(grid_index(b1.starts, b1.counts, dims, dev).reshape(-1))
                b.fill_(-2.0)
                c.fill_(-3.0)
            stream.synchronize()
            barrier()
            plan.execute(a, b, Execute.FORWARD, aux)
            stream.synchronize()
            # Z bricks are stored (z, x, y): get_pencil reports starts / counts in that local order
            with torch.cuda.stream(stream):
                nat_s = [b2.starts[1], b2.starts[2], b2.starts[0]]
                nat_c = [b2.counts[1], b2.counts[2], b2.counts[0]]
                want = grid_index(nat_s, nat_c, dims, dev)  # [z][y][x]
                got = b[:n_out].reshape(b2.counts[2], b2.counts[1], b2.counts[0])  # [y][x][z], z fastest
                ok_fwd = bool(torch.equal(got, want.permute(1, 2, 0)))
                del want
            # the forward may have destroyed `a` (the reference's contract): refill it
            with torch.cuda.stream(stream):
                a[:n_in].copy_(grid_index(b1.starts, b1.counts, dims, dev).reshape(-1))
            plan.execute(b, c, Execute.BACKWARD, aux)
            stream.synchronize()
            with torch.cuda.stream(stream):
                ok_bwd = bool(torch.equal(c[:n_in], a[:n_in]))
            stream.synchronize()
            ok = allsum([float(ok_fwd), float(ok_bwd)])
            ok_fwd, ok_bwd = bool(ok[0] == world), bool(ok[1] == world)

            def cyc():
                plan.execute(a, b, Execute.FORWARD, aux)
                plan.execute(b, c, Execute.BACKWARD, aux)

            ms = timed(plan, stream, cyc, args.iters)
            plan.execute(a, b, Execute.FORWARD, aux)
            st = plan.stats()
            stream.synchronize()
            emit({"config": "c5", "dims": list(dims), "n_gpus": world, "brick_grid": [2, 2 if world >= 4 else 1, 2 if world >= 8 else 1],
                  "grid": plan.grid_dims, "backend": plan.backend.name, "fwd_bwd_ms": ms, "forward_bit_exact": ok_fwd,
                  "backward_bit_exact": ok_bwd, "local_bytes_per_execute": st["local_bytes"],
                  "remote_bytes_per_execute": st["remote_bytes"], "launches_per_execute": st["kernel_launches"],
                  "effective_GBps": 2 * 2 * st["local_bytes"] * world / (ms * 1e-3) / 1e9,
                  "graph_replays": plan.graph_replays, "peer_error": plan.peer_error()})
            assert ok_fwd and ok_bwd, ("c5", bname)
            for x_ in bufs + ([aux] if aux is not None else []):
                plan.mem_free(x_)
            plan.destroy()
            del a, b, c
            torch.cuda.empty_cache()

    todo = args.configs.split(",")
    if "c2fft" in todo:
        run_fft_config("c2fft", PlanC2C, [sc(512)] * 3, Precision.DOUBLE)
    if "c3" in todo:
        run_fft_config("c3", PlanR2C, [sc(1024)] * 3, Precision.SINGLE)
    if "c4" in todo:
        run_fft_config("c4", PlanC2C, [sc(16384)] * 2, Precision.DOUBLE)
    if "c5" in todo:
        run_c5([sc(768), sc(512), sc(1024)])
    Config()._commit()
    if rank == 0 and args.out:
        with open(args.out, "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
