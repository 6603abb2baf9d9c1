#!/usr/bin/env python
"""Per-stage profile of BASELINE.json configs C2-C5 at FULL size on N GPUs (torchrun, one rank per GPU):
every transposition / reshape of the plan is (a) checked element for element against the analytic
destination layout of an index-encoded global array (= the reference's host MPI-datatype path,
src/dtfft_reshape_handle_datatype.F90:479-574, which only redistributes) and (b) timed on its own
with CUDA events on the plan stream, max over ranks, and put against its roofline:

    local kernels      2 x local bytes / time   vs MEASURED_PEAKS.json hbm_gbs       
    exchanging stages  remote bytes / time      vs 900 GB/s per direction per GPU (and the 737 GB/s a
                                                   copy engine reaches on this pool)

The whole forward + backward execute is timed too; with the cuFFT executor its time minus the
transpositions' is reported as `fft_ms` (library time, kept apart as SURVEY.md 8d asks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/configs_profile.py [--configs c2,c3,c4,c5] [--backends nvlink,nccl,nccl_pipe] [--iters 10] [--scale 1.0]

One JSON line per (config, backend) on rank 0."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DPERM3 = [[0, 1, 2], [1, 2, 0], [2, 0, 1]]  # local axis j of pencil d is natural axis DPERM[d][j]
DPERM2 = [[0, 1], [1, 0]]


def main():
    import torch
    import torch.distributed as dist

    from dtfft_b200.comm import TorchComm
    from dtfft_b200.plan import (Backend, Config, Execute, Executor, Layout, Pencil, PlanC2C, PlanR2C, PlanR2R, Precision,
                                 Reshape, Transpose)

    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2,c3,c4,c5")
    ap.add_argument("--backends", default="nvlink,nccl,nccl_pipe")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
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
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"

    def barrier():
        if world > 1:
            dist.barrier()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allmin(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t[0])

    names = {"nccl": Backend.NCCL, "nccl_pipe": Backend.NCCL_PIPELINED, "nvlink": Backend.NVLINK_FUSED}
    backends = [("single", Backend.NONE)] if world == 1 else [(b, names[b]) for b in args.backends.split(",")]
    sc = lambda n: max(8, int(round(n * args.scale)))

    def expected(pencil, d, dims, es):
        """Index-encoded content of a local box in ITS memory order, as the integer view of the element."""
        nd = len(dims)
        dperm = (DPERM3 if nd == 3 else DPERM2)[d]
        nat_s, nat_c = [0] * nd, [0] * nd
        for j in range(nd):
            nat_s[dperm[j]], nat_c[dperm[j]] = pencil.starts[j], pencil.counts[j]
        idt = torch.int32 if es == 4 else torch.int64
        ax = [torch.arange(nat_s[k], nat_s[k] + nat_c[k], device=dev, dtype=idt) for k in range(nd)]
        if nd == 2:
            want = ax[0][None, :] + dims[0] * ax[1][:, None]  # [y][x]
            return want.permute(1 - dperm[1], 1 - dperm[0]).reshape(-1)
        want = ax[0][None, None, :] + dims[0] * (ax[1][None, :, None] + dims[1] * ax[2][:, None, None])  # [z][y][x]
        return want.permute(2 - dperm[2], 2 - dperm[1], 2 - dperm[0]).reshape(-1)

    def put(buf, want, es):
        if es == 16:
            v = buf.view(torch.int64)[: 2 * want.numel()].view(-1, 2)
            v[:, 0] = want
            v[:, 1] = ~want
        elif es == 8:
            buf.view(torch.int64)[: want.numel()] = want
        else:
            buf.view(torch.int32)[: want.numel()] = want

    def holds(buf, want, es):
        if es == 16:
            v = buf.view(torch.int64)[: 2 * want.numel()].view(-1, 2)
            return bool(torch.equal(v[:, 0], want)) and bool(torch.equal(v[:, 1], ~want))
        if es == 8:
            return bool(torch.equal(buf.view(torch.int64)[: want.numel()], want))
        return bool(torch.equal(buf.view(torch.int32)[: want.numel()], want))

    def profile_stage(plan, stream, call, a, b, aux, src_pen, src_d, dst_pen, dst_d, dims, es):
        """Parity (index-encoded, every element) then timing of one transposition / reshape a -> b."""
        with torch.cuda.stream(stream):
            a.fill_(0)
            put(a, expected(src_pen, src_d, dims, es), es)
            b.fill_(0x5A)
        stream.synchronize()
        barrier()
        call(a, b, aux)
        stream.synchronize()
        barrier()
        with torch.cuda.stream(stream):
            ok = holds(b, expected(dst_pen, dst_d, dims, es), es)
        stream.synchronize()
        ok = allmin(1.0 if ok else 0.0) == 1.0
        st = plan.stats()
        for _ in range(args.warmup):
            call(a, b, aux)
        stream.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.iters):
            call(a, b, aux)
        e1.record(stream)
        stream.synchronize()
        barrier()
        ms = allmax(e0.elapsed_time(e1) / args.iters)
        local_b, remote_b = allmax(st["local_bytes"]), allmax(st["remote_bytes"])
        rec = {"ms": ms, "parity": "bit-exact" if ok else "MISMATCH", "local_bytes_per_gpu": local_b,
               "remote_bytes_per_gpu": remote_b, "launches": st["kernel_launches"]}
        if remote_b > 0:
            g = remote_b / (ms * 1e-3) / 1e9
            rec.update({"bound": "nvlink", "GBps_per_direction": g, "frac_of_900": g / 900.0, "frac_of_dma_737": g / 737.0})
        else:
            g = 2 * local_b / (ms * 1e-3) / 1e9
            rec.update({"bound": "hbm", "GBps": g, "frac_of_peak": g / peak})
        return rec

    def run(name, make_plan, dims_fourier, es, stages_of, bname, backend):
        stream = torch.cuda.Stream()
        plan = make_plan(backend, stream)
        nbytes = plan.alloc_bytes
        bufs = [plan.mem_alloc(nbytes) for _ in range(3)]
        a, b, c = (torch.as_tensor(x, device="cuda") for x in bufs)
        aux_buf = plan.mem_alloc(plan.aux_bytes) if plan.aux_bytes else None
        aux = torch.as_tensor(aux_buf, device="cuda") if aux_buf is not None else None
        stages = {}
        for sname, call, (sl, sd), (dl, dd) in stages_of(plan):
            stages[sname] = profile_stage(plan, stream, call, a, b, aux, plan.get_pencil(sl), sd, plan.get_pencil(dl), dd,
                                          dims_fourier, es)

        def cyc():
            plan.execute(a, b, Execute.FORWARD, aux)
            plan.execute(b, c, Execute.BACKWARD, aux)

        with torch.cuda.stream(stream):
            a.fill_(0)
        for _ in range(args.warmup):
            cyc()
        stream.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.iters):
            cyc()
        e1.record(stream)
        stream.synchronize()
        barrier()
        fwd_bwd = allmax(e0.elapsed_time(e1) / args.iters)
        rec = {"config": name, "n_gpus": world, "grid": plan.grid_dims, "backend": plan.backend.name, "element_bytes": es,
               "dims": list(plan.dims), "stages": stages, "fwd_bwd_ms": fwd_bwd, "hbm_peak": peak, "hbm_peak_source": peak_src,
               "overlap_chunks": plan.overlap_chunks, "z_slab": plan.z_slab_enabled, "peer_error": plan.peer_error()}
        worst = None
        for k, v in stages.items():
            f = v.get("frac_of_peak", v.get("frac_of_900"))
            if worst is None or f < worst[1]:
                worst = (k, f, v["bound"])
        rec["furthest_below_roofline"] = {"stage": worst[0], "frac": worst[1], "bound": worst[2]} if worst else None
        if rank == 0:
            print(json.dumps(rec), flush=True)
        bad = [k for k, v in stages.items() if v["parity"] != "bit-exact"]
        for x_ in bufs + ([aux_buf] if aux_buf is not None else []):
            plan.mem_free(x_)
        plan.destroy()
        del a, b, c, aux
        torch.cuda.empty_cache()
        assert not bad, (name, bname, bad)

    X, XF, Y, Z, XB, ZB = (Layout.X_PENCILS, Layout.X_PENCILS_FOURIER, Layout.Y_PENCILS, Layout.Z_PENCILS, Layout.X_BRICKS,
                           Layout.Z_BRICKS)

    def tcall(plan, t):
        return lambda a, b, aux: plan.transpose(a, b, t, aux)

    def rcall(plan, t):
        return lambda a, b, aux: plan.reshape(a, b, t, aux)

    def transposes_3d(plan, x_layout):
        if plan.z_slab_enabled:
            return [("X_TO_Z", tcall(plan, Transpose.X_TO_Z), (x_layout, 0), (Z, 2)),
                    ("Z_TO_X", tcall(plan, Transpose.Z_TO_X), (Z, 2), (x_layout, 0))]
        return [("X_TO_Y", tcall(plan, Transpose.X_TO_Y), (x_layout, 0), (Y, 1)),
                ("Y_TO_Z", tcall(plan, Transpose.Y_TO_Z), (Y, 1), (Z, 2)),
                ("Z_TO_Y", tcall(plan, Transpose.Z_TO_Y), (Z, 2), (Y, 1)),
                ("Y_TO_X", tcall(plan, Transpose.Y_TO_X), (Y, 1), (x_layout, 0))]

    todo = args.configs.split(",")
    for bname, backend in backends:
        if "c2" in todo:  # the bench.py workload with its per-stage view (transpose-only, 16-byte elements)
            d = [sc(512)] * 3
            run("c2", lambda be, st: PlanC2C(d, comm=comm, config=Config(backend=be, stream=st, enable_z_slab=False)),
                d, 16, lambda p: transposes_3d(p, X), bname, backend)
        if "c3" in todo:  # 1024^3 R2C fp32: transpositions on 8-byte complex, 513 x 1024 x 1024
            d = [sc(1024)] * 3
            pc = comm
            if world > 1:  # pencil grids of SURVEY.md section 8: 1x2x1 / 1x2x2 / 1x4x2
                p1 = {2: 2, 4: 2, 8: 4}.get(world, world)
                pc = TorchComm(cart_dims=[1, p1, world // p1])
            run("c3", lambda be, st: PlanR2C(d, comm=pc, precision=Precision.SINGLE, executor=Executor.CUFFT,
                                             config=Config(backend=be, stream=st, enable_z_slab=False)),
                [d[0] // 2 + 1, d[1], d[2]], 8, lambda p: transposes_3d(p, XF), bname, backend)
        if "c4" in todo:  # 16384^2 C2C fp64 slab: one exchange each way
            d = [sc(16384)] * 2
            run("c4", lambda be, st: PlanC2C(d, comm=comm, executor=Executor.CUFFT, config=Config(backend=be, stream=st)),
                d, 16, lambda p: [("X_TO_Y", tcall(p, Transpose.X_TO_Y), (X, 0), (Y, 1)),
                                  ("Y_TO_X", tcall(p, Transpose.Y_TO_X), (Y, 1), (X, 0))], bname, backend)
        if "c5" in todo and world in (2, 4, 8):  # bricks with uneven non-power-of-two cuts, 8-byte reals
            d = [sc(768), sc(512), sc(1024)]
            gx, gy, gz = 2, (2 if world >= 4 else 1), (2 if world >= 8 else 1)

            def cuts(n, g, frac):
                if g == 1:
                    return [n]
                first = int(round(n * frac))
                return [first, n - first]

            cx, cy, cz = cuts(d[0], gx, 300 / 768), cuts(d[1], gy, 200 / 512), cuts(d[2], gz, 500 / 1024)
            boxes = []
            for kz in range(gz):
                for jy in range(gy):
                    for ix in range(gx):
                        boxes.append(([sum(cx[:ix]), sum(cy[:jy]), sum(cz[:kz])], [cx[ix], cy[jy], cz[kz]]))

            def c5_stages(p):
                return ([("R_X_BRICKS_TO_PENCILS", rcall(p, Reshape.X_BRICKS_TO_PENCILS), (XB, 0), (X, 0))]
                        + transposes_3d(p, X)
                        + [("R_Z_PENCILS_TO_BRICKS", rcall(p, Reshape.Z_PENCILS_TO_BRICKS), (Z, 2), (ZB, 2)),
                           ("R_Z_BRICKS_TO_PENCILS", rcall(p, Reshape.Z_BRICKS_TO_PENCILS), (ZB, 2), (Z, 2)),
                           ("R_X_PENCILS_TO_BRICKS", rcall(p, Reshape.X_PENCILS_TO_BRICKS), (X, 0), (XB, 0))])

            run("c5", lambda be, st: PlanR2R(Pencil(*boxes[rank]), comm=comm, precision=Precision.DOUBLE,
                                             config=Config(backend=be, reshape_backend=be, stream=st, enable_z_slab=False,
                                                           enable_fourier_reshape=True)),
                d, 8, c5_stages, bname, backend)
    Config()._commit()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
