"""Worker of tests/test_multi_gpu.py: one process per GPU under torchrun (NCCL).  For every
backend (NCCL, NCCL_PIPELINED, NVLINK_FUSED) checks, through the public plan API:
  * every transposition against the global-array truth of the host MPI-datatype path (bit-exact);
  * C2C / R2C FFT forward spectra against numpy.fft and the backward round trip;
  * brick <-> pencil reshapes with uneven, non-power-of-two cuts (bit-exact)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from dtfft_b200.comm import TorchComm
    from dtfft_b200.plan import (Backend, Config, Effort, Execute, Executor, Layout, Pencil, PlanC2C, PlanR2C, PlanR2R,
                                 Precision, Reshape)
    from oracle import layout as L
    from oracle import pipeline as P

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # DTFFTB_ALLOW_SHARED_DEVICE=1 (tests/test_zz_shared_device_gpu.py): every rank on cuda:0, time-slicing the
    # one GPU of the driver's box; NCCL refuses that, so the metadata plumbing is gloo and only the
    # NVLINK_FUSED backend (cudaIpc + device barriers, both fine on one device) is exercised
    shared = os.environ.get("DTFFTB_ALLOW_SHARED_DEVICE", "0") == "1"
    if shared:
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = TorchComm()

    def allreduce_sum(np_vec):
        tt = torch.from_numpy(np.asarray(np_vec, dtype=np.float64))
        if not shared:
            tt = tt.cuda()
        dist.all_reduce(tt)
        return tt.cpu().numpy()
    lay = [Layout.X_PENCILS, Layout.Y_PENCILS, Layout.Z_PENCILS]

    def oracle_pencil(p):
        return L.Pencil(p.dim, p.starts, p.counts)

    def dev_buf(plan, nbytes, np_src=None):
        buf = plan.mem_alloc(nbytes)
        t = torch.as_tensor(buf, device="cuda")
        t.fill_(0xAB)
        if np_src is not None:
            h = torch.from_numpy(np.ascontiguousarray(np_src).view(np.uint8).copy())
            t[: h.numel()] = h.cuda()
        torch.cuda.synchronize()
        return buf, t

    def sync(plan):
        torch.cuda.ExternalStream(plan.stream).synchronize()

    def host(t, dtype, n):
        return t.cpu().numpy().view(dtype)[:n]

    backends = [Backend.NVLINK_FUSED] if shared else [Backend.NCCL, Backend.NCCL_PIPELINED, Backend.NVLINK_FUSED]
    if os.environ.get("DTFFTB_TEST_BACKENDS"):  # e.g. "NVLINK_FUSED": shorter runs on large boxes
        backends = [Backend[x] for x in os.environ["DTFFTB_TEST_BACKENDS"].split(",")]
    checked = 0
    for backend in backends:
        for dims, z_slab in (([64, 48, 40], False), ([129, 99, 33], False), ([40, 33, 96], True), ([90, 57], False)):
            nd = len(dims)
            cfg = Config(backend=backend, enable_z_slab=z_slab)
            plan = PlanC2C(dims, comm=comm, config=cfg)
            assert plan.backend == backend, (plan.backend, backend)
            G = P.global_array(dims, np.complex128, kind="random")
            pencils = [oracle_pencil(plan.get_pencil(lay[d])) for d in range(nd)]
            ttypes = [1, -1] if nd == 2 else [1, -1, 2, -2] + ([3, -3] if plan.z_slab_enabled else [])
            ab, at = dev_buf(plan, plan.alloc_bytes)
            bb, bt = dev_buf(plan, plan.alloc_bytes)
            for t in ttypes:
                si, ri = L.transpose_pencil_ids(t)
                src = P.pencil_slice(G, pencils[si])
                want = P.pencil_slice(G, pencils[ri])
                at.fill_(0xAB)
                bt.fill_(0xAB)
                at[: src.nbytes] = torch.from_numpy(src.view(np.uint8).copy()).cuda()
                torch.cuda.synchronize()
                dist.barrier()
                plan.transpose(at, bt, t)
                sync(plan)
                got = host(bt, np.complex128, want.size)
                assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), (backend.name, dims, t, rank)
                st = plan.stats()
                assert st["kernel_launches"] >= 1
                checked += 1
            assert plan.peer_error() == 0
            plan.mem_free(ab)
            plan.mem_free(bb)
            plan.destroy()

        # FFT parity: C2C fp64 and R2C fp32 (tolerances of BASELINE.json north_star)
        # ([64, 64, 64] C2C fp64 on pencils is BASELINE.json configs[0]; with 4 ranks sharing one GPU
        #  -- tests/test_zz_shared_device_gpu.py -- it runs at the reference's own 4 ranks on the driver's box)
        for cls, dims, prec, rdt, cdt, tol in ((PlanC2C, [64, 64, 64], Precision.DOUBLE, np.complex128, np.complex128, 1e-12),
                                               (PlanC2C, [64, 48, 40], Precision.DOUBLE, np.complex128, np.complex128, 1e-12),
                                               (PlanR2C, [66, 40, 36], Precision.SINGLE, np.float32, np.complex64, 1e-5),
                                               (PlanC2C, [64, 96], Precision.DOUBLE, np.complex128, np.complex128, 1e-12)):
            # NVLINK_FUSED also runs with the FFT <-> exchange stage overlap (3 uneven chunks)
            for overlap in ([1, 3] if backend == Backend.NVLINK_FUSED else [1]):
                nd = len(dims)
                cfg = Config(backend=backend, enable_z_slab=False)
                plan = cls(dims, comm=comm, precision=prec, executor=Executor.CUFFT, config=cfg)
                plan.set_overlap(overlap)
                G = P.global_array(dims, rdt, kind="random")
                ins, inc, outs, outc, alloc = plan.local_sizes
                xin = L.Pencil(1, ins, inc)
                x = P.pencil_slice(G, xin)
                rev = tuple(range(nd - 1, -1, -1))
                if cls is PlanR2C:
                    spec = np.fft.rfftn(G.astype(np.float64).transpose(rev)).transpose(rev)
                else:
                    spec = np.fft.fftn(G.astype(np.complex128))
                outp = L.Pencil(nd, outs, outc)
                want = P.pencil_slice(np.asfortranarray(spec), outp)
                ab, at = dev_buf(plan, plan.alloc_bytes, x)
                bb, bt = dev_buf(plan, plan.alloc_bytes)
                cb, ct = dev_buf(plan, plan.alloc_bytes)
                dist.barrier()
                plan.execute(at, bt, Execute.FORWARD)
                sync(plan)
                got = host(bt, cdt, want.size).astype(np.complex128)
                tt = allreduce_sum([np.linalg.norm(got - want) ** 2, np.linalg.norm(want) ** 2])
                rel = float(np.sqrt(tt[0] / tt[1]))
                assert rel <= tol, (backend.name, cls.__name__, rel)
                # the third call with the same buffers replays the CUDA graph captured during the second
                if backend == Backend.NVLINK_FUSED:
                    for _ in range(2):
                        at[: x.nbytes] = torch.from_numpy(np.ascontiguousarray(x).view(np.uint8).copy()).cuda()
                        bt.fill_(0xAB)
                        torch.cuda.synchronize()
                        dist.barrier()
                        plan.execute(at, bt, Execute.FORWARD)
                        sync(plan)
                    if plan.overlapped_stages == 0:  # schedules with overlapped stages stay eager
                        assert plan.graph_replays >= 1, plan.graph_replays
                    else:
                        assert plan.graph_replays == 0
                    got3 = host(bt, cdt, want.size).astype(np.complex128)
                    assert np.array_equal(got3, got), "graph replay differs from the eager run"
                if overlap > 1:  # at least the first FFT -> transposition stage ran chunked
                    assert plan.overlapped_stages >= 1, (cls.__name__, dims, plan.overlapped_stages)
                else:
                    assert plan.overlapped_stages == 0
                plan.execute(bt, ct, Execute.BACKWARD)
                sync(plan)
                back = host(ct, rdt, x.size).astype(np.complex128) / np.prod(dims)
                eps = np.finfo(np.float64 if prec == Precision.DOUBLE else np.float32).eps
                assert np.max(np.abs(back - x)) <= 5 * np.log2(float(np.prod(dims))) * 2 * eps
                for b_ in (ab, bb, cb):
                    plan.mem_free(b_)
                plan.destroy()
                checked += 1

        # bricks -> pencils -> bricks with uneven cuts (needs an even number of ranks)
        # NCCL backends: once with the reference's pack-free / unpack-free shortcuts
        # (reshape_handle_generic.F90:711-746; the Z reshapes always take them, the X reshapes when
        # the bricks are split along z) and once with the plain three-step schedule
        for shortcuts in (("1", "0") if backend != Backend.NVLINK_FUSED and world % 2 == 0 else ("1",)):
            if world % 2:
                break
            os.environ["DTFFTB_RESHAPE_SHORTCUTS"] = shortcuts
            # brick grid 2 x ny x nz; with nz = 2 the Z bricks differ from the Z pencils
            nz = 2 if world % 4 == 0 else 1
            ny = world // (2 * nz)
            ycuts = [9 + 2 * j for j in range(ny)]
            cuts = [[30, 34], ycuts, [70] if nz == 1 else [40, 44]]
            edges = [np.concatenate([[0], np.cumsum(c)]) for c in cuts]
            boxes = []
            for k in range(nz):
                for j in range(ny):
                    for i in range(2):
                        boxes.append(([int(edges[0][i]), int(edges[1][j]), int(edges[2][k])],
                                      [int(cuts[0][i]), int(cuts[1][j]), int(cuts[2][k])]))
            cfg = Config(backend=backend, reshape_backend=backend, enable_z_slab=False, enable_fourier_reshape=True)
            plan = PlanR2R(Pencil(*boxes[rank]), comm=comm, config=cfg)
            if world == 2:  # z = 70 > 32 * 2: z split, so the X reshapes are pack-free / unpack-free
                flags = [plan.describe_reshape(t) for t in (Reshape.X_BRICKS_TO_PENCILS, Reshape.X_PENCILS_TO_BRICKS)]
                assert flags[0]["is_pack_free"] and flags[1]["is_unpack_free"]
            dims = plan.dims
            G = P.global_array(dims, np.float64, kind="index")
            b1 = oracle_pencil(plan.get_pencil(Layout.X_BRICKS))
            xp = oracle_pencil(plan.get_pencil(Layout.X_PENCILS))
            zp = oracle_pencil(plan.get_pencil(Layout.Z_PENCILS))
            b2 = oracle_pencil(plan.get_pencil(Layout.Z_BRICKS))
            ab, at = dev_buf(plan, plan.alloc_bytes)
            bb, bt = dev_buf(plan, plan.alloc_bytes)
            for rtype, s_l, d_l in ((Reshape.X_BRICKS_TO_PENCILS, b1, xp), (Reshape.X_PENCILS_TO_BRICKS, xp, b1),
                                    (Reshape.Z_PENCILS_TO_BRICKS, zp, b2), (Reshape.Z_BRICKS_TO_PENCILS, b2, zp)):
                src, want = P.pencil_slice(G, s_l), P.pencil_slice(G, d_l)
                at.fill_(0xAB)
                bt.fill_(0xAB)
                at[: src.nbytes] = torch.from_numpy(src.view(np.uint8).copy()).cuda()
                torch.cuda.synchronize()
                dist.barrier()
                plan.reshape(at, bt, rtype)
                sync(plan)
                got = host(bt, np.float64, want.size)
                assert np.array_equal(got, want), (backend.name, rtype, rank)
                checked += 1
            # whole transpose-only execute through bricks: forward then backward = identity
            src = P.pencil_slice(G, b1)
            at.fill_(0xAB)
            at[: src.nbytes] = torch.from_numpy(src.view(np.uint8).copy()).cuda()
            cb, ct = dev_buf(plan, plan.alloc_bytes)
            torch.cuda.synchronize()
            dist.barrier()
            plan.execute(at, bt, Execute.FORWARD)
            sync(plan)
            want = P.pencil_slice(G, b2)
            assert np.array_equal(host(bt, np.float64, want.size), want), (backend.name, "execute fwd", rank)
            plan.execute(bt, ct, Execute.BACKWARD)
            sync(plan)
            assert np.array_equal(host(ct, np.float64, src.size), src), (backend.name, "execute bwd", rank)
            for b_ in (ab, bb, cb):
                plan.mem_free(b_)
            plan.destroy()
            checked += 1
        os.environ.pop("DTFFTB_RESHAPE_SHORTCUTS", None)

    # Plan::run_transpose_pair on a slab-shaped grid 1 x 1 x P (forward: X->Y local feeding Y->Z; backward: Z->Y exchange
    # feeding Y->X): peer-by-peer pipeline when the exchange runs in its copy-engine form
    experimental = os.environ.get("DTFFTB_TEST_EXPERIMENTAL", "0") == "1"
    if experimental and Backend.NVLINK_FUSED in backends:
        dims = [96, 36, 32 * world + 1]  # x = 96: the copy-engine form cuts the blocks into up to 3 slices
        plan = PlanC2C(dims, comm=comm, config=Config(backend=Backend.NVLINK_FUSED, enable_z_slab=False))
        assert plan.grid_dims == [1, 1, world], plan.grid_dims
        # the peer-by-peer pipeline runs when the exchange is in its copy-engine form (forced by DTFFTB_FUSED_MODE=dma
        # at these sizes); with the direct-store kernel the two transpositions run one after the other
        want_pipelined = 1 if "copy engines" in plan.exchange_form(2)["form"] else 0
        G = P.global_array(dims, np.complex128, kind="random")
        pencils = [oracle_pencil(plan.get_pencil(lay[d])) for d in range(3)]
        x, want = P.pencil_slice(G, pencils[0]), P.pencil_slice(G, pencils[2])
        ab, at = dev_buf(plan, plan.alloc_bytes, x)
        bb, bt = dev_buf(plan, plan.alloc_bytes)
        cb, ct = dev_buf(plan, plan.alloc_bytes)
        for _ in range(2):  # twice: the second call reuses the piece kernels and the barrier epochs
            bt.fill_(0xAB)
            ct.fill_(0xAB)
            torch.cuda.synchronize()
            dist.barrier()
            plan.execute(at, bt, Execute.FORWARD)
            sync(plan)
            assert plan.overlapped_stages == want_pipelined, plan.overlapped_stages
            assert np.array_equal(host(bt, np.complex128, want.size).view(np.uint8), want.view(np.uint8)), ("pair fwd", rank)
            plan.execute(bt, ct, Execute.BACKWARD)
            sync(plan)
            assert plan.overlapped_stages == want_pipelined
            assert np.array_equal(host(ct, np.complex128, x.size).view(np.uint8), x.view(np.uint8)), ("pair bwd", rank)
        assert plan.peer_error() == 0
        for b_ in (ab, bb, cb):
            plan.mem_free(b_)
        plan.destroy()
        checked += 1

    # Drop-in contract (src/dtfft_plan.F90:1769-1795: ANY device pointer is accepted): no backend named, no
    # dtfft_mem_alloc -- plain torch allocations (cudaMalloc segments) through dtfft_transpose / dtfft_execute.
    # The default backend on a peer-reachable box is NVLINK_FUSED; the buffers are published on first use.
    if "DTFFTB_DEFAULT_BACKEND" not in os.environ and "DTFFT_BACKEND" not in os.environ:
        dims = [72, 40, 56]
        plan = PlanC2C(dims, comm=comm, config=Config(enable_z_slab=False))
        assert plan.backend == Backend.NVLINK_FUSED, plan.backend
        G = P.global_array(dims, np.complex128, kind="random")
        pencils = [oracle_pencil(plan.get_pencil(lay[d])) for d in range(3)]
        x, ywant, zwant = (P.pencil_slice(G, pencils[d]) for d in range(3))
        nb = plan.alloc_bytes
        for round_ in range(3):  # round 1, 2: freshly allocated tensors, quite possibly at the addresses of round 0
            at = torch.full((nb,), 0xAB, dtype=torch.uint8, device="cuda")
            bt = torch.full((nb,), 0xAB, dtype=torch.uint8, device="cuda")
            ct = torch.full((nb,), 0xAB, dtype=torch.uint8, device="cuda")
            at[: x.nbytes] = torch.from_numpy(x.view(np.uint8).copy()).cuda()
            torch.cuda.synchronize()
            dist.barrier()
            plan.transpose(at, bt, 1)  # X -> Y into a plain buffer
            sync(plan)
            assert np.array_equal(host(bt, np.complex128, ywant.size).view(np.uint8), ywant.view(np.uint8)), ("drop-in X->Y", rank)
            for rep in range(3):  # eager, captured, replayed
                bt.fill_(0xAB)
                ct.fill_(0xAB)
                torch.cuda.synchronize()
                dist.barrier()
                plan.execute(at, bt, Execute.FORWARD)
                sync(plan)
                assert np.array_equal(host(bt, np.complex128, zwant.size).view(np.uint8), zwant.view(np.uint8)), ("drop-in fwd", rank, rep)
                plan.execute(bt, ct, Execute.BACKWARD)
                sync(plan)
                assert np.array_equal(host(ct, np.complex128, x.size).view(np.uint8), x.view(np.uint8)), ("drop-in bwd", rank, rep)
            assert plan.fallbacks == 0 and plan.peer_error() == 0
            del at, bt, ct
            torch.cuda.synchronize()
            dist.barrier()
            if round_ == 1:
                torch.cuda.empty_cache()  # really returns the segments: the next round maps new allocations
        # memory cudaIpc cannot export (stream-ordered allocations): the call must still work -- on the NCCL stand-in
        import ctypes

        if shared:  # no NCCL among ranks of one device
            plan.destroy()
            checked += 1
        else:
            rt = ctypes.CDLL("libcudart.so.12")
            ptrs = []
            for _ in range(3):
                pv = ctypes.c_void_p(0)
                assert rt.cudaMallocAsync(ctypes.byref(pv), ctypes.c_size_t(nb), ctypes.c_void_p(0)) == 0
                ptrs.append(pv.value)
            assert rt.cudaDeviceSynchronize() == 0
            rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
            xb = np.ascontiguousarray(x).view(np.uint8)
            assert rt.cudaMemcpy(ptrs[0], xb.ctypes.data, xb.nbytes, 1) == 0
            dist.barrier()
            plan.execute(ptrs[0], ptrs[1], Execute.FORWARD)
            plan.execute(ptrs[1], ptrs[2], Execute.BACKWARD)
            sync(plan)
            back = np.empty_like(xb)
            assert rt.cudaMemcpy(back.ctypes.data, ptrs[2], xb.nbytes, 2) == 0
            assert np.array_equal(back, xb), ("fallback round trip", rank)
            zb = np.empty(zwant.nbytes, np.uint8)
            assert rt.cudaMemcpy(zb.ctypes.data, ptrs[1], zb.nbytes, 2) == 0
            assert np.array_equal(zb, zwant.view(np.uint8)), ("fallback fwd", rank)
            assert plan.fallbacks > 0, "stream-ordered memory should have taken the NCCL stand-in"
            dist.barrier()
            plan.destroy()
            for pv in ptrs:
                rt.cudaFreeAsync(ctypes.c_void_p(pv), ctypes.c_void_p(0))
            rt.cudaDeviceSynchronize()
            checked += 1

    # DTFFT_PATIENT: timed backend choice (run_autotune_backend), then a correct transposition
    plan = PlanC2C([128, 64, 96], comm=comm, effort=Effort.PATIENT, config=Config(enable_z_slab=False))
    picked = plan.backend
    assert picked in (Backend.NCCL, Backend.NCCL_PIPELINED, Backend.NVLINK_FUSED)
    G = P.global_array([128, 64, 96], np.complex128)
    xp, yp = oracle_pencil(plan.get_pencil(Layout.X_PENCILS)), oracle_pencil(plan.get_pencil(Layout.Y_PENCILS))
    src, want = P.pencil_slice(G, xp), P.pencil_slice(G, yp)
    ab, at = dev_buf(plan, plan.alloc_bytes, src)
    bb, bt = dev_buf(plan, plan.alloc_bytes)
    dist.barrier()
    plan.transpose(at, bt, 1)
    sync(plan)
    assert np.array_equal(host(bt, np.complex128, want.size).view(np.uint8), want.view(np.uint8))
    plan.mem_free(ab)
    plan.mem_free(bb)
    plan.destroy()

    # DTFFT_EXHAUSTIVE on a brick plan: timed choice of the reshape backend too (autotune_reshape_plan,
    # src/dtfft_reshape_plan.F90:206-222), then a correct reshape
    picked_r = None
    if experimental and world % 2 == 0:
        plan = PlanR2R(Pencil(*boxes[rank]), comm=comm, effort=Effort.EXHAUSTIVE,
                       config=Config(enable_z_slab=False, enable_fourier_reshape=True))
        picked_r = plan.reshape_backend
        assert picked_r in (Backend.NCCL, Backend.NCCL_PIPELINED, Backend.NVLINK_FUSED)
        G = P.global_array(plan.dims, np.float64, kind="index")
        b1 = oracle_pencil(plan.get_pencil(Layout.X_BRICKS))
        xp = oracle_pencil(plan.get_pencil(Layout.X_PENCILS))
        src, want = P.pencil_slice(G, b1), P.pencil_slice(G, xp)
        ab, at = dev_buf(plan, plan.alloc_bytes, src)
        bb, bt = dev_buf(plan, plan.alloc_bytes)
        dist.barrier()
        plan.reshape(at, bt, Reshape.X_BRICKS_TO_PENCILS)
        sync(plan)
        assert np.array_equal(host(bt, np.float64, want.size), want), ("EXHAUSTIVE reshape", rank)
        plan.mem_free(ab)
        plan.mem_free(bb)
        plan.destroy()

    # DTFFT_MEASURE on an FFT plan with the NVLINK_FUSED backend: timed choice of the stage overlap
    # (choose_overlap), then a forward + backward round trip within the reference's tolerance
    dims = [128, 64, 96]
    plan = PlanC2C(dims, comm=comm, effort=Effort.MEASURE, executor=Executor.CUFFT,
                   config=Config(enable_z_slab=False, backend=Backend.NVLINK_FUSED))
    assert plan.overlap_chunks in (1, 4, 8), plan.overlap_chunks
    G = P.global_array(dims, np.complex128, kind="random")
    ins, inc, outs, outc, alloc = plan.local_sizes
    x = P.pencil_slice(G, L.Pencil(1, ins, inc))
    ab, at = dev_buf(plan, plan.alloc_bytes, x)
    bb, bt = dev_buf(plan, plan.alloc_bytes)
    cb, ct = dev_buf(plan, plan.alloc_bytes)
    dist.barrier()
    plan.execute(at, bt, Execute.FORWARD)
    plan.execute(bt, ct, Execute.BACKWARD)
    sync(plan)
    back = host(ct, np.complex128, x.size) / np.prod(dims)
    assert np.max(np.abs(back - x)) <= 5 * np.log2(float(np.prod(dims))) * 2 * np.finfo(np.float64).eps
    tuned = plan.overlap_chunks
    for b_ in (ab, bb, cb):
        plan.mem_free(b_)
    plan.destroy()
    Config()._commit()
    dist.barrier()
    print(f"rank {rank}/{world}: multi-GPU plan checks OK ({checked} cases, PATIENT picked {picked.name}, EXHAUSTIVE reshape backend {picked_r.name if picked_r else '-'}, MEASURE overlap {tuned})", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
