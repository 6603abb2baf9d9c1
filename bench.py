#!/usr/bin/env python
"""bench.py -- the BASELINE.json metric on B200: 512^3 C2C fp64 forward+backward
transpose-only cycle (X->Y->Z->Y->X), time per cycle and effective GB/s.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N ...             # reference CPU path (port)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
  value      effective GB/s of the whole job = 4 transpositions x 2 x 512^3 x 16 B / cycle time,
             inputs resident in HBM, CUDA events on the launching stream, max over ranks.
  e2e        same metric through the public call with HOST (pinned) buffers: the H2D copy of the
             input and the D2H copy of the result are inside the timed region.
  roofline   dominant kernel (transpose_tiles_kernel): algorithmic bytes per launch
             (2 x elements x 16 B) / its mean launch time (CUDA events around each launch).
  cpu_baseline  oracle port of the reference host kernels (C/OpenMP) timed on the box's cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GLOBAL = 512
ES = 16  # complex128
METRIC = "512^3 C2C fp64 fwd+bwd transpose cycle, effective GB/s"
UNIT = "GB/s"
CYCLE_BYTES = 4 * 2 * N_GLOBAL ** 3 * ES  # 17.18 GB: 4 transpositions, one read + one write each
WORKLOAD = ("3D C2C fp64 512^3 transpose-only pencil cycle X->Y->Z->Y->X on {world} B200 "
            "(BASELINE configs[1]) through dtfft_execute FORWARD + BACKWARD")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML, ~2 ms period)."""

    def __init__(self, index=0):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self.index = index

    def _run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                    "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while not self._stop.is_set():
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for nm, b in bits.items():
                    if r & b:
                        self.reasons.add(nm)
                self._stop.wait(0.002)
        except Exception as ex:  # NVML missing: fall back to one nvidia-smi query
            self.error = str(ex)
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
            except Exception:
                pass

    def __enter__(self):
        self._t.start()
        time.sleep(0.05)
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def oracle_lib():
    path = os.path.join(ROOT, "oracle", "_build", "liboracle_host.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "_build/liboracle_host.so"], check=True,
                       stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(path)
    lib.oracle_kernel_execute.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.c_int,
                                          ctypes.c_int]
    lib.oracle_kernel_execute_blocked.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return lib


CPU_VARIANTS = [0, 4, 8, 16, 32, 64]  # 0 = unblocked loops, else BLOCK_SIZE of the blocked permutes


def cpu_cycle_times(n, reps, block=0, bufs=None):
    """Time the oracle port (reference host kernels restated in C/OpenMP) on an n^3 c128 cycle.
    block = 0: the unblocked loops (_dtfft_kernel_host_routines.inc); 4..64: the blocked permutes
    (_dtfft_kernel_host_block_routines.inc)."""
    import numpy as np

    lib = oracle_lib()
    lib.oracle_set_num_threads(len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1
    dims = (ctypes.c_int32 * 3)(n, n, n)
    if bufs is None:
        a = np.random.default_rng(1234).random(2 * n ** 3)  # n^3 complex128 as float64 pairs
        b = np.empty_like(a)
    else:
        a, b = bufs
    pa, pb = a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p)
    FWD, BWD = 7, 8
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for kt, (src, dst) in ((FWD, (pa, pb)), (FWD, (pb, pa)), (BWD, (pa, pb)), (BWD, (pb, pa))):
            if block:
                rc = lib.oracle_kernel_execute_blocked(kt, 3, dims, ES, src, dst, block)
            else:
                rc = lib.oracle_kernel_execute(kt, 3, dims, ES, src, dst, None, 0, 0)
            assert rc == 0
        times.append(time.perf_counter() - t0)
    return times, lib.oracle_num_threads()


def cpu_pick_variant(n_sample=256):
    """The reference's host kernel times its unblocked loops and its blocked variants
    (BLOCK_SIZE 4..64) and keeps the fastest; do the same on an n_sample^3 cycle."""
    import numpy as np

    a = np.random.default_rng(1234).random(2 * n_sample ** 3)
    b = np.empty_like(a)
    cpu_cycle_times(n_sample, 1, 0, (a, b))  # page faults
    table = {}
    for blk in CPU_VARIANTS:
        t, _ = cpu_cycle_times(n_sample, 2, blk, (a, b))
        table[blk] = min(t)
    best = min(table, key=table.get)
    return best, {("unblocked" if k == 0 else f"block_{k}"): v * 1e3 for k, v in table.items()}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself
    (Fortran 2018 + MPI) cannot be built in this image, so this is the oracle PORT of its host
    kernels (src/include/_dtfft_kernel_host_routines.inc and the blocked variants of
    _dtfft_kernel_host_block_routines.inc, the fastest of them like the reference's own host-kernel
    autotune) with every host thread OpenMP gives.  --steps / --warmup are honoured: a step is one
    full 512^3 cycle (about 0.4 s)."""
    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = N_GLOBAL
    block, table = cpu_pick_variant(256)
    a = np.random.default_rng(1234).random(2 * n ** 3)
    b = np.empty_like(a)
    cpu_cycle_times(n, max(1, args.warmup), block, (a, b))
    times, threads = cpu_cycle_times(n, max(1, args.steps), block, (a, b))
    t = sum(times) / len(times)
    val = CYCLE_BYTES / t / 1e9
    variant = "unblocked" if block == 0 else f"block_{block}"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": max(1, args.steps),
        "warmup": max(1, args.warmup), "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128 (opaque 16-byte moves)", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(world=args.gpus),
                   "where": "single process, host memory, every host thread",
                   "host_kernel_variant": variant, "variant_ms_on_256^3": table,
                   "note": "reference CPU path = oracle port of dtFFT host kernels (reference needs Fortran+MPI: unbuildable here)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{max(1, args.steps)} full 512^3 cycle(s) after {max(1, args.warmup)} warm-up cycle(s), "
                                   f"host-kernel variant {variant} (fastest of unblocked / block 4..64 on a 256^3 sample)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


DPERM = [[0, 1, 2], [1, 2, 0], [2, 0, 1]]  # local axis j of pencil d is natural axis DPERM[d][j] (transpose_plan.F90:1046-1082)


def expected_pencil(torch, pencil, d, dims, device):
    """What rank-local pencil `d` (0 = X, 1 = Y, 2 = Z) must hold when element (x, y, z) of the global
    array carries G = x + Nx (y + Ny z): the result of the reference's host MPI-datatype path
    (src/dtfft_reshape_handle_datatype.F90:479-574 is a redistribution of G, nothing else).  Returned in
    the pencil's memory order (local axis 0 fastest), float64."""
    nat_s, nat_c = [0, 0, 0], [0, 0, 0]
    for j in range(3):
        nat_s[DPERM[d][j]], nat_c[DPERM[d][j]] = pencil.starts[j], pencil.counts[j]
    ax = [torch.arange(nat_s[k], nat_s[k] + nat_c[k], device=device, dtype=torch.float64) for k in range(3)]
    want = ax[0][None, None, :] + dims[0] * (ax[1][None, :, None] + dims[1] * ax[2][:, None, None])  # [z][y][x]
    return want.permute(2 - DPERM[d][2], 2 - DPERM[d][1], 2 - DPERM[d][0]).reshape(-1)


def encode(torch, buf, want):
    """16-byte element = (G, -G - 1) as two float64: both halves identify the element."""
    v = buf[: 2 * want.numel()].view(-1, 2)
    v[:, 0] = want
    v[:, 1] = -want - 1.0


def holds(torch, buf, want):
    v = buf[: 2 * want.numel()].view(-1, 2)
    return bool(torch.equal(v[:, 0], want)) and bool(torch.equal(v[:, 1], -want - 1.0))


def _plan_cycle(plan, Execute, a, b, aux):
    """One fwd+bwd transpose-only cycle through the public API: X->Y->Z then Z->Y->X (2 calls)."""
    plan.execute(a, b, Execute.FORWARD, aux)
    plan.execute(b, a, Execute.BACKWARD, aux)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from dtfft_b200.comm import TorchComm
    from dtfft_b200.plan import Backend, Config, Execute, PlanC2C, Transpose

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = TorchComm()

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    n = N_GLOBAL
    dims = [n, n, n]
    names = {"nccl": Backend.NCCL, "nccl_pipe": Backend.NCCL_PIPELINED, "nvlink": Backend.NVLINK_FUSED}
    if world == 1:
        cand = [("single", Backend.NONE)]
    elif args.backend == "auto":
        cand = list(names.items())
    else:
        cand = [(args.backend, names[args.backend])]

    results = {}
    best = None
    for bname, backend in cand:
        # the plan enqueues on a stream torch owns (dtfft_config_t.stream), so torch's allocators
        # may record events on it for as long as they like
        stream = torch.cuda.Stream()
        cfg = Config(enable_z_slab=False, backend=backend, stream=stream)
        plan = PlanC2C(dims, comm=comm, config=cfg)
        assert plan.stream == stream.cuda_stream
        nbytes = plan.alloc_bytes
        bufs = [plan.mem_alloc(nbytes) for _ in range(2)]
        aux_buf = plan.mem_alloc(plan.aux_bytes)
        a, b = (torch.as_tensor(x, device="cuda").view(torch.float64) for x in bufs)
        aux = torch.as_tensor(aux_buf, device="cuda")
        n_local = nbytes // ES
        # ---- parity BEFORE anything is timed: index-encoded fill, every transposition and both execute
        # directions checked element for element, on device, against the analytic pencils ----------
        from dtfft_b200.plan import Layout

        pen = [plan.get_pencil(l) for l in (Layout.X_PENCILS, Layout.Y_PENCILS, Layout.Z_PENCILS)]
        parity = {}
        with torch.cuda.stream(stream):
            want = [expected_pencil(torch, pen[d], d, dims, a.device) for d in range(3)]

            def step(name, fn, src, dst, d_dst):
                dst.fill_(-7.0)  # a stale result of an earlier step must not pass ...
                aux.fill_(0x5B)  # ... nor a stale intermediate pencil / staging block of an earlier step
                stream.synchronize()
                barrier()
                fn()
                stream.synchronize()
                barrier()
                parity[name] = holds(torch, dst, want[d_dst])

            encode(torch, a, want[0])
            step("X_TO_Y", lambda: plan.transpose(a, b, Transpose.X_TO_Y, aux), a, b, 1)
            step("Y_TO_Z", lambda: plan.transpose(b, a, Transpose.Y_TO_Z, aux), b, a, 2)
            step("Z_TO_Y", lambda: plan.transpose(a, b, Transpose.Z_TO_Y, aux), a, b, 1)
            step("Y_TO_X", lambda: plan.transpose(b, a, Transpose.Y_TO_X, aux), b, a, 0)
            # the timed path: dtfft_execute (first call eager, second captured into a CUDA graph, third
            # replayed where graphs apply) -- all three must deliver the Z pencils / return the X pencils
            for rep in ("eager", "captured", "replayed"):
                encode(torch, a, want[0])
                step(f"execute_forward_{rep}", lambda: plan.execute(a, b, Execute.FORWARD, aux), a, b, 2)
                step(f"execute_backward_{rep}", lambda: plan.execute(b, a, Execute.BACKWARD, aux), b, a, 0)
            encode(torch, a, want[0])
        stream.synchronize()
        torch.cuda.synchronize()
        barrier()
        for _ in range(args.warmup):
            _plan_cycle(plan, Execute, a, b, aux)
        stream.synchronize()
        with torch.cuda.stream(stream):
            parity["after_warmup_cycles"] = holds(torch, a, want[0])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = 0
        with ClockSampler(local_rank) as clocks:
            barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(args.steps):
                plan.execute(a, b, Execute.FORWARD, aux)
                launches += plan.stats()["kernel_launches"]
                plan.execute(b, a, Execute.BACKWARD, aux)
                launches += plan.stats()["kernel_launches"]
            e1.record(stream)
            stream.synchronize()
            barrier()
        ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        with torch.cuda.stream(stream):
            parity["after_timed_cycles"] = holds(torch, a, want[0])
        stream.synchronize()
        del want
        # every rank must agree on every check
        flags = torch.tensor([1.0 if parity[k] else 0.0 for k in sorted(parity)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        parity = {k: ("bit-exact" if float(flags[i]) == 1.0 else "MISMATCH") for i, k in enumerate(sorted(parity))}
        peer_err = plan.peer_error() if hasattr(plan, "peer_error") else 0
        results[bname] = {"plan": plan, "ms": ms, "launches": launches, "clocks": clocks.summary(), "a": a, "b": b,
                          "aux": aux, "stream": stream, "bufs": bufs + [aux_buf], "n_local": n_local,
                          "backend": plan.backend.name, "parity": parity, "peer_error": peer_err}
        if best is None or ms < results[best]["ms"]:
            best = bname
    bad = {k: v["parity"] for k, v in results.items() if any(x != "bit-exact" for x in v["parity"].values())}
    if bad:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "error": "parity mismatch against the analytic pencils", "parity": bad}), flush=True)
        sys.exit(1)
    R = results[best]
    plan, a, b, aux, stream, n_local = R["plan"], R["a"], R["b"], R["aux"], R["stream"], R["n_local"]
    ms_per_step = R["ms"]
    value = CYCLE_BYTES / (ms_per_step * 1e-3) / 1e9

    # ---- per-transposition timing (events around each dtfft_transpose) -----------------------
    order = [(Transpose.X_TO_Y, a, b), (Transpose.Y_TO_Z, b, a), (Transpose.Z_TO_Y, a, b), (Transpose.Y_TO_X, b, a)]
    evs = []
    reps = min(args.steps, 10)
    barrier()
    for _ in range(reps):
        for t, x, y in order:
            s_, t_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record(stream)
            plan.transpose(x, y, t, aux)
            t_.record(stream)
            evs.append((s_, t_, plan.stats()))
    stream.synchronize()
    kms = [s_.elapsed_time(t_) for s_, t_, _ in evs]
    k_avg_ms = max_over_ranks(sum(kms) / len(kms))
    per_type = {}
    for i, (t, _, _) in enumerate(order):
        per_type[t.name] = max_over_ranks(sum(kms[i::4]) / len(kms[i::4]))
    k_bytes = 2 * n_local * ES  # one read + one write of the local pencil
    # transpositions without an exchange are one HBM-bound kernel; those with one are NVLink-bound
    remote_of = {t.name: max_over_ranks(evs[i][2]["remote_bytes"]) for i, (t, _, _) in enumerate(order)}
    local_names = [nm for nm, rb in remote_of.items() if rb == 0]
    exch_names = [nm for nm, rb in remote_of.items() if rb > 0]
    loc_ms = sum(per_type[nm] for nm in local_names) / len(local_names) if local_names else None
    achieved = k_bytes / (loc_ms * 1e-3) / 1e9 if loc_ms else k_bytes / (k_avg_ms * 1e-3) / 1e9
    link = {}
    for nm in exch_names:
        rb = remote_of[nm]
        link[nm] = {"remote_bytes_per_gpu": rb, "ms": per_type[nm],
                    "GBps_per_direction": rb / (per_type[nm] * 1e-3) / 1e9,
                    "frac_of_900": rb / (per_type[nm] * 1e-3) / 1e9 / 900.0}
    peak, peak_src = measured_peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu capture
    if world == 1 and os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("transpose_tiles_kernel_bytes_per_launch")
        except Exception:
            traffic = None
    # name the kernel instantiation from the kernel plugin itself (B200 tile table, kernel_object.cu)
    from dtfft_b200.kernel import KERNEL_PERMUTE_FORWARD, Kernel

    probe = Kernel().create([64, 64, 64], 0, ES, KERNEL_PERMUTE_FORWARD)
    ki = probe.info()
    probe.destroy()
    tname = f"transpose_tiles_kernel<uint4,{ki['tile_a'] // 32},{ki['tile_b'] // 32},{ki['threads'] // 32}>"
    kernel_name = f"{tname} (one launch per local transposition)"
    forms = {t.name: plan.exchange_form(t) for t, _, _ in order}
    if R["backend"] == "NVLINK_FUSED" and any(v["form"] == "copy engines" for v in forms.values()):  # forced everywhere
        ncop = max(v["copies_per_execute"] for v in forms.values())
        exchange_kernel_name = (f"{tname} packing each peer slice locally + {ncop} strided cudaMemcpy3DAsync copies to the "
                                "peers' final addresses (+ peer barriers / pairwise flags)")
    elif R["backend"] == "NVLINK_FUSED" and any(v["form"].startswith("direct-store kernel alone") for v in forms.values()):
        exchange_kernel_name = (f"timed alone: {tname} with peer-mapped destinations (+ peer barriers); inside dtfft_execute: "
                                f"{tname} packing each peer slice + strided cudaMemcpy3DAsync copies, pipelined peer by peer "
                                "with the local transposition next to it")
    elif R["backend"] == "NVLINK_FUSED":
        exchange_kernel_name = f"{tname} with peer-mapped destinations (+ peer barriers)"
    else:
        exchange_kernel_name = "transpose_tiles_kernel (pack) + ncclSend/Recv + rows_copy_kernel (unpack)"

    # ---- end to end with host buffers (H2D + cycle + D2H inside the timed region) -----------
    # Every step copies ITS input pencil from pinned host memory to the device, runs the cycle
    # through the plan API and copies the result back.  `serial`: one step after the other on the
    # plan stream.  `pipelined` (the reported e2e): two buffer sets and three streams, so the D2H
    # of step k overlaps the H2D of step k+1 (PCIe is full duplex) -- what a streaming caller does.
    h_in = [torch.empty(2 * n_local, dtype=torch.float64, pin_memory=True) for _ in range(2)]
    for h in h_in:
        h.copy_(a[: 2 * n_local])
    h_out = [torch.empty(2 * n_local, dtype=torch.float64, pin_memory=True) for _ in range(2)]
    e2e_steps = max(2, min(args.steps, 6))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def e2e_step():
        with torch.cuda.stream(stream):
            a[: 2 * n_local].copy_(h_in[0], non_blocking=True)
            _plan_cycle(plan, Execute, a, b, aux)
            h_out[0].copy_(a[: 2 * n_local], non_blocking=True)

    e2e_step()
    stream.synchronize()
    barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    e1.record(stream)
    stream.synchronize()
    barrier()
    e2e_serial_ms = max_over_ranks(e0.elapsed_time(e1) / e2e_steps)
    assert torch.equal(h_out[0], h_in[0]), "e2e cycle is not the identity"

    set_bufs = [(a, b)]
    extra = [plan.mem_alloc(nbytes) for _ in range(2)]
    R["bufs"] += extra
    set_bufs.append(tuple(torch.as_tensor(x, device="cuda").view(torch.float64) for x in extra))
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_exec = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def e2e_pipelined(nsteps):
        for i in range(nsteps):
            k = i % 2
            xa, xb = set_bufs[k]
            if i >= 2:
                s_in.wait_event(ev_out[k])  # the result of step i-2 has left this buffer set
            with torch.cuda.stream(s_in):
                xa[: 2 * n_local].copy_(h_in[k], non_blocking=True)
                ev_in[k].record(s_in)
            stream.wait_event(ev_in[k])
            _plan_cycle(plan, Execute, xa, xb, aux)
            ev_exec[k].record(stream)
            s_out.wait_event(ev_exec[k])
            with torch.cuda.stream(s_out):
                h_out[k].copy_(xa[: 2 * n_local], non_blocking=True)
                ev_out[k].record(s_out)

    e2e_pipelined(2)
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    e0.record(s_in)
    e2e_pipelined(e2e_steps)
    s_out.wait_stream(stream)
    e1.record(s_out)
    torch.cuda.synchronize()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1) / e2e_steps)
    assert torch.equal(h_out[0], h_in[0]) and torch.equal(h_out[1], h_in[1]), "pipelined e2e cycle is not the identity"
    e2e_val = CYCLE_BYTES / (e2e_ms * 1e-3) / 1e9

    # ---- CPU baseline: oracle port on a bounded sample (rank 0, N = 1 only) -------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            block, table = cpu_pick_variant(256)
            variant = "unblocked" if block == 0 else f"block_{block}"
            cpu_cycle_times(n, 1, block)  # warm-up at full size (page faults)
            t512, threads = cpu_cycle_times(n, 2, block)
            tt = min(t512)
            cpu = {"value": CYCLE_BYTES / tt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"2 full 512^3 c128 cycles (4 permutes each) after 1 warm-up cycle, best of 2; oracle C/OpenMP port of the "
                             f"reference host kernels, variant {variant} = fastest of unblocked / block 4..64 on a 256^3 sample "
                             "(the reference's host kernel autotunes the same set)",
                   "ms_per_step": tt * 1e3, "variant": variant, "variant_ms_on_256^3": table}
        except Exception as ex:  # the baseline is reported, never a gate
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    grid = plan.grid_dims
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kernel_name, "bytes_per_launch": k_bytes,
                "avg_launch_ms": loc_ms if loc_ms else k_avg_ms, "peak_source": peak_src,
                "frac_of_nominal_8TBs": achieved / 8000.0}
    if world > 1:
        # N > 1: the cycle has two kinds of transposition with two different roofs; the top-level fields
        # describe the local (HBM-bound) kernel, `exchange` the NVLink-bound one -- never a blend
        roofline["scope"] = ("local transpositions only (" + ", ".join(local_names) + ")") if local_names else \
            "every transposition exchanges: HBM figure = whole transposition incl. its exchange"
        roofline["local"] = {"bound": "hbm", "kernel": f"{tname} (one launch, no exchange)", "transpositions": local_names,
                             "bytes_per_launch": k_bytes, "avg_launch_ms": loc_ms, "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak} if local_names else None
        ex_bytes = sum(v["remote_bytes_per_gpu"] for v in link.values())
        ex_ms = sum(v["ms"] for v in link.values())
        ach = ex_bytes / (ex_ms * 1e-3) / 1e9 if ex_ms > 0 else 0.0
        roofline["exchange"] = {"bound": "nvlink", "kernel": exchange_kernel_name, "transpositions": exch_names,
                                "bytes_out_per_launch_per_gpu": ex_bytes / max(1, len(link)),
                                "avg_launch_ms": ex_ms / max(1, len(link)), "achieved": ach, "peak": 900.0,
                                "unit": "GB/s per direction per GPU", "frac": ach / 900.0,
                                "frac_of_measured_dma_737": ach / 737.0,
                                "note": "remote payload / whole transposition time (device barriers + local tiles + remote "
                                        "stores); 737 GB/s = cudaMemcpyPeerAsync on this pool (profiles/r01c_p2p_n8.json)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128 (opaque 16-byte moves)", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(world=world),
                   "global_dims": dims, "grid": grid, "element_bytes": ES, "bytes_per_step": CYCLE_BYTES,
                   "backend": R["backend"], "transposition_ms": per_type, "exchange_form": forms,
                   "switches": {k: os.environ[k] for k in ("DTFFTB_PAIR_OVERLAP", "DTFFTB_FUSED_MODE", "DTFFTB_DMA_SUB_BYTES",
                                                            "DTFFTB_GRAPHS", "DTFFTB_TILE") if k in os.environ},
                   "backends_ms_per_step": {k: v["ms"] for k, v in results.items()},
                   "l2": f"working set 2 x {n_local * ES / 2**20:.0f} MiB per transposition per GPU >> 126 MB L2 (no flush needed)"
                   if n_local * ES > 2**28 else
                   f"working set {3 * n_local * ES / 2**20:.0f} MiB over 3 buffers per GPU cycles through L2 (126 MB); buffers alternate so no transposition re-reads what the previous one wrote from L2 alone"},
        "roofline": roofline,
        "parity": {"checked": "every element of every rank, on device, against the analytic X/Y/Z pencils of the "
                              "index-encoded global array (= the reference's host MPI-datatype path), before the timed region; "
                              "the round trip again after the warm-up and after the timed cycles",
                   "per_backend": {k: v["parity"] for k, v in results.items()},
                   "peer_error": {k: v["peer_error"] for k, v in results.items()}},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": n_local * ES * world,
                "d2h_bytes_per_step": n_local * ES * world, "steps": e2e_steps,
                "how": "pinned host -> device copy of every step's input, dtfft_execute FORWARD + BACKWARD, device -> "
                       "pinned host copy of the result; two buffer sets so the D2H of step k overlaps the H2D of step k+1",
                "serial_ms_per_step": e2e_serial_ms, "serial_value": CYCLE_BYTES / (e2e_serial_ms * 1e-3) / 1e9},
        "gpu_launches": R["launches"], "graph_replays": plan.graph_replays,
        "clocks": R["clocks"],
    }
    if world > 1:
        ex_bytes = sum(v["remote_bytes_per_gpu"] for v in link.values())
        line["nvlink"] = dict(roofline["exchange"], per_exchange_transposition=link,
                              bytes_out_per_cycle_per_gpu=ex_bytes)
    for v in results.values():
        for x in v["bufs"]:
            v["plan"].mem_free(x)
        v["plan"].destroy()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--backend", default="auto", choices=["auto", "nccl", "nccl_pipe", "nvlink"],
                    help="N>1: exchange backend; auto times all three and reports the fastest")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
