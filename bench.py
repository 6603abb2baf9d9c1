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
    return lib


def cpu_cycle_times(n, reps):
    """Time the oracle port (reference host kernels restated in C/OpenMP) on an n^3 c128 cycle."""
    import numpy as np

    lib = oracle_lib()
    dims = (ctypes.c_int32 * 3)(n, n, n)
    a = np.random.default_rng(1234).random(2 * n ** 3)  # n^3 complex128 as float64 pairs
    b = np.empty_like(a)
    pa, pb = a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p)
    FWD, BWD = 7, 8
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for kt, (src, dst) in ((FWD, (pa, pb)), (FWD, (pb, pa)), (BWD, (pa, pb)), (BWD, (pb, pa))):
            rc = lib.oracle_kernel_execute(kt, 3, dims, ES, src, dst, None, 0, 0)
            assert rc == 0
        times.append(time.perf_counter() - t0)
    return times, lib.oracle_num_threads()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself
    (Fortran 2018 + MPI) cannot be built in this image, so this is the oracle PORT of its host
    kernels (src/include/_dtfft_kernel_host_routines.inc) with every host thread OpenMP gives."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = N_GLOBAL
    steps = max(1, min(args.steps, 10))
    cpu_cycle_times(n, 1)  # warm-up (page faults)
    times, threads = cpu_cycle_times(n, steps)
    t = sum(times) / len(times)
    val = CYCLE_BYTES / t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128 (opaque 16-byte moves)", "data": "synthetic",
        "config": {"workload": "3D C2C fp64 512^3 transpose-only cycle X->Y->Z->Y->X, single process, host memory",
                   "note": "reference CPU path = oracle port of dtFFT host kernels (reference needs Fortran+MPI: unbuildable here)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{steps} full 512^3 cycle(s) after 1 warm-up"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from dtfft_b200.kernel import KERNEL_PERMUTE_BACKWARD, KERNEL_PERMUTE_FORWARD, Kernel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != 1:
        raise SystemExit("multi-GPU bench arrives with the plan layer (N=1 only in this revision)")

    n = N_GLOBAL
    dims = [n, n, n]
    N = n ** 3
    stream = torch.cuda.current_stream()
    a = torch.empty(2 * N, dtype=torch.float64, device="cuda")
    a.uniform_(0, 1)
    b = torch.empty_like(a)
    fwd = Kernel().create(dims, 0, ES, KERNEL_PERMUTE_FORWARD)
    bwd = Kernel().create(dims, 0, ES, KERNEL_PERMUTE_BACKWARD)
    info = fwd.info()

    def cycle(x, y):  # X->Y->Z->Y->X ; result back in x
        fwd.execute(x, y, stream)
        fwd.execute(y, x, stream)
        bwd.execute(x, y, stream)
        bwd.execute(y, x, stream)

    ref_sum = float(a.sum())
    for _ in range(args.warmup):
        cycle(a, b)
    torch.cuda.synchronize()
    assert float(a.sum()) == ref_sum, "cycle is not the identity"

    # ---- device-resident timing ------------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            cycle(a, b)
        e1.record(stream)
        torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    ms_per_step = total_ms / args.steps
    value = CYCLE_BYTES / (ms_per_step * 1e-3) / 1e9
    launches = 4 * args.steps

    # ---- per-launch timing of the dominant kernel (events around each launch) ---------------
    evs = []
    reps = min(args.steps, 10)
    for _ in range(reps):
        for k, (x, y) in ((fwd, (a, b)), (fwd, (b, a)), (bwd, (a, b)), (bwd, (b, a))):
            s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            k.execute(x, y, stream)
            t.record(stream)
            evs.append((s, t))
    torch.cuda.synchronize()
    kms = [s.elapsed_time(t) for s, t in evs]
    k_avg_ms = sum(kms) / len(kms)
    k_bytes = 2 * N * ES
    achieved = k_bytes / (k_avg_ms * 1e-3) / 1e9
    peak, peak_src = measured_peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu capture
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("transpose_tiles_kernel_bytes_per_launch")
        except Exception:
            traffic = None

    # ---- end to end with host buffers (H2D + cycle + D2H inside the timed region) -----------
    h_in = torch.empty(2 * N, dtype=torch.float64, pin_memory=True)
    h_in.copy_(a)
    h_out = torch.empty(2 * N, dtype=torch.float64, pin_memory=True)
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(1):
        a.copy_(h_in, non_blocking=True)
        cycle(a, b)
        h_out.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(e2e_steps):
        a.copy_(h_in, non_blocking=True)
        cycle(a, b)
        h_out.copy_(a, non_blocking=True)
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    assert torch.equal(h_out, h_in), "e2e cycle is not the identity"
    e2e_val = CYCLE_BYTES / (e2e_ms * 1e-3) / 1e9

    # ---- CPU baseline: oracle port on a bounded sample ------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu_cycle_times(256, 1)
            t256, threads = cpu_cycle_times(256, 3)
            t512, _ = cpu_cycle_times(n, 1)
            tt = min(t512)
            cpu = {"value": CYCLE_BYTES / tt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": "1 full 512^3 c128 cycle (4 permutes) after a 256^3 warm-up; oracle C/OpenMP port of the reference host kernels",
                   "ms_per_step": tt * 1e3, "ms_per_step_256": min(t256) * 1e3}
        except Exception as ex:  # the baseline is reported, never a gate
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "c128 (opaque 16-byte moves)", "data": "synthetic",
        "config": {"workload": "3D C2C fp64 512^3 transpose-only pencil cycle X->Y->Z->Y->X on 1 B200 (BASELINE configs[1])",
                   "global_dims": dims, "element_bytes": ES, "bytes_per_step": CYCLE_BYTES,
                   "l2": "working set 2 x 2 GiB per permute >> 126 MB L2 (no flush needed)",
                   "kernel": info},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "transpose_tiles_kernel<uint4,2,2,8>", "bytes_per_launch": k_bytes,
                     "avg_launch_ms": k_avg_ms, "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": N * ES,
                "d2h_bytes_per_step": N * ES, "steps": e2e_steps},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
