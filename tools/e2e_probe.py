#!/usr/bin/env python
"""Where does the end-to-end (host buffers) figure of bench.py go?  One experiment per knob, on the link
between pinned host memory and each GPU (run alone or under torchrun with one rank per GPU):

  * bare cudaMemcpyAsync ceilings: H2D alone, D2H alone, both at once (what bench.py's pipelined e2e needs);
  * pinned memory allocated on the GPU's own NUMA node (process affinity set to that node's cores BEFORE the
    allocation, first touch there) against the default placement;
  * write-combined pinned memory as the H2D source;
  * chunked copies (how early the first transposition could start if H2D were split).

    python tools/e2e_probe.py [--mb 2048]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_probe.py

Prints one JSON line per experiment (rank 0 prints the max time over ranks = aggregate bandwidth)."""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_numa_node(index):
    import torch

    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        return int(open(path).read().strip()), path
    except Exception as ex:
        return -1, repr(ex)


def node_cpus(node):
    try:
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    except Exception:
        return None
    cpus = []
    for part in txt.split(","):
        if "-" in part:
            lo, hi = part.split("-")
            cpus += list(range(int(lo), int(hi) + 1))
        elif part:
            cpus.append(int(part))
    return cpus


def main():
    import torch
    import torch.distributed as dist

    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=4)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nbytes = (args.mb << 20) // world  # the bench's 2 GiB pencil is split over the ranks
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    rt.cudaFreeHost.argtypes = [ctypes.c_void_p]

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def barrier():
        if world > 1:
            dist.barrier()

    nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()) \
        if os.path.isdir("/sys/devices/system/node") else []
    gnode, gpath = gpu_numa_node(local_rank)
    emit({"what": "topology", "numa_nodes": nodes, "gpu_numa_node_rank0": gnode, "sysfs": gpath,
          "cpus_allowed": len(os.sched_getaffinity(0)), "n_ranks": world, "bytes_per_rank": nbytes})

    def host_alloc(flags=0):
        p = ctypes.c_void_p(0)
        rc = rt.cudaHostAlloc(ctypes.byref(p), nbytes, flags)
        assert rc == 0, rc
        ctypes.memset(p.value, 1, nbytes)  # first touch by this thread
        return p.value

    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def measure(label, h_src, h_dst, h2d, d2h, chunks=1, extra=None):
        def go():
            step = nbytes // chunks
            for c in range(chunks):
                off = c * step
                n = step if c < chunks - 1 else nbytes - off
                if h2d:
                    rt.cudaMemcpyAsync(d_in.data_ptr() + off, h_src + off, n, 1, ctypes.c_void_p(s1.cuda_stream))
                if d2h:
                    rt.cudaMemcpyAsync(h_dst + off, d_out.data_ptr() + off, n, 2, ctypes.c_void_p(s2.cuda_stream))

        go()
        torch.cuda.synchronize()
        barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize()
        e0.record(s1)
        s2.wait_event(e0)
        for _ in range(args.iters):
            go()
        e1.record(s1)
        s2.wait_event(e1)
        e2.record(s2)
        torch.cuda.synchronize()
        barrier()
        ms = allmax(max(e0.elapsed_time(e1), e0.elapsed_time(e2)) / args.iters)
        per_dir = nbytes * world / (ms * 1e-3) / 1e9
        rec = {"what": label, "h2d": h2d, "d2h": d2h, "chunks": chunks, "ms": ms, "aggregate_GBps_per_direction": per_dir,
               "per_gpu_GBps_per_direction": per_dir / world}
        if extra:
            rec.update(extra)
        emit(rec)

    # default placement
    hs, hd = host_alloc(), host_alloc()
    measure("default pinned", hs, hd, True, False)
    measure("default pinned", hs, hd, False, True)
    measure("default pinned", hs, hd, True, True)
    for ch in (4, 16):
        measure("default pinned, chunked", hs, hd, True, True, chunks=ch)
    rt.cudaFreeHost(hs)
    rt.cudaFreeHost(hd)
    # write-combined source
    hs, hd = host_alloc(0x04), host_alloc()
    measure("write-combined H2D source", hs, hd, True, False)
    measure("write-combined H2D source", hs, hd, True, True)
    rt.cudaFreeHost(hs)
    rt.cudaFreeHost(hd)
    # NUMA-local placement
    cpus = node_cpus(gnode) if gnode >= 0 else None
    if cpus:
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            hs, hd = host_alloc(), host_alloc()
            measure("NUMA-local pinned", hs, hd, True, False, extra={"node": gnode, "cpus": len(allowed)})
            measure("NUMA-local pinned", hs, hd, False, True, extra={"node": gnode, "cpus": len(allowed)})
            measure("NUMA-local pinned", hs, hd, True, True, extra={"node": gnode, "cpus": len(allowed)})
            rt.cudaFreeHost(hs)
            rt.cudaFreeHost(hd)
        else:
            emit({"what": "NUMA-local pinned", "skipped": "the GPU's node has no CPU this process may run on"})
    else:
        emit({"what": "NUMA-local pinned", "skipped": f"no NUMA information (gpu node {gnode}, nodes {nodes})"})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
