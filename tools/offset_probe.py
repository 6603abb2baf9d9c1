"""Does the relative placement of `in` and `out` matter for the 512^3 c128 permute?  (B200 probe)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtfft_b200.kernel import KERNEL_PERMUTE_BACKWARD, KERNEL_PERMUTE_FORWARD, Kernel  # noqa: E402
from dtfft_b200.plan import Config, PlanC2C, Transpose  # noqa: E402

n = 512
N = n ** 3
es = 16
GiB = 1 << 30
big = torch.empty(7 * GiB, dtype=torch.uint8, device="cuda")
base = big.data_ptr()
base = (base + 2 * 1024 * 1024 - 1) // (2 * 1024 * 1024) * (2 * 1024 * 1024)
k = Kernel().create([n, n, n], 0, es, KERNEL_PERMUTE_FORWARD)


def timeit(fn, stream=None, reps=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for delta in (2 * GiB, 2 * GiB + 256, 2 * GiB + 4096, 2 * GiB + 65536, 2 * GiB + (1 << 20), 2 * GiB + (2 << 20) + 4096,
              2 * GiB + (37 << 20), 3 * GiB, 4 * GiB):
    a, b = base, base + delta
    ms = timeit(lambda: k.execute(a, b))
    print(f"default stream, out = in + 2GiB + {delta - 2 * GiB:>10d} B: {ms:.4f} ms  {2 * N * es / ms / 1e6:.0f} GB/s", flush=True)

s = torch.cuda.Stream()
ms = timeit(lambda: k.execute(base, base + 2 * GiB + (37 << 20), s), s)
print(f"side stream: {ms:.4f} ms", flush=True)

x = torch.empty(2 * N, dtype=torch.float64, device="cuda")
y = torch.empty(2 * N, dtype=torch.float64, device="cuda")
print("torch tensors at", hex(x.data_ptr()), hex(y.data_ptr()), "delta", y.data_ptr() - x.data_ptr())
ms = timeit(lambda: k.execute(x, y))
print(f"torch tensors: {ms:.4f} ms", flush=True)

plan = PlanC2C([n, n, n], config=Config(enable_z_slab=False, stream=s))
pa, pb = plan.mem_alloc(plan.alloc_bytes), plan.mem_alloc(plan.alloc_bytes)
print("plan buffers at", hex(pa.ptr), hex(pb.ptr), "delta", pb.ptr - pa.ptr)
ms = timeit(lambda: k.execute(pa.ptr, pb.ptr))
print(f"plan buffers, kernel API: {ms:.4f} ms", flush=True)
ms = timeit(lambda: plan.transpose(pa, pb, Transpose.X_TO_Y), s)
print(f"plan buffers, plan.transpose: {ms:.4f} ms", flush=True)
ms = timeit(lambda: plan.transpose(x, y, Transpose.X_TO_Y), s)
print(f"torch tensors, plan.transpose: {ms:.4f} ms", flush=True)
