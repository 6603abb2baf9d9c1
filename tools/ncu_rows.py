"""Launches the family-R kernel (rows_copy_kernel) on the unpack of an 8-peer slab transposition at
512^3 complex128 -- the target of `ncu -k regex:rows_copy` (profiles/r01e_ncu_rows_copy.md)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtfft_b200.kernel import KERNEL_UNPACK, Kernel  # noqa: E402

n, P, es = 512, 8, 16
a = torch.empty(n ** 3 * es, dtype=torch.uint8, device="cuda").random_(0, 255)
b = torch.empty_like(a)
nxx = n // P
nd = np.zeros((P, 5), dtype=np.int32)
for i in range(P):
    nd[i] = (nxx, n, n, i * nxx * n * n, i * nxx)
k = Kernel().create([n, n, n], 0, es, KERNEL_UNPACK, nd)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    k.execute(a, b)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    k.execute(a, b)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print({"kernel": "rows_copy (unpack, 8 peers, one launch)", "ms": ms, "GBps": 2 * n ** 3 * es / ms / 1e6, "info": k.info()})
