"""dtfft_b200 -- B200-native (sm_100a) implementation of dtFFT's GPU reshape path.

Host-side mirror of the reference's plugin interfaces over ``libdtfft_b200.so``:
``kernel.Kernel`` (abstract_kernel / kernel_device).  The CUDA extension is mandatory:
there is no CPU fallback on this path.
"""
from ._lib import DtfftB200Error, LIB_PATH, lib  # noqa: F401
from . import kernel  # noqa: F401
from . import comm, plan  # noqa: F401
from .plan import (Backend, Config, DtfftError, Effort, Execute, Executor, Layout, Pencil, PlanC2C, PlanR2C,  # noqa: F401
                   PlanR2R, Precision, R2RKind, Reshape, Transpose)

__version__ = "0.1.0"
