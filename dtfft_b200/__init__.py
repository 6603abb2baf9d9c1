"""dtfft_b200 -- B200-native (sm_100a) implementation of dtFFT's GPU reshape path.

Host-side mirror of the reference's plugin interfaces over ``libdtfft_b200.so``:
``kernel.Kernel`` (abstract_kernel / kernel_device).  The CUDA extension is mandatory:
there is no CPU fallback on this path.
"""
from ._lib import DtfftB200Error, LIB_PATH, lib  # noqa: F401
from . import kernel  # noqa: F401
from . import comm, plan  # noqa: F401
from .plan import (AccessMode, Backend, Config, DtfftError, Effort, Execute, Executor, Layout, Pencil, PlanC2C,  # noqa: F401
                   PlanR2C, PlanR2R, Platform, Precision, R2RKind, Request, Reshape, Transpose, TransposeMode, Version,
                   get_backend_string, is_compression_enabled, is_cuda_enabled, is_cufft_enabled, is_fftw_enabled,
                   is_mkl_enabled, is_nccl_enabled, is_nvshmem_enabled, is_transpose_only_enabled, is_vkfft_enabled)

dtfft_Exception = DtfftError  # the reference module's name for it (src/interfaces/python/__init__.py:112)

__version__ = "0.1.0"
