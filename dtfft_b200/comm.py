"""Process group handed to the C ABI in place of an ``MPI_Comm``.

The reference takes an ``mpi4py`` communicator (src/interfaces/python/__init__.py:134-136)
and uses it for metadata only.  Here the same role is played by a ``dtfftb_comm_t``
(include/dtfft_b200.h): rank, size and ONE collective -- allgather of N bytes per rank --
implemented over ``torch.distributed`` (``gloo`` on CPU boxes, ``nccl`` on GPUs).
PyTorch is plumbing only.
"""
from __future__ import annotations

import ctypes as C

_ALLGATHER_T = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)


class CommStruct(C.Structure):
    """``dtfftb_comm_t``."""

    _fields_ = [("rank", C.c_int32), ("size", C.c_int32), ("ctx", C.c_void_p), ("allgather", _ALLGATHER_T),
                ("cart_ndims", C.c_int32), ("cart_dims", C.c_int32 * 3)]


class TorchComm:
    """``dtfftb_comm_t`` over a ``torch.distributed`` process group.

    ``cart_dims`` (optional, dims[0] fastest like everywhere in dtFFT) plays the role of a
    user ``MPI_Cart_create`` communicator (src/dtfft_transpose_plan.F90:128-170)."""

    def __init__(self, group=None, cart_dims=None):
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised; pass comm=None for a single-rank plan")
        self._torch, self._dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        backend = dist.get_backend(group)
        self._device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        self._cb = _ALLGATHER_T(self._allgather)  # keep the thunk alive as long as the comm
        self.struct = CommStruct(self.rank, self.size, None, self._cb, 0, (C.c_int32 * 3)(1, 1, 1))
        if cart_dims is not None:
            self.struct.cart_ndims = len(cart_dims)
            for i, d in enumerate(cart_dims):
                self.struct.cart_dims[i] = int(d)

    def _allgather(self, ctx, send, recv, nbytes):
        try:
            torch = self._torch
            n = int(nbytes)
            src = torch.frombuffer((C.c_ubyte * n).from_address(send), dtype=torch.uint8).clone().to(self._device)
            dst = torch.empty(n * self.size, dtype=torch.uint8, device=self._device)
            self._dist.all_gather_into_tensor(dst, src, group=self.group)
            host = dst.cpu().contiguous()
            C.memmove(recv, host.data_ptr(), n * self.size)
            return 0
        except Exception as ex:  # never unwind through the C frame
            import sys

            print(f"[dtfft_b200] allgather callback failed: {ex!r}", file=sys.stderr, flush=True)
            return 1

    def pointer(self):
        return C.byref(self.struct)


def as_comm_pointer(comm):
    """None -> NULL (single rank); TorchComm -> pointer; 'world' -> default process group."""
    if comm is None:
        return None, None
    if isinstance(comm, str) and comm == "world":
        comm = TorchComm()
    if isinstance(comm, TorchComm) or (hasattr(comm, "pointer") and hasattr(comm, "struct")):
        return comm.pointer(), comm
    raise TypeError(f"unsupported communicator {type(comm)}: pass None, 'world' or a dtfft_b200.comm.TorchComm")
