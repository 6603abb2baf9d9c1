"""In-process stand-in for a multi-rank world: P threads, one per simulated rank, each holding a
``dtfftb_comm_t`` whose allgather callback meets the others at a ``threading.Barrier``.
Lets the CPU suite drive the C++ plan logic (decomposition, exchange geometry) at 2..8 ranks
without spawning processes; the real ``torch.distributed`` (gloo) path is covered separately
in tests/test_plan_gloo.py."""
import ctypes as C
import threading

from dtfft_b200.comm import _ALLGATHER_T, CommStruct, TorchComm  # noqa: F401


class _ThreadComm:
    """Duck-types dtfft_b200.comm.TorchComm for as_comm_pointer()."""

    def __init__(self, world, rank, cart_dims=None):
        self.world, self.rank, self.size = world, rank, world.size
        self._cb = _ALLGATHER_T(self._allgather)
        self.struct = CommStruct(rank, world.size, None, self._cb, 0, (C.c_int32 * 3)(1, 1, 1))
        if cart_dims is not None:
            self.struct.cart_ndims = len(cart_dims)
            for i, d in enumerate(cart_dims):
                self.struct.cart_dims[i] = int(d)

    def _allgather(self, ctx, send, recv, nbytes):
        try:
            n = int(nbytes)
            w = self.world
            w.slots[self.rank] = C.string_at(send, n)
            w.barrier.wait(timeout=60)
            data = b"".join(w.slots)
            C.memmove(recv, data, n * w.size)
            w.barrier.wait(timeout=60)
            return 0
        except Exception as ex:  # pragma: no cover
            print("fake allgather failed:", repr(ex), flush=True)
            return 1

    def pointer(self):
        return C.byref(self.struct)


class ThreadWorld:
    def __init__(self, size):
        self.size = size
        self.barrier = threading.Barrier(size)
        self.slots = [b""] * size

    def run(self, fn, cart_dims=None):
        """Run ``fn(rank, comm)`` on every simulated rank; returns the list of results."""
        results, errors = [None] * self.size, [None] * self.size

        def body(r):
            try:
                results[r] = fn(r, _ThreadComm(self, r, cart_dims))
            except BaseException as ex:  # noqa: BLE001
                errors[r] = ex
                self.barrier.abort()

        ts = [threading.Thread(target=body, args=(r,)) for r in range(self.size)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        # a rank that fails validation aborts the barrier, which the others then see as a failed
        # allgather (DTFFTB_ERROR_COMM = -30002): report the root cause first
        for e in errors:
            if e is not None and not isinstance(e, threading.BrokenBarrierError) and getattr(e, "code", 0) != -30002:
                raise e
        for e in errors:
            if e is not None:
                raise e
        return results
