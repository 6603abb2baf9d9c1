"""Helpers for the GPU parity tests: numpy <-> device buffers as opaque bytes."""
import numpy as np


def to_device(torch, arr: np.ndarray):
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).copy()).cuda()


def device_filled(torch, n_elems: int, dtype, fill_byte=0xAB):
    return torch.full((n_elems * np.dtype(dtype).itemsize,), fill_byte, dtype=torch.uint8, device="cuda")


def to_host(t, dtype) -> np.ndarray:
    return t.cpu().numpy().view(dtype)


def host_filled(n_elems: int, dtype, fill_byte=0xAB) -> np.ndarray:
    return np.full(n_elems * np.dtype(dtype).itemsize, fill_byte, dtype=np.uint8).view(dtype)
