"""Device byte buffers for the gpu-marked tests.  On a GPU these are exactly the torch calls the tests used to make;
when the tests run against the emulated library (B200_EMU=1, tests/emu/) "device" pointers are host pointers and a
256-byte aligned numpy array stands in.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import os

import numpy as np


def _emu() -> bool:
    return os.environ.get("B200_EMU") == "1"


class _HostDev:
    def __init__(self, nbytes: int):
        self._raw = np.zeros(nbytes + 256, dtype=np.uint8)
        off = (-self._raw.ctypes.data) % 256
        self.arr = self._raw[off:off + nbytes]

    def data_ptr(self) -> int:
        return self.arr.ctypes.data

    def __setitem__(self, key, value):
        self.arr[key] = value


def zeros(nbytes: int):
    """torch.zeros(nbytes, dtype=uint8, device='cuda') — or its host stand-in under the emulator."""
    if _emu():
        return _HostDev(nbytes)
    import torch
    return torch.zeros(nbytes, dtype=torch.uint8, device="cuda")


def to_dev(a: np.ndarray):
    """torch.from_numpy(a).cuda() — or the array itself under the emulator."""
    if _emu():
        return np.ascontiguousarray(a)
    import torch
    return torch.from_numpy(a).cuda()


def sync():
    if not _emu():
        import torch
        torch.cuda.synchronize()
