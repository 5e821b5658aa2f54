"""Thin device-array wrapper: the handful of ``pyopencl.array.Array`` operations the BER
drivers apply to decoder / quantizer buffers, on top of a torch CUDA tensor.

The reference's drivers and classes use exactly these on device buffers:
``.shape``, ``.data`` (discrete_LDPC_decoder.py:204-207), slicing ``[:data_len]``, ``__lt__``,
``.astype`` and ``.get()`` (discrete_LDPC_decoder_irreg.py:346-347).  PyTorch is used for device
memory only; all arithmetic of the hot path happens in libibldpc.so.
"""
from __future__ import annotations

import numpy as np
import torch

_NP2T = {np.dtype(np.uint8): torch.uint8, np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
         np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.bool_): torch.bool}


class DeviceArray:
    __slots__ = ("tensor",)

    def __init__(self, tensor: torch.Tensor):
        if not tensor.is_cuda:
            raise ValueError("DeviceArray wraps CUDA tensors only")
        self.tensor = tensor

    # --- pyopencl.array.Array look-alikes -------------------------------------------------
    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def dtype(self):
        return np.dtype(str(self.tensor.dtype).replace("torch.", ""))

    @property
    def data(self):
        return self

    @property
    def ptr(self) -> int:
        return self.tensor.data_ptr()

    def get(self) -> np.ndarray:
        return self.tensor.cpu().numpy()

    def astype(self, dtype):
        return DeviceArray(self.tensor.to(_NP2T[np.dtype(dtype)]))

    def __getitem__(self, idx):
        return DeviceArray(self.tensor[idx])

    def __lt__(self, other):
        return DeviceArray(self.tensor < other)

    def sum(self):
        return DeviceArray(self.tensor.sum().reshape(1))

    def __len__(self):
        return self.tensor.shape[0]

    @property
    def __cuda_array_interface__(self):
        return self.tensor.__cuda_array_interface__

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, device={self.tensor.device})"


def as_tensor(buf) -> torch.Tensor:
    """Accept a DeviceArray, a torch CUDA tensor or anything exposing ``.tensor``."""
    if isinstance(buf, DeviceArray):
        return buf.tensor
    if isinstance(buf, torch.Tensor):
        if not buf.is_cuda:
            raise ValueError("expected a CUDA tensor (buffer_in=True means a device buffer)")
        return buf
    raise TypeError(f"cannot interpret {type(buf).__name__} as a device buffer")


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """numpy view of page-locked host memory: host buffers handed to decode_OpenCL /
    ibldpc_decode_ib_host from such arrays are DMA'd without an intermediate staging copy."""
    t = torch.empty(tuple(shape), dtype=_NP2T[np.dtype(dtype)], pin_memory=True)
    return t.numpy()
