"""AWGN channel with the interface of the reference's ``AWGN_channel`` (AWGN_Channel_Transmission/
AWGN_channel.py:14-50), real noise, generated on the GPU (Philox4x32-10 + Box-Muller, ``ibldpc_awgn``).

``transmission(input)``: numpy in -> numpy out (computed on the device), DeviceArray in -> DeviceArray out.
``transmission_bits(bits)`` fuses the BPSK mapping of coded bits with the noise (one kernel, no float copy of
the symbols).  Complex noise (used by the reference's QAM transmitter only) is not on the decode path and is
not provided.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..device_array import DeviceArray, as_tensor
from ..engine import current_device, stream_ptr
from ..rng import PhiloxStream


class AWGN_channel(PhiloxStream):
    def __init__(self, sigma_n2_, complex=False):
        if complex:
            raise NotImplementedError("complex noise is only used by the reference's QAM transmitter")
        self.sigma_n2 = float(sigma_n2_)
        self.complex = False
        self.seed = 20181001 ^ 0x5DEECE66D
        self._offset = 0
        self._stream = None           # sub-stream: set_stream(rank); default = rank of the process group, else 0

    def _run(self, x_t, bits_t, shape):
        dev = current_device()
        n = int(np.prod(shape))
        out = torch.empty(tuple(shape), dtype=torch.float64, device=f"cuda:{dev}")
        _lib.check(_lib.lib().ibldpc_awgn(dev, C.c_void_p(x_t.data_ptr()) if x_t is not None else None,
                                          C.c_void_p(bits_t.data_ptr()) if bits_t is not None else None, n, self.sigma_n2,
                                          int(self._philox_key()), int(self._offset), C.c_void_p(out.data_ptr()),
                                          C.c_void_p(stream_ptr())))
        self._offset += n
        return out

    def transmission(self, input):
        if isinstance(input, (DeviceArray, torch.Tensor)):
            x = as_tensor(input).to(torch.float64).contiguous()
            return DeviceArray(self._run(x, None, x.shape))
        x = torch.from_numpy(np.ascontiguousarray(input, dtype=np.float64)).to(f"cuda:{current_device()}")
        return self._run(x, None, x.shape).cpu().numpy()

    def transmission_bits(self, bits):
        """y = (1 - 2 bit) + noise for coded bits (uint8 0/1) on the device -> DeviceArray float64."""
        b = as_tensor(bits).to(torch.uint8).contiguous()
        return DeviceArray(self._run(None, b, b.shape))
