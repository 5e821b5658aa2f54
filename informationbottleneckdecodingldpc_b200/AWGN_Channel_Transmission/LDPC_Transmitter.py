"""LDPC-encoded BPSK transmitter with the interface of the reference's ``LDPC_BPSK_Transmitter``
(AWGN_Channel_Transmission/LDPC_Transmitter.py:17-132): random information bits, systematic LDPC
encoding, BPSK mapping -- a whole batch per call on the GPU instead of a Python loop over frames
(:113-119).

``transmit()`` returns the BPSK symbols ``(codeword_len, msg_at_time)`` and keeps the information bits in
``last_transmitted_bits`` exactly like the reference.  With ``return_buffer_only = True`` both stay on the
device (DeviceArray), which is what the on-device BER loop uses; ``transmit_bits()`` returns the coded bits
themselves (uint8) for the fused BPSK + AWGN kernel (``AWGN_channel.transmission_bits``).

``data_len`` is ``N - M`` (the encoder's K).  The reference derives it from the degree distribution in
floating point (:36-64), which for DVB-S2-like codes gives K - 1 and then indexes past the message
(``encode_c`` reads ``NumInfoBits`` entries); for every code whose float formula is exact the two agree.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse as sp
import torch

from .. import _lib
from ..device_array import DeviceArray
from ..Discrete_LDPC_decoding.LDPC_encoder import LDPCEncoder
from ..engine import current_device, stream_ptr
from ..graph import code_rate_from_degrees, load_check_matrix
from ..rng import PhiloxStream


class LDPC_BPSK_Transmitter(PhiloxStream):
    def __init__(self, filename_H_, msg_at_time=1):
        self.filename_H = filename_H_
        if isinstance(filename_H_, (str, os.PathLike)):
            self.H_sparse = load_check_matrix(str(filename_H_))
        else:
            self.H_sparse = sp.csr_matrix(filename_H_)
        self.set_code_parameters()
        try:
            self.encoder = LDPCEncoder(self.H_sparse)
            self.data_len = int(self.encoder.K)
        except ValueError as e:
            # The reference prints 'Not invertible Matrix' and carries on (LDPC_encoder.py:246-247); its all-zero
            # codeword drivers build a transmitter only to read R_c (Regular_LDPC_Decoding/BPSK/
            # BER_simulation_OpenCL.py:76,85).  Same here: the object stays usable for the code parameters and
            # transmit() reports why it cannot encode.
            print("Not invertible Matrix")
            self.encoder = None
            self._encoder_error = str(e)
            self.data_len = int(self.N_v - self.N_c)
        self.last_transmitted_bits = []
        self.msg_at_time = int(msg_at_time)
        self.return_buffer_only = False
        self.seed = 20181001
        self._offset = 0          # Philox counter offset: consecutive calls draw disjoint ranges of the stream
        self._stream = None       # sub-stream: set_stream(rank); default = rank of the process group, else 0

    def set_code_parameters(self):
        H = self.H_sparse
        self.degree_checknode_nr = np.asarray(H.sum(1)).astype(np.int64)[:, 0]
        self.degree_varnode_nr = np.asarray(H.sum(0)).astype(np.int64)[0, :]
        self.N_v, self.N_c = H.shape[1], H.shape[0]
        self.d_c_max, self.d_v_max = int(self.degree_checknode_nr.max()), int(self.degree_varnode_nr.max())
        self.codeword_len = H.shape[1]
        self.R_c = code_rate_from_degrees(H)

    # ---- device pieces ------------------------------------------------------------------------
    def random_bits(self):
        """(data_len, msg_at_time) uint8 information bits on the device (replaces np.random.randint, :111)."""
        dev = current_device()
        n = self.data_len * self.msg_at_time
        t = torch.empty((self.data_len, self.msg_at_time), dtype=torch.uint8, device=f"cuda:{dev}")
        _lib.check(_lib.lib().ibldpc_random_bits(dev, int(self._philox_key()), int(self._offset), n, C.c_void_p(t.data_ptr()),
                                                 C.c_void_p(stream_ptr())))
        self._offset += n
        return t

    def transmit_bits(self, uncoded_msgs=None):
        """Coded bits (codeword_len, msg_at_time) uint8 on the device; ``last_transmitted_bits`` is updated."""
        if self.encoder is None:
            raise ValueError(f"this parity-check matrix has no systematic encoder: {self._encoder_error}")
        bits = self.random_bits() if uncoded_msgs is None else uncoded_msgs
        coded = self.encoder.encode_batch(bits)
        if self.return_buffer_only:
            self.last_transmitted_bits = DeviceArray(bits) if isinstance(bits, torch.Tensor) else bits
        else:
            self.last_transmitted_bits = (bits.cpu().numpy() if isinstance(bits, torch.Tensor) else np.asarray(bits)).astype(np.int64)
        return coded

    def transmit(self):
        coded = self.transmit_bits()
        data = self.BPSK_mapping(coded)
        return data if self.return_buffer_only else data.get()

    def BPSK_mapping(self, X):
        """0 -> +1, 1 -> -1 (:127-132); numpy in -> numpy out, device in -> device out."""
        if isinstance(X, DeviceArray):
            return DeviceArray(1.0 - 2.0 * X.tensor.to(torch.float64))
        X = np.asarray(X)
        data = np.ones(X.shape)
        data[X == 1] = -1
        return data
