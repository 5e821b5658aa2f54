"""Min-sum benchmark decoder -- B200 back-end.

Drop-in for ``Continous_LDPC_Decoding/min_sum_decoder_irreg.py`` of the reference
(constructor :23, ``decode_OpenCL_min_sum`` :221-287, ``return_errors_all_zero`` :290-295).
Messages follow the dtype of the input buffer: float64 (the reference's dtype and the default of
the quantizer / of numpy inputs) reproduces the reference arithmetic -- min-sum bit for bit --
while float32 buffers (``quanti.llr_dtype = np.float32`` or ``self.precision = 'f32'`` for host
inputs) select the faster fp32 kernels, whose hard decisions differ from float64 only where an
a-posteriori LLR is a rounding-level tie (see tests/test_gpu_parity.py for the measured agreement).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..device_array import DeviceArray
from ..engine import GraphDecoderBase, count_errors, stream_ptr


class _LlrDecoderBase(GraphDecoderBase):
    _algo = _lib.ALGO_MINSUM

    def __init__(self, filename, imax_, cardinality_T_channel_, msg_at_time_):
        self._load_graph(filename)
        self.imax = imax_
        self.cardinality_T_channel = cardinality_T_channel_
        self._set_rate()
        self.msg_at_time = msg_at_time_
        self.early_termination = True     # the reference always checks the syndrome (:262-270)
        self.precision = 'f64'            # dtype used for host (numpy) inputs ('f32' = fast path)
        # 'flooding' = the reference's schedule (every check node, then every variable node: :242-273); 'layered' = the
        # opt-in row-message-passing schedule (csrc/llr_layered.cu): one a-posteriori LLR per variable node, updated
        # check by check in layers of checks that share no variable.  About half the passes for the same error rate;
        # results differ from the reference by construction.
        self.schedule = 'flooding'
        self.last_i_num = None

    def init_OpenCL_decoding(self, msg_at_time_, context_=False):
        """min_sum_decoder_irreg.py:167-218 -- uploads the graph tables; ``context_`` is ignored."""
        self.msg_at_time = msg_at_time_
        self.context = context_
        self._ensure_handle()

    init_decoding = init_OpenCL_decoding

    def _decode_llr(self, received_blocks, buffer_in, return_buffer, early_termination=None):
        h = self._ensure_handle()
        early = self.early_termination if early_termination is None else early_termination
        if buffer_in:
            t = received_blocks.tensor if isinstance(received_blocks, DeviceArray) else received_blocks
            tdt = t.dtype if t.dtype in (torch.float32, torch.float64) else torch.float64
            ch = self._device_input(received_blocks, tdt)
        else:
            rb = np.asarray(received_blocks, dtype=np.float64)
            if rb.ndim == 1:
                rb = rb[:, None]
            tdt = torch.float64 if self.precision == 'f64' else torch.float32
            ch = self._device_input(torch.from_numpy(np.ascontiguousarray(rb)).cuda(), tdt)
        out = torch.empty_like(ch)
        # asynchronous on the current stream; i_num is read back lazily (self.last_i_num)
        if self.schedule not in ('flooding', 'layered'):
            raise ValueError("schedule must be 'flooding' or 'layered'")
        fn = _lib.lib().ibldpc_decode_llr_layered if self.schedule == 'layered' else _lib.lib().ibldpc_decode_llr
        _lib.check(fn(
            h, self._algo, _lib.F32 if tdt == torch.float32 else _lib.F64, C.c_void_p(ch.data_ptr()), ch.shape[1],
            int(self.imax), int(bool(early)), C.c_void_p(out.data_ptr()), None, C.c_void_p(stream_ptr())))
        self._inum_pending = True
        if return_buffer:
            return DeviceArray(out)
        return out.cpu().numpy().astype(np.float64)

    def layer_count(self):
        """Number of layers of the layered schedule for this code (greedy colouring of the checks in index order)."""
        n = C.c_int32(0)
        _lib.check(_lib.lib().ibldpc_layer_count(self._ensure_handle(), C.byref(n)))
        return int(n.value)

    def return_errors_all_zero(self, varnode_output_buffer):
        """Decoded 1-bits (LLR < 0) in the first data_len rows (min_sum_decoder_irreg.py:290-295)."""
        if isinstance(varnode_output_buffer, np.ndarray):
            varnode_output_buffer = DeviceArray(torch.from_numpy(np.ascontiguousarray(varnode_output_buffer)).cuda())
        return count_errors(varnode_output_buffer, int(self.data_len))[0]

    def count_errors(self, varnode_output_buffer, ref_bits=None, rows=None):
        return count_errors(varnode_output_buffer, int(self.data_len) if rows is None else rows, None, ref_bits)

    def decode_on_host(self, channel_values_):
        """Single frame, host in / host out, no early termination; runs on the GPU (the reference's
        own method raises, min_sum_decoder_irreg.py:320-383)."""
        ch = np.asarray(channel_values_, dtype=np.float64).reshape(-1)
        return self._decode_llr(ch[:, None], False, False, early_termination=False)[:, 0]


class Min_Sum_Decoder_class_irregular(_LlrDecoderBase):
    _algo = _lib.ALGO_MINSUM

    def decode_OpenCL_min_sum(self, received_blocks, buffer_in=False, return_buffer=False, early_termination=None):
        return self._decode_llr(received_blocks, buffer_in, return_buffer, early_termination)

    decode = decode_OpenCL_min_sum

    def discrete_cn_operation(self, vec_y_c, iter_):
        """sign*min chain (min_sum_decoder_irreg.py:298-308); numpy helper, not a decode path."""
        y = np.asarray(vec_y_c, dtype=np.float64)
        t = y[:, 0]
        for l in range(y.shape[1] - 1):
            t = np.sign(y[:, l + 1] * t) * np.minimum(np.abs(t), np.abs(y[:, l + 1]))
        return t

    def discrete_vn_operation(self, vec_y_v, iter_):
        """Running sum (min_sum_decoder_irreg.py:310-318)."""
        y = np.asarray(vec_y_v, dtype=np.float64)
        t = y[:, 0]
        for l in range(y.shape[1] - 1):
            t = y[:, l + 1] + t
        return t
