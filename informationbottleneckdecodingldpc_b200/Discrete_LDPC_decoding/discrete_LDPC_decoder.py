"""Information-bottleneck (look-up table) decoder for regular LDPC codes -- B200 back-end.

Drop-in for ``Discrete_LDPC_decoding/discrete_LDPC_decoder.py`` of the reference: same
constructor, ``init_OpenCL_decoding`` / ``decode_OpenCL`` / ``return_errors_all_zero`` /
``decode_on_host`` / ``update_trellis_vectors`` and the attributes the BER drivers read.
The OpenCL context, Mako rendering and every CPU path are gone: all decoding runs in
libibldpc.so (hand-written sm_100a kernels) and raises if that library or a GPU is missing.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..device_array import DeviceArray
from ..engine import GraphDecoderBase, count_errors, current_device, stream_ptr
from ..luts import as_int32


class Discrete_LDPC_Decoder_class(GraphDecoderBase):
    """Reference signature: discrete_LDPC_decoder.py:30-31."""

    _irregular = False

    def __init__(self, filename, imax_, cardinality_T_channel_, cardinality_T_decoder_ops_,
                 Trellis_checknode_vector_a_, Trellis_varnode_vector_a_, msg_at_time_):
        self._load_graph(filename)
        if np.unique(self.degree_checknode_nr).size != 1 or np.unique(self.degree_varnode_nr).size != 1:
            raise ValueError("Discrete_LDPC_Decoder_class needs a regular code; use "
                             "Discrete_LDPC_Decoder_class_irregular for irregular parity-check matrices")
        self.imax = imax_
        self.cardinality_T_channel = cardinality_T_channel_
        self.cardinality_T_decoder_ops = cardinality_T_decoder_ops_
        self.matching_vector_checknode = None
        self.matching_vector_varnode = None
        self.match = 'false'
        self.update_trellis_vectors(Trellis_checknode_vector_a_, Trellis_varnode_vector_a_)
        self.msg_at_time = msg_at_time_
        self._post_init()

    def _post_init(self):
        # The reference always stops as soon as the whole batch has zero syndrome
        # (discrete_LDPC_decoder.py:233-276).  Set False for fixed-imax (throughput) runs, or 'frame' for the opt-in
        # per-frame stop with frame compaction: every frame gets exactly the result (and i_num, see
        # last_i_num_per_frame) the reference gives when that frame is decoded on its own (msg_at_time = 1).
        self.early_termination = True
        self.last_i_num_per_frame = None
        self.host_output_dtype = np.int32   # dtype decode_OpenCL(..., return_buffer=False) returns
        self.last_i_num = None
        self._luts_uploaded = False
        self._host_out = None
        self._host_bits = None

    # ---- tables ------------------------------------------------------------------------
    def update_trellis_vectors(self, Trellis_checknode_vector_a_, Trellis_varnode_vector_a_):
        """discrete_LDPC_decoder.py:53-55"""
        T = int(self.cardinality_T_decoder_ops)
        self.Trellis_checknode_vector_a = as_int32(Trellis_checknode_vector_a_, "Trellis_checknode_vector_a", T).astype(int)
        self.Trellis_varnode_vector_a = as_int32(Trellis_varnode_vector_a_, "Trellis_varnode_vector_a", T).astype(int)
        self._luts_uploaded = False

    def _upload_luts(self):
        h = self._ensure_handle()
        if self._luts_uploaded:
            return h
        T = int(self.cardinality_T_decoder_ops)
        cn = np.ascontiguousarray(self.Trellis_checknode_vector_a, dtype=np.int32)
        vn = np.ascontiguousarray(self.Trellis_varnode_vector_a, dtype=np.int32)
        use_match = self._irregular and str(self.match).lower() == 'true'
        mc = mv = None
        if use_match:
            if self.matching_vector_checknode is None or self.matching_vector_varnode is None:
                raise ValueError("match='true' needs matching_vector_checknode and matching_vector_varnode")
            mc = as_int32(self.matching_vector_checknode, "matching_vector_checknode", T)
            mv = as_int32(self.matching_vector_varnode, "matching_vector_varnode", T)
        d = _lib.LutDesc()
        d.card_channel = int(self.cardinality_T_channel)
        d.card_decoder = T
        d.imax = int(self.imax)
        d.cn_degree, d.vn_degree = int(self.d_c_max), int(self.d_v_max)
        d.cn_lut, d.cn_lut_len = cn.ctypes.data, cn.size
        d.vn_lut, d.vn_lut_len = vn.ctypes.data, vn.size
        if use_match:
            d.cn_match, d.cn_match_len = mc.ctypes.data, mc.size
            d.vn_match, d.vn_match_len = mv.ctypes.data, mv.size
        _lib.check(_lib.lib().ibldpc_set_luts(h, C.byref(d)))
        self._luts_uploaded = True
        return h

    # ---- reference entry points ------------------------------------------------------------
    def init_OpenCL_decoding(self, msg_at_time_, context_=False):
        """Upload graph tables and LUTs to the current GPU (discrete_LDPC_decoder.py:132-200).
        ``context_`` is accepted and ignored (there is no OpenCL context)."""
        self.msg_at_time = msg_at_time_
        self.context = context_
        self._upload_luts()

    init_decoding = init_OpenCL_decoding

    def decode_OpenCL(self, received_blocks, buffer_in=False, return_buffer=False, early_termination=None):
        """discrete_LDPC_decoder.py:202-295 / discrete_LDPC_decoder_irreg.py:245-341.

        buffer_in=True: ``received_blocks`` is a device buffer (DeviceArray / CUDA tensor) of
        cluster indices (N_v, B); otherwise a numpy array, copied inside the call.
        return_buffer=True returns a DeviceArray (uint8), else a numpy array of
        ``host_output_dtype``.  ``self.last_i_num`` holds the reference's ``i_num`` afterwards
        (device-buffer calls: only when early termination is on, to stay asynchronous otherwise)."""
        h = self._upload_luts()
        early = self.early_termination if early_termination is None else early_termination
        L = _lib.lib()
        inum = C.c_int32(0)
        per_frame = isinstance(early, str) and early.lower() == 'frame'
        if per_frame and not buffer_in:
            # host arrays: one upload, the device path below, one download (no chunking: the stop is per frame)
            rb = np.asarray(received_blocks)
            if rb.ndim == 1:
                rb = rb[:, None]
            if rb.dtype != np.uint8:
                if rb.size and (rb.min() < 0 or rb.max() >= int(self.cardinality_T_channel)):
                    raise ValueError("channel cluster indices must lie in [0, cardinality_T_channel)")
                rb = rb.astype(np.uint8)
            out = self.decode_OpenCL(DeviceArray(torch.from_numpy(np.ascontiguousarray(rb)).cuda()), buffer_in=True,
                                     return_buffer=True, early_termination='frame')
            if return_buffer:
                return out
            return out.get().astype(self.host_output_dtype, copy=False)
        if buffer_in:
            ch = self._device_input(received_blocks, torch.uint8)
            B = ch.shape[1]
            out = torch.empty_like(ch)
            if per_frame:
                inum_f = torch.empty(B, dtype=torch.int32, device=ch.device)
                _lib.check(L.ibldpc_decode_ib_perframe(h, C.c_void_p(ch.data_ptr()), B, int(self.imax), C.c_void_p(out.data_ptr()),
                                                       C.c_void_p(inum_f.data_ptr()), C.c_void_p(stream_ptr())))
                self.last_i_num_per_frame = DeviceArray(inum_f)
                self._inum_pending = True
                if return_buffer:
                    return DeviceArray(out)
                res = out.cpu().numpy().astype(self.host_output_dtype, copy=False)
                self.last_i_num
                return res
            # asynchronous on the current stream: i_num stays on the device until someone reads self.last_i_num
            _lib.check(L.ibldpc_decode_ib(h, C.c_void_p(ch.data_ptr()), B, int(self.imax), int(bool(early)),
                                          C.c_void_p(out.data_ptr()), None, C.c_void_p(stream_ptr())))
            self._inum_pending = True
            if return_buffer:
                return DeviceArray(out)
            res = out.cpu().numpy().astype(self.host_output_dtype, copy=False)
            self.last_i_num          # resolves i_num and raises on out-of-range cluster indices
            return res
        # host buffers
        rb = np.asarray(received_blocks)
        if rb.ndim == 1:
            rb = rb[:, None]
        if rb.shape[0] != self.N_v:
            raise ValueError(f"expected {self.N_v} rows (variable nodes), got {rb.shape[0]}")
        B = rb.shape[1]
        current_device()
        if rb.dtype != np.uint8:
            # the reference's own contract: integer numpy in, int32 numpy out (discrete_LDPC_decoder.py:207-209,
            # :292-295).  Narrowing / widening run on host threads inside the library, overlapped with the copies.
            if rb.dtype != np.int32:
                if rb.size and not np.issubdtype(rb.dtype, np.integer) and np.any(rb != np.floor(rb)):
                    raise ValueError("channel cluster indices must be integers")
                if rb.size and (rb.min() < 0 or rb.max() >= int(self.cardinality_T_channel)):
                    raise ValueError("channel cluster indices must lie in [0, cardinality_T_channel)")
                rb = rb.astype(np.int32)
            rb = np.ascontiguousarray(rb)
            out32 = np.empty(rb.shape, dtype=np.int32)
            try:
                _lib.check(L.ibldpc_decode_ib_host_i32(h, C.c_void_p(rb.ctypes.data), B, int(self.imax), int(bool(early)),
                                                       C.c_void_p(out32.ctypes.data), C.byref(inum)))
            except RuntimeError as e:
                if "cluster indices" in str(e):
                    raise ValueError(str(e)) from None
                raise
            self.last_i_num = int(inum.value)
            if return_buffer:
                return DeviceArray(torch.from_numpy(out32).to(torch.uint8).cuda())
            return out32 if np.dtype(self.host_output_dtype) == np.int32 else out32.astype(self.host_output_dtype)
        rb = np.ascontiguousarray(rb)
        if self._host_out is None or self._host_out.shape != rb.shape:
            from ..device_array import pinned_empty
            self._host_out = pinned_empty(rb.shape, np.uint8)
        out = self._host_out
        try:
            _lib.check(L.ibldpc_decode_ib_host(h, C.c_void_p(rb.ctypes.data), B, int(self.imax), int(bool(early)),
                                               C.c_void_p(out.ctypes.data), C.byref(inum)))
        except RuntimeError as e:
            if "cluster indices" in str(e):
                raise ValueError(str(e)) from None
            raise
        self.last_i_num = int(inum.value)
        if return_buffer:
            return DeviceArray(torch.from_numpy(out).cuda())
        if np.dtype(self.host_output_dtype) == np.uint8:
            return out          # pinned buffer owned by the decoder, overwritten by the next call
        return out.astype(self.host_output_dtype)

    decode = decode_OpenCL

    # ---- packed host buffers (opt-in): half the H2D bytes, 1/16 of the D2H bytes ------------------------------
    @staticmethod
    def pack_channel_values(received_blocks, out=None):
        """(N_v, B) cluster indices (< 16) -> (N_v, ceil(B/2)) uint8, frame f in nibble f & 1 of byte f >> 1."""
        rb = np.asarray(received_blocks)
        if rb.ndim == 1:
            rb = rb[:, None]
        rb = rb.astype(np.uint8, copy=False)
        if rb.shape[1] % 2:
            rb = np.concatenate([rb, np.zeros((rb.shape[0], 1), np.uint8)], axis=1)
        res = (rb[:, 0::2] | (rb[:, 1::2] << 4)).astype(np.uint8)
        if out is not None:
            out[...] = res
            return out
        return res

    @staticmethod
    def unpack_bits(bits, B):
        """(rows, ceil(B/8)) bit-packed hard decisions -> (rows, B) uint8 0/1."""
        return np.unpackbits(np.asarray(bits, dtype=np.uint8), axis=1, bitorder="little")[:, :B]

    def decode_packed(self, packed_blocks, B, rows=None, early_termination=None):
        """Host path for BER loops: ``packed_blocks`` = nibble-packed cluster indices (``pack_channel_values``;
        ideally in pinned memory, ``pinned_empty``), result = bit-packed hard decisions (cluster < |T|/2, the decoded
        bit of ``return_errors_all_zero`` and of the _enc drivers' comparison,
        WLAN/BER_simulation_OpenCL_enc.py:134) of the first ``rows`` rows (default: ``data_len`` for irregular
        decoders, all rows otherwise) as a (rows, ceil(B/8)) uint8 array owned by the decoder (overwritten by the
        next call).  Same kernels and stop rule as ``decode_OpenCL``; only the host formats differ."""
        h = self._upload_luts()
        early = self.early_termination if early_termination is None else early_termination
        pk = np.ascontiguousarray(packed_blocks, dtype=np.uint8)
        B = int(B)
        if pk.shape != (self.N_v, (B + 1) // 2):
            raise ValueError(f"expected packed shape {(self.N_v, (B + 1) // 2)}, got {pk.shape}")
        rows = (int(self.data_len) if self._irregular else self.N_v) if rows is None else int(rows)
        shape = (rows, (B + 7) // 8)
        if self._host_bits is None or self._host_bits.shape != shape:
            from ..device_array import pinned_empty
            self._host_bits = pinned_empty(shape, np.uint8)
        inum = C.c_int32(0)
        current_device()
        _lib.check(_lib.lib().ibldpc_decode_ib_host_packed(h, C.c_void_p(pk.ctypes.data), B, int(self.imax), int(bool(early)),
                                                           C.c_void_p(self._host_bits.ctypes.data), rows, C.byref(inum)))
        self.last_i_num = int(inum.value)
        return self._host_bits

    def return_errors_all_zero(self, varnode_output_buffer):
        """Number of decoded 1-bits, all-zero codeword assumed, over ALL rows
        (discrete_LDPC_decoder.py:297-300)."""
        return self._errors(varnode_output_buffer, self.N_v)

    def _errors(self, buf, rows):
        thr = int(self.cardinality_T_decoder_ops / 2)
        if isinstance(buf, np.ndarray):
            buf = DeviceArray(torch.from_numpy(np.ascontiguousarray(buf.astype(np.uint8))).cuda())
        bit, _frame = count_errors(buf, rows, thr)
        return bit

    def count_errors(self, varnode_output_buffer, ref_bits=None, rows=None):
        """(bit_errors, frame_errors); ``ref_bits`` (rows, B) device buffer of transmitted bits or
        None for the all-zero codeword (host comparison of WLAN/BER_simulation_OpenCL_enc.py:134)."""
        rows = (self.N_v if not self._irregular else int(self.data_len)) if rows is None else rows
        return count_errors(varnode_output_buffer, rows, int(self.cardinality_T_decoder_ops / 2), ref_bits)

    def decode_on_host(self, channel_values_):
        """Single frame, host vector in / host vector out (discrete_LDPC_decoder.py:357-400, which
        raises ValueError as shipped).  Runs imax-1 passes without early termination like the
        reference's loop (:374) -- on the GPU: this package has no CPU decoding path."""
        ch = np.asarray(channel_values_).reshape(-1)
        out = self.decode_OpenCL(ch[:, None], buffer_in=False, return_buffer=False, early_termination=False)
        return np.array(out[:, 0], dtype=np.float64)

    # ---- table-inspection helpers kept from the reference API -----------------------------------
    def discrete_cn_operation(self, vec_y_c, iter_):
        """One check-node look-up chain per row of ``vec_y_c`` (discrete_LDPC_decoder.py:302-335).
        Host-side numpy helper for inspecting tables; no decode path uses it."""
        Tc, T = int(self.cardinality_T_channel), int(self.cardinality_T_decoder_ops)
        C_ = self.Trellis_checknode_vector_a
        y = np.asarray(vec_y_c).astype(int)
        n_in = y.shape[1]
        if iter_ == 0:
            t = C_[y[:, 0] * Tc + y[:, 1]]
            for l in range(n_in - 2):
                t = C_[t * T + y[:, l + 2] + Tc ** 2 + l * T * Tc]
        else:
            off = (self.d_c_max - 3) * Tc * T + Tc ** 2 + (iter_ - 1) * (self.d_c_max - 2) * T ** 2
            t = y[:, 0]
            for l in range(n_in - 1):
                t = C_[off + l * T ** 2 + t * T + y[:, l + 1]]
        return t

    def discrete_vn_operation(self, vec_y_v, iter_):
        """One variable-node chain per row (column 0 = channel value) (discrete_LDPC_decoder.py:337-355)."""
        Tc, T = int(self.cardinality_T_channel), int(self.cardinality_T_decoder_ops)
        V_ = self.Trellis_varnode_vector_a
        y = np.asarray(vec_y_v).astype(int)
        off = (Tc * T + (self.d_v_max - 1) * T ** 2) * iter_
        t = V_[off + y[:, 0] * T + y[:, 1]]
        for l in range(y.shape[1] - 2):
            t = V_[off + Tc * T + l * T ** 2 + t * T + y[:, l + 2]]
        return t
