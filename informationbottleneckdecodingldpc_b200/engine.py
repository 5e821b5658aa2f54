"""Host-side plumbing shared by the decoder classes: graph loading, handle lifetime, buffer
conversion.  Mirrors what the reference spreads over ``__init__`` / ``init_OpenCL_decoding``
of its four decoder classes (discrete_LDPC_decoder.py:30-51,132-200 and the _irreg / min-sum /
BP twins); the device half of those methods is libibldpc.so.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .device_array import DeviceArray, as_tensor
from .graph import EdgeTables, alist_to_csr, code_rate_from_degrees, edge_tables, load_check_matrix


def current_device() -> int:
    """GPU of this process: torch's current device (one process per GPU; LOCAL_RANK under torchrun)."""
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device visible: the decoders run on the GPU only (no CPU fallback)")
    return torch.cuda.current_device()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class GraphDecoderBase:
    """Parity-check graph + libibldpc handle."""

    _handle = None
    _handle_device = None

    # ---- graph -------------------------------------------------------------------------
    def _load_graph(self, filename_or_matrix):
        if isinstance(filename_or_matrix, (str, os.PathLike)):
            H = load_check_matrix(str(filename_or_matrix))
        else:
            import scipy.sparse as sp
            H = sp.csr_matrix(filename_or_matrix)
            H.sort_indices()
        self.H_sparse = H
        t = edge_tables(H)
        self.tables: EdgeTables = t
        self.degree_checknode_nr = t.degree_chk.astype(np.int64)
        self.degree_varnode_nr = t.degree_var.astype(np.int64)
        self.N_v, self.N_c = t.n_var, t.n_chk
        self.d_c_max, self.d_v_max = t.d_c_max, t.d_v_max
        self.codeword_len = t.n_var
        self.inbox_memory_start_checknodes = t.inbox_start_chk.astype(np.int64)
        self.inbox_memory_start_varnodes = t.inbox_start_var.astype(np.int64)
        self.target_memory_cells_checknodes = t.target_cells_chk.astype(np.int64)
        self.target_memory_cells_varnodes = t.target_cells_var.astype(np.int64)
        self.customers_checknode_nr = t.var_of_chk_slot
        self.customers_varnode_nr = t.chk_of_var_slot

    def _set_rate(self):
        # set_code_parameters (discrete_LDPC_decoder_irreg.py:69-100)
        self.R_c = code_rate_from_degrees(self.H_sparse)
        self.data_len = (self.R_c * self.codeword_len).astype(int)

    @property
    def H(self):
        """Dense 0/1 matrix like the regular reference class keeps (built on demand only)."""
        return np.asarray(self.H_sparse.toarray(), dtype=np.int64)

    def alistToNumpy(self, lines):
        """AList (list of int lists) -> dense 0/1 array (discrete_LDPC_decoder.py:57-81)."""
        return np.asarray(alist_to_csr(lines).toarray(), dtype=np.int64)

    def load_check_mat(self, filename):
        return load_check_matrix(filename)

    # ---- handle ------------------------------------------------------------------------
    def _ensure_handle(self):
        dev = current_device()
        if self._handle is not None and self._handle_device == dev:
            return self._handle
        self._release()
        t = self.tables
        keep = [np.ascontiguousarray(a, dtype=np.int32) for a in
                (t.inbox_start_chk, t.degree_chk, t.target_cells_chk, t.inbox_start_var, t.degree_var, t.target_cells_var)]
        desc = _lib.CodeDesc(t.n_var, t.n_chk, t.n_edge, *[a.ctypes.data for a in keep])
        h = C.c_void_p()
        _lib.check(_lib.lib().ibldpc_create(C.byref(desc), dev, C.byref(h)))
        self._handle, self._handle_device = h, dev
        self._luts_uploaded = False
        return h

    def _release(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.lib().ibldpc_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        self._release()

    def info(self):
        """(fast_path, launches_of_last_decode, grid, dynamic_smem_bytes)"""
        w = (C.c_int32 * 4)()
        _lib.check(_lib.lib().ibldpc_info(self._ensure_handle(), w))
        return tuple(int(x) for x in w)

    # ---- buffers -----------------------------------------------------------------------
    def _device_input(self, received_blocks, torch_dtype):
        t = as_tensor(received_blocks)
        if t.dim() == 1:
            t = t[:, None]
        if t.shape[0] != self.N_v:
            raise ValueError(f"expected {self.N_v} rows (variable nodes), got {t.shape[0]}")
        if self._handle_device is not None and t.device.index != self._handle_device:
            raise ValueError(f"buffer lives on cuda:{t.device.index} but this decoder was initialised on "
                             f"cuda:{self._handle_device} (one decoder object per GPU)")
        if t.dtype != torch_dtype:
            if torch_dtype == torch.uint8 and t.numel():
                # cluster indices handed over as int32 / int64 like the reference's buffers: a value outside
                # [0, |T_channel|) must not wrap around in the uint8 cast (uint8 buffers are range-checked on the
                # device by the decode itself)
                lo, hi = int(t.min()), int(t.max())
                if lo < 0 or hi >= int(getattr(self, "cardinality_T_channel", 256)):
                    raise ValueError("channel cluster indices must lie in [0, cardinality_T_channel)")
            t = t.to(torch_dtype)
        return t.contiguous()

    # ---- i_num of the last decode, read back lazily ------------------------------------------
    @property
    def last_i_num(self):
        """The reference's ``i_num`` of the last decode.  Device-buffer decodes are asynchronous: the value is read
        back (one stream synchronisation) only when this attribute is looked at."""
        if getattr(self, "_inum_pending", False):
            v = C.c_int32(0)
            self._inum_pending = False
            _lib.check(_lib.lib().ibldpc_last_i_num(self._handle, C.byref(v)))
            self._last_i_num = int(v.value)
        return getattr(self, "_last_i_num", None)

    @last_i_num.setter
    def last_i_num(self, value):
        self._inum_pending = False
        self._last_i_num = value


def count_errors(buf, rows: int, threshold=None, ref_bits=None):
    """(bit_errors, frame_errors) over the first ``rows`` rows of a decoder output buffer.
    Cluster buffers (uint8): bit = value < threshold; LLR buffers (f32/f64): bit = value < 0."""
    t = as_tensor(buf)
    if t.dim() == 1:
        t = t[:, None]
    t = t.contiguous()
    B = t.shape[1]
    rows = int(min(rows, t.shape[0]))
    ref_ptr = None
    if ref_bits is not None:
        r = as_tensor(ref_bits).to(torch.uint8).contiguous()
        ref_ptr = C.c_void_p(r.data_ptr())
    cnt = (C.c_int64 * 2)()
    dev = t.device.index if t.device.index is not None else current_device()
    L = _lib.lib()
    if t.dtype == torch.uint8:
        _lib.check(L.ibldpc_count_errors_u8(dev, C.c_void_p(t.data_ptr()), rows, B, int(threshold), ref_ptr, cnt,
                                            C.c_void_p(stream_ptr())))
    elif t.dtype in (torch.float32, torch.float64):
        _lib.check(L.ibldpc_count_errors_llr(dev, C.c_void_p(t.data_ptr()), _lib.F32 if t.dtype == torch.float32 else _lib.F64,
                                             rows, B, ref_ptr, cnt, C.c_void_p(stream_ptr())))
    else:
        raise TypeError(f"unsupported buffer dtype {t.dtype}")
    return int(cnt[0]), int(cnt[1])


def count_errors_async(buf, rows: int, threshold: int, counters: torch.Tensor, ref_bits=None) -> None:
    """Asynchronous twin (cluster or LLR buffers): ``counters`` is an int64 CUDA tensor whose elements 0/1
    are incremented by the bit / frame errors on the current stream; nothing is read back."""
    t = as_tensor(buf)
    if t.dim() == 1:
        t = t[:, None]
    t = t.contiguous()
    if counters.dtype != torch.int64 or not counters.is_cuda or counters.numel() < 2:
        raise TypeError("count_errors_async needs an int64 CUDA counter tensor")
    ref_ptr = None
    if ref_bits is not None:
        r = as_tensor(ref_bits).to(torch.uint8).contiguous()
        ref_ptr = C.c_void_p(r.data_ptr())
    dev = t.device.index if t.device.index is not None else current_device()
    if t.dtype == torch.uint8:
        _lib.check(_lib.lib().ibldpc_count_errors_u8_async(dev, C.c_void_p(t.data_ptr()), int(min(rows, t.shape[0])), t.shape[1],
                                                           int(threshold), ref_ptr, C.c_void_p(counters.data_ptr()),
                                                           C.c_void_p(stream_ptr())))
    elif t.dtype in (torch.float32, torch.float64):   # LLR buffers: bit = (LLR < 0), `threshold` is ignored
        _lib.check(_lib.lib().ibldpc_count_errors_llr_async(dev, C.c_void_p(t.data_ptr()),
                                                            _lib.F32 if t.dtype == torch.float32 else _lib.F64,
                                                            int(min(rows, t.shape[0])), t.shape[1], ref_ptr,
                                                            C.c_void_p(counters.data_ptr()), C.c_void_p(stream_ptr())))
    else:
        raise TypeError(f"unsupported buffer dtype {t.dtype}")
