"""AWGN channel-output quantizer for BPSK -- B200 back-end.

Drop-in for ``AWGN_Channel_Transmission/AWGN_Quantizer_BPSK.py`` of the reference: same
constructor and ``init_OpenCL_quanti`` / ``quantize_OpenCL`` / ``quantize_direct_OpenCL`` /
``quantize_direct_OpenCL_LLR`` / ``quantize_on_host`` / ``quantize_direct`` methods, same attributes
(``limits``, ``cdf_t_given_x_equals_zero``, ``output_LLRs``, ``p_x_and_t``, ``context``).
Device work (``quantize``, ``quantize_LLR`` kernels, direct sampling) runs in libibldpc.so; the
uniform numbers of the direct methods are drawn on the GPU (Philox) instead of
``np.random.rand`` + H2D (AWGN_Quantizer_BPSK.py:210-214).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from scipy.stats import norm

from .. import _lib
from ..design import symmetric_mi_quantizer
from ..device_array import DeviceArray, as_tensor
from ..engine import current_device, stream_ptr
from ..rng import PhiloxStream


class AWGN_Channel_Quantizer(PhiloxStream):
    """Reference signature: AWGN_Quantizer_BPSK.py:46."""

    def __init__(self, sigma_n2_, AD_max_abs_, cardinality_T_, cardinality_Y_, dont_calc=False):
        self.nror = 5
        self.limits = np.zeros(cardinality_T_)
        self.sigma_n2 = sigma_n2_
        self.cardinality_T = cardinality_T_
        self.cardinality_Y = cardinality_Y_
        self.AD_max_abs = AD_max_abs_
        self.y_vec = np.linspace(-self.AD_max_abs, +self.AD_max_abs, self.cardinality_Y)
        self.x_vec = np.array([-1, 1])
        self.delta = self.y_vec[1] - self.y_vec[0]
        self.seed = 20181001          # Philox seed of the direct-sampling methods (key = rng.stream_key(seed, stream))
        self._offset = 0              # Philox counter: advances by N_var*msg_at_time per call
        self._stream = None           # sub-stream: set_stream(rank); default = rank of the process group, else 0
        self.llr_dtype = np.float64   # dtype of quantize_direct_OpenCL_LLR buffers (np.float32 = fast path)
        self.return_buffer_only = False
        self.context = None
        if not dont_calc:
            self.calc_quanti()

    # ---- design (host, one-off per Eb/N0; AWGN_Quantizer_BPSK.py:62-124) ----------------------
    def calc_quanti(self):
        p0 = norm.pdf(self.y_vec, loc=1, scale=np.sqrt(self.sigma_n2)) * self.delta
        p0[-1] += self.gaussian_over_prob(self.AD_max_abs, 1)
        p0[0] += self.gaussian_under_prob(-self.AD_max_abs, 1)
        p1 = p0[::-1]
        self.p_xy = 0.5 * np.hstack((p0[:, np.newaxis], p1[:, np.newaxis]))
        self.p_xy = self.p_xy / self.p_xy.sum()
        # ib_base's symmetric_sIB replaced by a deterministic symmetric max-MI design (design.py)
        self.p_t_given_y, self.p_x_given_t, self.p_t = symmetric_mi_quantizer(self.p_xy, self.cardinality_T)
        self.p_x_given_t = self.p_x_given_t / self.p_x_given_t.sum(1)[:, np.newaxis]
        self.p_x_and_t = self.p_x_given_t * self.p_t[:, np.newaxis]
        p_t_given_x_equals_zero = self.p_x_and_t[:, 0] / 0.5
        self.cdf_t_given_x_equals_zero = np.append([0], np.cumsum(p_t_given_x_equals_zero))
        self.output_LLRs = np.log(self.p_x_and_t[:, 0] / self.p_x_and_t[:, 1])
        self.calc_limits()

    def gaussian_over_prob(self, x, mu):
        return norm.sf((x - mu + self.delta / 2) / np.sqrt(self.sigma_n2))

    def gaussian_under_prob(self, x, mu):
        return 1 - self.gaussian_over_prob(x - self.delta, mu)

    def calc_limits(self):
        for i in range(self.cardinality_T):
            cur = (self.p_t_given_y[:, i] == 1).nonzero()
            self.limits[i] = self.y_vec[cur[0].min()]
        self.limits[int(self.cardinality_T / 2)] = 0

    # ---- host twins (AWGN_Quantizer_BPSK.py:126-154), numpy, used by the _enc drivers ---------
    def quantize_direct(self, input_bits):
        rand_u = np.random.rand(input_bits.shape[0], input_bits.shape[1])
        out = ((rand_u[:, :, np.newaxis] - self.cdf_t_given_x_equals_zero) > 0).sum(2) - 1
        flip = input_bits.astype(bool)
        out[flip] = self.cardinality_T - 1 - out[flip]
        return out

    def quantize_on_host(self, x):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[:, None]
        cluster = ((x[:, :, np.newaxis] - self.limits) > 0).sum(2) - 1
        cluster[cluster == -1] = 0
        return cluster

    # ---- device side ---------------------------------------------------------------------------
    def init_OpenCL_quanti(self, N_var, msg_at_time, return_buffer_only=False):
        """AWGN_Quantizer_BPSK.py:156-181.  ``context`` is a placeholder object so that drivers can
        keep passing ``quanti.context`` to ``init_OpenCL_decoding``."""
        self.device = current_device()
        _lib.lib()
        self.context = ("cuda", self.device)
        self.return_buffer_only = return_buffer_only
        self._N_var, self._msg_at_time = N_var, msg_at_time

    init_quanti = init_OpenCL_quanti

    def _ret(self, t):
        return DeviceArray(t) if self.return_buffer_only else t.cpu().numpy()

    def quantize_OpenCL(self, x):
        """AWGN_Quantizer_BPSK.py:183-199: cluster indices of received samples x (numpy or device)."""
        dev = current_device()
        if isinstance(x, np.ndarray):
            xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
        else:
            xt = as_tensor(x).to(torch.float64).contiguous()
        out = torch.empty(xt.shape, dtype=torch.uint8, device=xt.device)
        lim = np.ascontiguousarray(self.limits, dtype=np.float64)
        _lib.check(_lib.lib().ibldpc_quantize(dev, C.c_void_p(xt.data_ptr()), xt.numel(), C.c_void_p(lim.ctypes.data),
                                              int(self.cardinality_T), C.c_void_p(out.data_ptr()), C.c_void_p(stream_ptr())))
        if self.return_buffer_only:
            return DeviceArray(out)
        return out.cpu().numpy().astype(np.int32)

    def quantize_direct_OpenCL(self, N_var, msg_at_time):
        """AWGN_Quantizer_BPSK.py:201-228: cluster indices ~ p(t|x=0) by the inversion method,
        (N_var, msg_at_time) uint8, all-zero codeword."""
        dev = current_device()
        n = int(N_var) * int(msg_at_time)
        out = torch.empty((int(N_var), int(msg_at_time)), dtype=torch.uint8, device=f"cuda:{dev}")
        cdf = np.ascontiguousarray(self.cdf_t_given_x_equals_zero, dtype=np.float64)
        _lib.check(_lib.lib().ibldpc_sample_direct(dev, C.c_void_p(cdf.ctypes.data), int(self.cardinality_T) + 1,
                                                   int(self._philox_key()), int(self._offset), n, C.c_void_p(out.data_ptr()),
                                                   C.c_void_p(stream_ptr())))
        self._offset += n
        if self.return_buffer_only:
            return DeviceArray(out)
        return out.cpu().numpy().astype(np.int32)

    def quantize_direct_OpenCL_LLR(self, N_var, msg_at_time):
        """AWGN_Quantizer_BPSK.py:230-248: ``output_LLRs[cluster]`` of the same draws."""
        dev = current_device()
        n = int(N_var) * int(msg_at_time)
        f32 = np.dtype(self.llr_dtype) == np.float32
        out = torch.empty((int(N_var), int(msg_at_time)), dtype=torch.float32 if f32 else torch.float64,
                          device=f"cuda:{dev}")
        cdf = np.ascontiguousarray(self.cdf_t_given_x_equals_zero, dtype=np.float64)
        llr = np.ascontiguousarray(np.append(self.output_LLRs, self.output_LLRs[-1]), dtype=np.float64)
        _lib.check(_lib.lib().ibldpc_sample_direct_llr(dev, C.c_void_p(cdf.ctypes.data), int(self.cardinality_T) + 1,
                                                       C.c_void_p(llr.ctypes.data), int(self._philox_key()), int(self._offset), n,
                                                       _lib.F32 if f32 else _lib.F64, C.c_void_p(out.data_ptr()),
                                                       C.c_void_p(stream_ptr())))
        self._offset += n
        if self.return_buffer_only:
            return DeviceArray(out)
        return out.cpu().numpy().astype(np.float64)
