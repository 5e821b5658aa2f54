"""Systematic LDPC encoder with the interface of the reference's ``LDPCEncoder``
(Discrete_LDPC_decoding/LDPC_encoder.py), batched on the GPU (csrc/encoder.cu behind
``ibldpc_encoder_create`` / ``ibldpc_encode``).

Like the reference (and MATLAB's comm.LDPCEncoder it follows) the codeword is ``[x ; p]`` where the
parity bits solve ``H[:, K:] p = H[:, :K] x`` over GF(2); the last ``N-K`` columns of H must be
invertible.  The host analysis mirrors ``getLDPCEncoderParamters`` (LDPC_encoder.py:196-262):

* last part lower / upper triangular with full diagonal        -> 'Forward' / 'Backward Substitution'
* the same after reversing the row order (``RowOrder``)         -> ditto
* otherwise                                                     -> 'Matrix Inverse' (GF(2) elimination)

All substitution variants become one schedule "step t solves parity bit var[t] from equation eq[t]";
'Matrix Inverse' uploads the dense inverse of the last part (bit-packed rows).  Because the parity
vector is unique, the results equal the reference's codewords bit for bit.  Two notes on the reference,
both visible when it is run under numpy 2 (oracle/make_golden_encoder.py): its int8 ``EncodingMethod`` overflows the
column counter of ``GF2MatrixMul`` for more than 127 columns (:187), and its 'Backward Substitution' branch
sets the substitution direction to +1 (:236), which does not solve an upper-triangular system; this
implementation returns the valid codeword in both cases.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

from .. import _lib
from ..graph import load_check_matrix

_MAX_DENSE = 16384   # parity rows up to which the GF(2) elimination of 'Matrix Inverse' is attempted


# --------------------------------------------------------------------------------------------
# host analysis (pure numpy / scipy, no GPU): also used by the CPU test-suite
# --------------------------------------------------------------------------------------------
def is_full_diag_triangular(X: sp.spmatrix) -> int:
    """1: lower triangular with full diagonal, -1: upper triangular with full diagonal, else 0
    (isfulldiagtriangular, LDPC_encoder.py:337-356; a diagonal matrix counts as lower)."""
    X = sp.csr_matrix(X)
    n = X.shape[0]
    if X.shape[1] != n or not np.all(X.diagonal() != 0):
        return 0
    nnz = X.nnz
    if sp.tril(X).nnz == nnz:
        return 1
    if sp.triu(X).nnz == nnz:
        return -1
    return 0


def gf2_inverse_packed(X: np.ndarray) -> np.ndarray:
    """Inverse of a square 0/1 matrix over GF(2) by Gauss-Jordan elimination on bit-packed rows.
    Returns the inverse as (n, ceil(n/32)) uint32, bit k of row r = inv[r, k].  Raises ValueError when
    the matrix is singular (the reference prints 'Not invertible Matrix', LDPC_encoder.py:247)."""
    X = np.asarray(X, dtype=np.uint8) & 1
    n = X.shape[0]
    if X.shape != (n, n):
        raise ValueError("square matrix expected")
    aug = np.concatenate([X, np.eye(n, dtype=np.uint8)], axis=1)
    # pack along the columns into uint64 words (little-endian bit order inside a word)
    pad = (-aug.shape[1]) % 64
    aug = np.pad(aug, ((0, 0), (0, pad)))
    words = np.packbits(aug, axis=1, bitorder="little").view(np.uint64)     # (n, W)
    for col in range(n):
        w, b = divmod(col, 64)
        colbits = (words[:, w] >> np.uint64(b)) & np.uint64(1)
        cand = np.nonzero(colbits[col:])[0]
        if cand.size == 0:
            raise ValueError("the last N-K columns of the parity-check matrix are not invertible in GF(2)")
        piv = col + int(cand[0])
        if piv != col:
            words[[col, piv]] = words[[piv, col]]
            colbits[[col, piv]] = colbits[[piv, col]]
        rows = np.nonzero(colbits)[0]
        rows = rows[rows != col]
        if rows.size:
            words[rows] ^= words[col]
    bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")[:, n:2 * n]
    pad32 = (-n) % 32
    bits = np.pad(bits, ((0, 0), (0, pad32)))
    return np.ascontiguousarray(np.packbits(bits, axis=1, bitorder="little").view(np.uint32))


class EncoderPlan:
    """Result of the host analysis: algorithm name, CSR of H_first and either the substitution schedule or
    the dense inverse of the last part."""

    def __init__(self, H: sp.spmatrix):
        H = sp.csr_matrix(H)
        H.data = np.ones_like(H.data)
        H.sort_indices()
        M, N = H.shape
        K = N - M
        if K <= 0:
            raise ValueError("the parity-check matrix needs more columns than rows")
        self.N, self.K, self.M = N, K, M
        A = sp.csr_matrix(H[:, :K])
        A.sort_indices()
        self.a_rowptr = A.indptr.astype(np.int32)
        self.a_col = A.indices.astype(np.int32)
        last = sp.csr_matrix(H[:, K:])
        last.sort_indices()
        self.RowOrder = np.array([-1], dtype=np.int32)
        self.dense_inverse = None
        shape = is_full_diag_triangular(last)
        rows = np.arange(M)
        if shape == 0:
            rev = sp.csr_matrix(last[::-1, :])
            rshape = is_full_diag_triangular(rev)
            if rshape != 0:
                shape, last, rows = rshape, rev, rows[::-1].copy()
                self.RowOrder = np.arange(M, dtype=np.int32)[::-1].copy()
        if shape != 0:
            self.EncodingAlgorithm = "Forward Substitution" if shape == 1 else "Backward Substitution"
            self.EncodingMethod = 1 if shape == 1 else -1
            last = sp.csr_matrix(last)
            last.sort_indices()
            # row i of `last` (= equation rows[i] of H) has its diagonal at column i: it solves parity bit i
            order = np.arange(M) if shape == 1 else np.arange(M)[::-1]
            strict = sp.csr_matrix(sp.tril(last, -1) if shape == 1 else sp.triu(last, 1))
            strict.sort_indices()
            cnt = np.diff(strict.indptr)[order]
            self.eq = rows[order].astype(np.int32)
            self.var = order.astype(np.int32)
            self.oth_ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
            self.oth = (np.concatenate([strict.indices[strict.indptr[i]:strict.indptr[i + 1]] for i in order]).astype(np.int32)
                        if strict.nnz else np.zeros(0, dtype=np.int32))
            self.method = _lib.ENC_SUBSTITUTION
        else:
            self.EncodingAlgorithm = "Matrix Inverse"
            self.EncodingMethod = 0
            if M > _MAX_DENSE:
                raise NotImplementedError(f"'Matrix Inverse' encoding needs a dense GF(2) inverse of {M} x {M} bits; "
                                          f"codes with more than {_MAX_DENSE} parity bits must have a triangular last part")
            self.dense_inverse = gf2_inverse_packed(sp.csr_matrix(H[:, K:]).toarray())
            self.method = _lib.ENC_DENSE


class LDPCEncoder:
    """``LDPCEncoder(filename, alist_file=True)`` -- constructor and attributes of the reference class
    (LDPC_encoder.py:15-38); ``filename`` may also be a scipy sparse / dense matrix."""

    _handle = None

    def __init__(self, filename, alist_file=True):
        if isinstance(filename, (str, os.PathLike)):
            self.H_sparse = load_check_matrix(str(filename))
        else:
            self.H_sparse = sp.csr_matrix(filename)
        self.setParityCheckMatrix(self.H_sparse)

    # ---- reference-named host methods ---------------------------------------------------------
    def setParityCheckMatrix(self, H):
        self.getLDPCEncoderParamters(H)
        self.storedParityCheckMatrix = H

    def getLDPCEncoderParamters(self, H):
        self._release()
        plan = EncoderPlan(H)
        self.plan = plan
        self.N, self.K = plan.N, plan.K
        self.NumInfoBits, self.NumParityBits, self.BlockLength = plan.K, plan.M, plan.N
        self.EncodingAlgorithm = plan.EncodingAlgorithm
        self.EncodingMethod = np.int64(plan.EncodingMethod)
        self.RowOrder = plan.RowOrder

    isfulldiagtriangular = staticmethod(is_full_diag_triangular)

    # ---- device ---------------------------------------------------------------------------------
    def _ensure_handle(self):
        from ..engine import current_device
        dev = current_device()
        if self._handle is None:
            p = self.plan
            keep = [p.a_rowptr, p.a_col]
            d = _lib.EncoderDesc(n_var=p.N, n_info=p.K, method=p.method,
                                 a_rowptr=p.a_rowptr.ctypes.data, a_col=p.a_col.ctypes.data)
            if p.method == _lib.ENC_SUBSTITUTION:
                d.eq, d.var, d.oth_ptr, d.oth = (p.eq.ctypes.data, p.var.ctypes.data, p.oth_ptr.ctypes.data,
                                                 p.oth.ctypes.data if p.oth.size else None)
                keep += [p.eq, p.var, p.oth_ptr, p.oth]
            else:
                d.dense_inverse = p.dense_inverse.ctypes.data
                keep.append(p.dense_inverse)
            h = C.c_void_p()
            _lib.check(_lib.lib().ibldpc_encoder_create(C.byref(d), dev, C.byref(h)))
            self._handle, self._handle_device = h, dev
        elif self._handle_device != dev:
            raise RuntimeError("encoder handle belongs to another CUDA device")
        return self._handle

    def _release(self):
        if self._handle is not None:
            try:
                _lib.lib().ibldpc_encoder_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        self._release()

    def encode_batch(self, bits):
        """(K, B) information bits (DeviceArray / CUDA tensor / numpy, any integer dtype, values 0/1) ->
        DeviceArray (N, B) uint8 codewords on the GPU."""
        import torch
        from ..device_array import DeviceArray, as_tensor
        from ..engine import current_device, stream_ptr
        dev = current_device()
        if isinstance(bits, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(bits).astype(np.uint8)).to(f"cuda:{dev}")
        else:
            t = as_tensor(bits).to(torch.uint8)
        if t.dim() == 1:
            t = t.reshape(-1, 1)
        if t.shape[0] != self.K:
            raise ValueError(f"expected {self.K} information bits per frame, got {t.shape[0]}")
        t = t.contiguous()
        out = torch.empty((self.N, t.shape[1]), dtype=torch.uint8, device=t.device)
        _lib.check(_lib.lib().ibldpc_encode(self._ensure_handle(), C.c_void_p(t.data_ptr()), int(t.shape[1]),
                                            C.c_void_p(out.data_ptr()), C.c_void_p(stream_ptr())))
        return DeviceArray(out)

    def encode(self, X):
        """One frame, numpy in -> numpy codeword out (LDPC_encoder.py:86-123), computed on the GPU."""
        X = np.asarray(X)
        return self.encode_batch(X.reshape(-1, 1)).get()[:, 0].astype(X.dtype if X.dtype.kind in "iu" else np.int64)

    encode_c = encode   # the reference's Cython-accelerated twin (LDPC_encoder.py:125-163)
