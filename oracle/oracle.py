"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/README.md).

Two back-ends with one calling convention:

* ``port``      -- oracle/ldpc_oracle.c, this repository's C restatement of the reference kernels;
* ``reference`` -- oracle/_ref/libref_kernels.so, the reference's own ``.cl`` sources compiled
                   for the CPU by oracle/build_ref.py (exists only where it was built).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may import
this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SRC = os.path.join(HERE, "ldpc_oracle.c")
PORT_LIB = os.path.join(HERE, "libldpc_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libref_kernels.so")

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build_port(force: bool = False) -> str:
    if force or not os.path.exists(PORT_LIB) or os.path.getmtime(PORT_LIB) < os.path.getmtime(PORT_SRC):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC",
                               PORT_SRC, "-o", PORT_LIB, "-lm"])
    return PORT_LIB


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        lib = C.CDLL(build_port())
        g = [C.c_int] * 3 + [_i32p] * 6
        lib.oracle_ib_decode.argtypes = g + [C.c_int] * 6 + [_i32p, _i32p, C.c_void_p, C.c_void_p,
                                                             _i32p, C.c_int, C.c_int, _i32p]
        lib.oracle_ib_decode.restype = C.c_int
        lib.oracle_llr_decode.argtypes = g + [C.c_int, C.c_int, _f64p, C.c_int, C.c_int, _f64p]
        lib.oracle_llr_decode.restype = C.c_int
        lib.oracle_quantize.argtypes = [C.c_int, _f64p, _f64p, C.c_long, _i32p]
        lib.oracle_quantize_llr.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_long, _f64p]
        _port = lib
    return _port


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_LIB)
        g = [C.c_int] * 3 + [_i32p] * 6
        lib.ref_ib_decode.argtypes = [C.c_int] * 4 + g + [C.c_int] * 3 + [_i32p, _i32p, C.c_void_p, C.c_void_p,
                                                                          _i32p, C.c_int, C.c_int, _i32p]
        lib.ref_ib_decode.restype = C.c_int
        lib.ref_llr_decode.argtypes = [C.c_int] * 3 + g + [C.c_int, _f64p, C.c_int, C.c_int, _f64p]
        lib.ref_llr_decode.restype = C.c_int
        lib.ref_quantize.argtypes = [C.c_int, _f64p, _f64p, C.c_int, C.c_int, _i32p]
        lib.ref_quantize_llr.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_int, _f64p]
        lib.ref_num_threads.restype = C.c_int
        _ref = lib
    return _ref


def _graph_args(t):
    c = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return (int(t.n_var), int(t.n_chk), int(t.n_edge), c(t.inbox_start_chk), c(t.degree_chk),
            c(t.target_cells_chk), c(t.inbox_start_var), c(t.degree_var), c(t.target_cells_var))


def _opt(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(np.asarray(a).astype(np.int32))
    return a, a.ctypes.data_as(C.c_void_p)


def ib_decode(tables, ch, *, T, imax, cn_lut, vn_lut, cn_match=None, vn_match=None, Tc=None,
              DC=None, DV=None, early=False, backend="port", irregular=True):
    """Run the IB decoder on int cluster indices ``ch`` (N, B).  Returns (out int32 (N,B), i_num)."""
    Tc = T if Tc is None else Tc
    DC = tables.d_c_max if DC is None else DC
    DV = tables.d_v_max if DV is None else DV
    ch = np.ascontiguousarray(np.asarray(ch).astype(np.int32))
    if ch.ndim == 1:
        ch = ch[:, None]
    B = ch.shape[1]
    out = np.empty_like(ch)
    Cn = np.ascontiguousarray(np.asarray(cn_lut).astype(np.int32))
    Vn = np.ascontiguousarray(np.asarray(vn_lut).astype(np.int32))
    match = cn_match is not None
    mc_keep, mc = _opt(cn_match)
    mv_keep, mv = _opt(vn_match)
    if backend == "port":
        r = port_lib().oracle_ib_decode(*_graph_args(tables), DC, DV, Tc, T, imax, int(match), Cn, Vn, mc, mv,
                                        ch, B, int(early), out)
    elif backend == "reference":
        if not irregular and match:
            raise ValueError("the regular reference kernels have no matching")
        r = ref_lib().ref_ib_decode(int(irregular), DC, DV, int(match), *_graph_args(tables), Tc, T, imax,
                                    Cn, Vn, mc, mv, ch, B, int(early), out)
        if r == -1000:
            raise KeyError(f"oracle/_ref was not built for irregular={irregular} DC={DC} DV={DV} match={match}")
    else:
        raise ValueError(backend)
    if r < 0:
        raise MemoryError("oracle allocation failed")
    return out, int(r)


def llr_decode(tables, ch, *, algo, imax, early=False, backend="port", DC=None, DV=None):
    """Min-sum (algo='minsum') or BP (algo='bp') on float64 LLRs (N, B).  Returns (out f64, i_num)."""
    a = {"minsum": 0, "bp": 1}[algo]
    ch = np.ascontiguousarray(np.asarray(ch, dtype=np.float64))
    if ch.ndim == 1:
        ch = ch[:, None]
    B = ch.shape[1]
    out = np.empty_like(ch)
    if backend == "port":
        r = port_lib().oracle_llr_decode(*_graph_args(tables), a, imax, ch, B, int(early), out)
    else:
        DC = tables.d_c_max if DC is None else DC
        DV = tables.d_v_max if DV is None else DV
        r = ref_lib().ref_llr_decode(DC, DV, a, *_graph_args(tables), imax, ch, B, int(early), out)
        if r == -1000:
            raise KeyError(f"oracle/_ref was not built for LLR DC={DC} DV={DV}")
    return out, int(r)


def quantize(x, limits, card, backend="port"):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    lim = np.ascontiguousarray(np.asarray(limits, dtype=np.float64))
    out = np.empty(x.shape, dtype=np.int32)
    if backend == "port":
        port_lib().oracle_quantize(card, x.ravel(), lim, x.size, out.reshape(-1))
    else:
        ref_lib().ref_quantize(card, x.ravel(), lim, x.shape[0], x.shape[1], out.reshape(-1))
    return out


def quantize_llr(x, limits, card, llr_values, backend="port"):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    lim = np.ascontiguousarray(np.asarray(limits, dtype=np.float64))
    lv = np.ascontiguousarray(np.asarray(llr_values, dtype=np.float64))
    out = np.empty(x.shape, dtype=np.float64)
    if backend == "port":
        port_lib().oracle_quantize_llr(card, x.ravel(), lim, lv, x.size, out.reshape(-1))
    else:
        ref_lib().ref_quantize_llr(card, x.ravel(), lim, lv, x.shape[0], x.shape[1], out.reshape(-1))
    return out


def gf2_solve(Hlast, S):
    """Solve Hlast P = S over GF(2) by plain Gauss-Jordan elimination on 0/1 arrays (Hlast square and
    invertible, S one column per frame).  Independent of the product's host analysis."""
    A = (np.asarray(Hlast, dtype=np.uint8) & 1).copy()
    P = (np.asarray(S, dtype=np.uint8) & 1).copy()
    n = A.shape[0]
    for col in range(n):
        cand = np.nonzero(A[col:, col])[0]
        if cand.size == 0:
            raise ValueError("singular over GF(2)")
        piv = col + int(cand[0])
        if piv != col:
            A[[col, piv]] = A[[piv, col]]
            P[[col, piv]] = P[[piv, col]]
        rows = np.nonzero(A[:, col])[0]
        rows = rows[rows != col]
        A[rows] ^= A[col]
        P[rows] ^= P[col]
    return P


def encode(H, bits):
    """Systematic encoding as the reference's LDPCEncoder defines it (Discrete_LDPC_decoding/LDPC_encoder.py:15-21,
    :86-123): codeword = [x ; p] with H[:, K:] p = H[:, :K] x over GF(2).  ``bits`` (K, B) -> (N, B) uint8.
    The parity vector is unique, so forward / backward substitution and the GF(2) factorisation of the
    reference all produce this result."""
    import scipy.sparse as sp
    H = sp.csr_matrix(H)
    M, N = H.shape
    K = N - M
    x = (np.asarray(bits, dtype=np.int64) & 1).reshape(K, -1)
    s = (H[:, :K].astype(np.int64) @ x) & 1
    p = gf2_solve(H[:, K:].toarray(), s.astype(np.uint8))
    return np.concatenate([x.astype(np.uint8), p.astype(np.uint8)], axis=0)
