#!/usr/bin/env python3
"""Generate tests/golden/enc_*.npz from the REFERENCE encoder (run in the build container only).

TEST INFRASTRUCTURE ONLY.  The reference's ``LDPCEncoder`` (Discrete_LDPC_decoding/LDPC_encoder.py) is imported
from /root/reference with a stub for its un-built Cython extension (``GF2MatrixMul_c``; the pure-Python
``encode`` / ``GF2MatrixMul`` path is used), ``np.int = int``, and ``EncodingMethod`` widened from int8 to int64
after construction (under numpy 2 the int8 turns the column counter of GF2MatrixMul into an int8 that overflows
after 127 columns, LDPC_encoder.py:187).  Its codewords are frozen together with the oracle's (oracle.encode) after
asserting that both agree and satisfy H c = 0.  The reference's 'Backward Substitution' branch does not produce
codewords (substitution direction +1 for an upper-triangular system, :236); that case is recorded with the oracle's
result only and flagged ``reference_valid = 0``.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")

from informationbottleneckdecodingldpc_b200 import codes  # noqa: E402
from oracle import oracle  # noqa: E402


def reference_encoder_module():
    np.int = int
    stub = types.ModuleType("Discrete_LDPC_decoding.GF2MatrixMul_c")
    pkg = types.ModuleType("Discrete_LDPC_decoding")
    pkg.GF2MatrixMul_c = stub
    sys.modules["Discrete_LDPC_decoding"] = pkg
    sys.modules["Discrete_LDPC_decoding.GF2MatrixMul_c"] = stub
    spec = importlib.util.spec_from_file_location("ref_ldpc_encoder", os.path.join(REF, "Discrete_LDPC_decoding", "LDPC_encoder.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def case(mod, name, H, B, seed):
    H = sp.csr_matrix(H)
    H.sort_indices()
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, name + ".npz")
    codes.save_csr_npz(H, path)
    with np.errstate(all="ignore"):
        enc = mod.LDPCEncoder(path)
    enc.EncodingMethod = np.int64(enc.EncodingMethod)
    M, N = H.shape
    K = N - M
    x = np.random.Generator(np.random.PCG64(seed)).integers(0, 2, size=(K, B))
    cw_oracle = oracle.encode(H, x)
    assert not (H.astype(np.int64) @ cw_oracle.astype(np.int64) % 2).any() and np.array_equal(cw_oracle[:K], x)
    cw_ref = np.stack([np.asarray(enc.encode(x[:, b].copy())) for b in range(B)], axis=1).astype(np.uint8)
    valid = not (H.astype(np.int64) @ cw_ref.astype(np.int64) % 2).any()
    if valid:
        assert np.array_equal(cw_ref, cw_oracle), f"{name}: oracle != reference"
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), H_indptr=H.indptr.astype(np.int32),
                        H_indices=H.indices.astype(np.int32), H_shape=np.array(H.shape, dtype=np.int64),
                        bits=x.astype(np.uint8), codeword=cw_oracle, reference_valid=int(valid),
                        algorithm=str(enc.EncodingAlgorithm), row_order_reversed=int(enc.RowOrder[0] >= 0 and enc.EncodingAlgorithm != "Matrix Inverse"))
    print(f"{name}: N={N} K={K} B={B} reference algorithm '{enc.EncodingAlgorithm}' reference codewords valid: {valid}")


def main():
    if not os.path.isdir(REF):
        sys.exit("the reference tree is needed to (re)generate golden vectors")
    mod = reference_encoder_module()
    Hw = codes.wlan_80211n(54)
    Hd = codes.dvbs2_like_half_rate(6480, q_groups=36)
    A = Hd.toarray()
    K = A.shape[1] - A.shape[0]
    case(mod, "enc_wlan1296_matrix_inverse", Hw, 6, 600)
    case(mod, "enc_wlan1944_matrix_inverse", codes.wlan_80211n(81), 3, 601)
    case(mod, "enc_dvb6480_forward", Hd, 6, 602)
    case(mod, "enc_dvb6480_rows_reversed", A[::-1].copy(), 5, 603)
    case(mod, "enc_dvb6480_backward", np.concatenate([A[::-1, :K], A[::-1, K:][:, ::-1]], axis=1), 4, 604)
    # a small dense-ish code whose last part needs the factorisation
    rng = np.random.Generator(np.random.PCG64(7))
    while True:
        Hs = (rng.random((40, 100)) < 0.12).astype(np.uint8)
        try:
            oracle.gf2_solve(Hs[:, 60:], np.zeros((40, 1), dtype=np.uint8))
            break
        except ValueError:
            continue
    case(mod, "enc_random100_matrix_inverse", Hs, 9, 605)


if __name__ == "__main__":
    main()
