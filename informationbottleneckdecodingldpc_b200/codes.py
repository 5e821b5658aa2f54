"""Parity-check matrix generators for the three code families of BASELINE.json.

The reference ships no H files; the drivers open
``LDPC_codes/...`` which is absent.  These generators produce codes with the
degree profiles the reference's scripts assume:

* ``regular_random``     -- (d_v, d_c)-regular, e.g. the (3,6) n=8000 code standing in for
  MacKay's ``8000.4000.3.483`` (Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py:35);
* ``qc_expand`` / ``wlan_80211n`` -- quasi-cyclic expansion of an 802.11n prototype matrix
  (same construction as Irregular_LDPC_Decoding/WLAN/generate_802.11_matrix.py:21-33,
  i.e. block (i,j) with shift s is the identity cyclically shifted right by s columns);
* ``dvbs2_like_half_rate`` -- IRA code with the exact node-degree profile of the DVB-S2
  rate-1/2 normal frame used in Irregular_LDPC_Decoding/DVB-S2/decoder_config_generation.py:32-34
  (d_c = {6:1, 7:32399}, d_v = {1:1, 2:32399, 3:19440, 8:12960}).  The ETSI address
  tables are not available offline, so the accumulator addresses are drawn from a seeded
  generator: "DVB-S2-like", same sizes/degrees, not the standardised matrix.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

# IEEE 802.11n-2009 Annex R, rate 1/2 prototype matrices (-1 = zero block).
# Z=54 (n=1296) is the matrix the reference generates; Z=81 (n=1944) is the
# block length BASELINE.json names.  The Z=81 table is reproduced from the
# standard from memory (it is not in the reference); only its degree profile
# matters for throughput, and parity tests are pinned on Z=54.
_WLAN_R12 = {
    54: """
40 -1 -1 -1 22 -1 49 23 43 -1 -1 -1  1  0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1
50  1 -1 -1 48 35 -1 -1 13 -1 30 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1 -1
39 50 -1 -1  4 -1  2 -1 -1 -1 -1 49 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1
33 -1 -1 38 37 -1 -1  4  1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1
45 -1 -1 -1  0 22 -1 -1 20 42 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1
51 -1 -1 48 35 -1 -1 -1 44 -1 18 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1
47 11 -1 -1 -1 17 -1 -1 51 -1 -1 -1  0 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1
 5 -1 25 -1  6 -1 45 -1 13 40 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1
33 -1 -1 34 24 -1 -1 -1 23 -1 -1 46 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1
 1 -1 27 -1  1 -1 -1 -1 38 -1 44 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1
-1 18 -1 -1 23 -1 -1  8  0 35 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0
49 -1 17 -1 30 -1 -1 -1 34 -1 -1 19  1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0
""",
    81: """
57 -1 -1 -1 50 -1 11 -1 50 -1 79 -1  1  0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1
 3 -1 28 -1  0 -1 -1 -1 55  7 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1 -1
30 -1 -1 -1 24 37 -1 -1 56 14 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1
62 53 -1 -1 53 -1 -1  3 35 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1
40 -1 -1 20 66 -1 -1 22 28 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1
 0 -1 -1 -1  8 -1 42 -1 50 -1 -1  8 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1
69 79 79 -1 -1 -1 56 -1 52 -1 -1 -1  0 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1
65 -1 -1 -1 38 57 -1 -1 72 -1 27 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1
64 -1 -1 -1 14 52 -1 -1 30 -1 -1 32 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1
-1 45 -1 70  0 -1 -1 -1 77  9 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1
 2 56 -1 57 35 -1 -1 -1 -1 -1 12 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0
24 -1 61 -1 60 -1 -1 27 51 -1 -1 16  1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0
""",
}


def wlan_prototype(Z: int = 54) -> np.ndarray:
    return np.array([[int(t) for t in row.split()] for row in _WLAN_R12[Z].strip().splitlines()],
                    dtype=np.int64)


def qc_expand(proto: np.ndarray, Z: int) -> sp.csr_matrix:
    """Expand a prototype (shift) matrix: entry s>=0 -> ZxZ identity rolled right by s."""
    rows, cols = [], []
    base = np.arange(Z)
    for i in range(proto.shape[0]):
        for j in range(proto.shape[1]):
            s = int(proto[i, j])
            if s >= 0:
                rows.append(i * Z + base)
                cols.append(j * Z + (base + s) % Z)
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    H = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)),
                      shape=(proto.shape[0] * Z, proto.shape[1] * Z))
    H.sort_indices()
    return H


def wlan_80211n(Z: int = 54) -> sp.csr_matrix:
    """802.11n rate-1/2 code: Z=54 -> n=1296 (the reference's matrix), Z=81 -> n=1944."""
    return qc_expand(wlan_prototype(Z), Z)


def random_from_degrees(deg_var, deg_chk, seed: int = 0) -> sp.csr_matrix:
    """Simple bipartite graph with prescribed node degrees (configuration model); parallel
    edges are repaired by swapping the variable socket of one of them with a random other
    edge until the graph is simple."""
    deg_var = np.asarray(deg_var, dtype=np.int64)
    deg_chk = np.asarray(deg_chk, dtype=np.int64)
    if deg_var.sum() != deg_chk.sum():
        raise ValueError("socket counts differ")
    n, m = deg_var.size, deg_chk.size
    rng = np.random.Generator(np.random.PCG64(seed))
    chk = np.repeat(np.arange(m), deg_chk)
    var = rng.permutation(np.repeat(np.arange(n), deg_var))
    for _round in range(10000):
        key = chk.astype(np.int64) * n + var
        order = np.argsort(key, kind="stable")
        dup = order[1:][key[order][1:] == key[order][:-1]]
        if dup.size == 0:
            break
        for e in dup:
            o = int(rng.integers(0, var.size))
            var[e], var[o] = var[o], var[e]
    else:  # pragma: no cover
        raise RuntimeError("could not remove double edges")
    H = sp.csr_matrix((np.ones(var.size, dtype=np.int8), (chk, var)), shape=(m, n))
    H.sort_indices()
    assert H.nnz == deg_var.sum() and int(H.data.max()) == 1
    return H


def regular_random(n: int = 8000, d_v: int = 3, d_c: int = 6, seed: int = 20181001) -> sp.csr_matrix:
    """(d_v,d_c)-regular code from a random socket matching (see ``random_from_degrees``)."""
    if (n * d_v) % d_c:
        raise ValueError("n*d_v must be a multiple of d_c")
    return random_from_degrees(np.full(n, d_v), np.full(n * d_v // d_c, d_c), seed)


# kept as the documented name of the C1 stand-in
regular_gallager = regular_random


def dvbs2_like_half_rate(n: int = 64800, seed: int = 20181001, q_groups: int = 360) -> sp.csr_matrix:
    """IRA code with the DVB-S2 rate-1/2 normal-frame degree profile (see module docstring).

    Layout (as in the standard): columns [0,K) information bits in groups of 360, columns
    [K,N) the accumulator (staircase) parity bits.  Column m of a group with base addresses
    x_j connects rows (x_j + q*m) mod (N-K), q = (N-K)/360.  Base addresses are drawn so
    every residue class mod q receives exactly 5 of them -> every check sees 5 information
    edges + 2 staircase edges (the first check: 1).
    Scaled-down instances (n a multiple of 2*360*... ) keep the same proportions.
    """
    k = n // 2
    m = n - k
    if k % q_groups or m % q_groups:
        raise ValueError("n/2 must be a multiple of the group size")
    q = m // q_groups
    n_groups = k // q_groups
    # 40 % of the information groups have degree 8, 60 % degree 3 (36 / 54 of 90).
    g8 = (n_groups * 2) // 5
    g3 = n_groups - g8
    deg = np.array([8] * g8 + [3] * g3)
    total = int(deg.sum())
    if total % q:
        raise ValueError("degree total must be a multiple of q")
    per_class = total // q
    rng = np.random.Generator(np.random.PCG64(seed))
    for _attempt in range(200):
        residues = np.repeat(np.arange(q), per_class)
        addr = residues + q * rng.integers(0, q_groups, size=total)
        addr = rng.permutation(addr)
        ok = True
        groups = []
        pos = 0
        for d in deg:
            a = addr[pos:pos + d]
            pos += d
            if np.unique(a).size != d:
                ok = False
                break
            groups.append(a)
        if ok:
            break
    else:  # pragma: no cover
        raise RuntimeError("could not draw distinct accumulator addresses")
    rows, cols = [], []
    mm = np.arange(q_groups)
    for g, a in enumerate(groups):
        for x in a:
            rows.append((x + q * mm) % m)
            cols.append(g * q_groups + mm)
    # staircase: parity j sits in checks j and j+1
    pj = np.arange(m)
    rows.append(pj)
    cols.append(k + pj)
    rows.append(pj[:-1] + 1)
    cols.append(k + pj[:-1])
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    H = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(m, n))
    H.sum_duplicates()
    H.data[:] = 1
    H.sort_indices()
    return H


def save_csr_npz(H: sp.spmatrix, filename: str) -> None:
    """Write the ``.npz`` CSR container ``load_sparse_csr`` expects
    (discrete_LDPC_decoder_irreg.py:102-105)."""
    H = sp.csr_matrix(H)
    np.savez(filename, data=H.data, indices=H.indices, indptr=H.indptr, shape=np.array(H.shape))
