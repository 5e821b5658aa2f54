"""Tanner-graph tables for the flooding decoders (host side, numpy/scipy only).

What the reference builds in ``map_node_connections``
(Discrete_LDPC_decoding/discrete_LDPC_decoder.py:88-130 dense,
Discrete_LDPC_decoding/discrete_LDPC_decoder_irreg.py:121-170 sparse) is
restated here once, from the CSR structure of H, without ever forming a dense
matrix:

* check-node inbox (VN->CN messages) is CN-major: slots ``sc[c] .. sc[c]+deg(c)-1``,
  slot k <-> k-th neighbour variable of c in ascending variable index;
* variable-node inbox (CN->VN messages) is VN-major: slots ``sv[v] .. sv[v]+deg(v)-1``,
  slot k <-> k-th neighbour check of v in ascending check index;
* ``tc[e_c]`` = VN-inbox slot of CN-major edge ``e_c``; ``tv[e_v]`` = CN-inbox slot of
  VN-major edge ``e_v``.  They are inverse permutations of each other.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp


def alist_to_csr(lines) -> sp.csr_matrix:
    """Parse AList lines (already split to ints) into a CSR 0/1 matrix.

    Same accepted dialects as the reference parser ``alistToNumpy``
    (discrete_LDPC_decoder.py:57-81): full format (lines 3/4 hold the column and
    row weights) or the reduced one without them; zero entries are padding.
    """
    n_cols, n_rows = int(lines[0][0]), int(lines[0][1])
    if len(lines) > 3 and len(lines[2]) == n_cols and len(lines[3]) == n_rows:
        start = 4
    else:
        start = 2
    rows, cols = [], []
    for col, nonzeros in enumerate(lines[start:start + n_cols]):
        for r in nonzeros:
            if r != 0:
                rows.append(int(r) - 1)
                cols.append(col)
    data = np.ones(len(rows), dtype=np.int8)
    H = sp.csr_matrix((data, (rows, cols)), shape=(n_rows, n_cols))
    H.sum_duplicates()
    H.data[:] = 1
    return H


def load_check_matrix(filename: str) -> sp.csr_matrix:
    """Load H from ``.npy`` (dense), ``.npz`` (CSR parts) or an AList text file.

    File conventions follow discrete_LDPC_decoder_irreg.py:102-119.
    """
    if filename.endswith(".npy"):
        H = sp.csr_matrix(np.load(filename) != 0, dtype=np.int8)
    elif filename.endswith(".npz"):
        z = np.load(filename)
        H = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
        H = sp.csr_matrix(H != 0, dtype=np.int8)
    else:
        with open(filename) as fh:
            lines = [[int(tok) for tok in line.split()] for line in fh]
        lines = [ln for ln in lines]
        H = alist_to_csr(lines)
    H = H.tocsr()
    H.sort_indices()
    return H


def write_alist(H: sp.spmatrix, filename: str) -> None:
    """Write H in the full AList format (MacKay), the input the regular decoder reads."""
    H = sp.csr_matrix(H)
    Hc = H.tocsc()
    Hc.sort_indices()
    H.sort_indices()
    n_rows, n_cols = H.shape
    dv = np.diff(Hc.indptr)
    dc = np.diff(H.indptr)
    with open(filename, "w") as fh:
        fh.write(f"{n_cols} {n_rows}\n{dv.max()} {dc.max()}\n")
        fh.write(" ".join(map(str, dv)) + "\n")
        fh.write(" ".join(map(str, dc)) + "\n")
        for v in range(n_cols):
            r = Hc.indices[Hc.indptr[v]:Hc.indptr[v + 1]] + 1
            r = list(r) + [0] * (dv.max() - len(r))
            fh.write(" ".join(map(str, r)) + "\n")
        for c in range(n_rows):
            r = H.indices[H.indptr[c]:H.indptr[c + 1]] + 1
            r = list(r) + [0] * (dc.max() - len(r))
            fh.write(" ".join(map(str, r)) + "\n")


@dataclass
class EdgeTables:
    """The six int32 tables the reference uploads (discrete_LDPC_decoder_irreg.py:191-205)."""
    n_var: int
    n_chk: int
    n_edge: int
    degree_chk: np.ndarray        # [M]
    degree_var: np.ndarray        # [N]
    inbox_start_chk: np.ndarray   # [M]  (sc)
    inbox_start_var: np.ndarray   # [N]  (sv)
    target_cells_chk: np.ndarray  # [E]  (tc) CN-major edge -> VN-inbox slot
    target_cells_var: np.ndarray  # [E]  (tv) VN-major edge -> CN-inbox slot
    var_of_chk_slot: np.ndarray   # [E]  variable index of every CN-major slot (H.indices)
    chk_of_var_slot: np.ndarray   # [E]  check index of every VN-major slot

    @property
    def d_c_max(self) -> int:
        return int(self.degree_chk.max())

    @property
    def d_v_max(self) -> int:
        return int(self.degree_var.max())


def edge_tables(H: sp.spmatrix) -> EdgeTables:
    H = sp.csr_matrix(H)
    H.sort_indices()
    M, N = H.shape
    E = int(H.nnz)
    deg_c = np.diff(H.indptr).astype(np.int32)
    deg_v = np.bincount(H.indices, minlength=N).astype(np.int32)
    sc = np.concatenate(([0], np.cumsum(deg_c[:-1]))).astype(np.int32)
    sv = np.concatenate(([0], np.cumsum(deg_v[:-1]))).astype(np.int32)
    var_of = H.indices.astype(np.int32)
    # CSC position of every CSR edge: stable sort by variable keeps ascending check order.
    tv = np.argsort(var_of, kind="stable").astype(np.int32)   # VN-major slot -> CN-major slot
    tc = np.empty(E, dtype=np.int32)
    tc[tv] = np.arange(E, dtype=np.int32)                     # CN-major slot -> VN-major slot
    chk_of_cslot = np.repeat(np.arange(M, dtype=np.int32), deg_c)
    chk_of = chk_of_cslot[tv]
    return EdgeTables(N, M, E, deg_c, deg_v, sc, sv, tc, tv, var_of, chk_of)


def code_rate_from_degrees(H: sp.spmatrix) -> float:
    """Design rate 1 - mean(d_v)/mean(d_c) exactly as ``set_code_parameters`` computes it
    (discrete_LDPC_decoder_irreg.py:69-100)."""
    H = sp.csr_matrix(H)
    dv = np.asarray(H.sum(0)).ravel().astype(np.int64)
    dc = np.asarray(H.sum(1)).ravel().astype(np.int64)

    def mean_degree(deg):
        # node-perspective degree histogram, normalised, then dotted with 1..dmax:
        # the same floating-point operation order as the reference, so that
        # data_len = int(R_c * N) truncates identically.
        hist = np.bincount(deg, minlength=int(deg.max()) + 1)[1:].astype(np.float64)
        hist = hist / hist.sum()
        return np.dot(hist, np.arange(int(deg.max())) + 1)

    return 1 - mean_degree(dv) / mean_degree(dc)
