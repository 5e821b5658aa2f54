"""Deterministic channel-quantizer design: symmetric maximum-mutual-information quantizer.

The reference designs its quantizer with the sequential information-bottleneck algorithm of the
un-vendored ``information_bottleneck`` (ib_base) package (AWGN_Quantizer_BPSK.py:2,81-85).  That
package is not available and no reference artefact pins its output (random restarts, ``nror``),
so this is a replacement with the properties the callers rely on:
clusters contiguous in LLR order, indexed by increasing LLR, mirror-symmetric, deterministic
one-hot ``p(t|y)``.  Threshold design parity with the authors' tables: UNPINNED.

For a binary-input symmetric channel the mutual-information-optimal quantizer has contiguous
decision regions in the LLR, so a dynamic programme over the sorted outputs is exact; the
mirror symmetry is imposed by designing the positive half and reflecting it.
"""
from __future__ import annotations

import numpy as np


def _pair_gain(a, b):
    """Contribution of a cluster with joint masses (a, b) and of its mirror (b, a) to I(X;T), nats."""
    s = a + b
    with np.errstate(divide="ignore", invalid="ignore"):
        ga = np.where(a > 0, a * np.log(a / (0.5 * s)), 0.0)
        gb = np.where(b > 0, b * np.log(b / (0.5 * s)), 0.0)
    return 2.0 * (ga + gb)


def _dp_bounds(a_mass: np.ndarray, b_mass: np.ndarray, K: int):
    """Optimal split of `n` LLR-sorted symbols (joint masses a = p(x=0,y), b = p(x=1,y)) into K
    contiguous clusters maximising the sum of `_pair_gain`.  Returns the K+1 boundaries
    0 = b_0 < ... < b_K = n.  Vectorised over (i, j): O(K n^2) numpy work."""
    n = a_mass.shape[0]
    ca = np.concatenate(([0.0], np.cumsum(a_mass)))
    cb = np.concatenate(([0.0], np.cumsum(b_mass)))
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    gain = np.where(ii < jj, _pair_gain(np.maximum(ca[None, :] - ca[:, None], 0.0),
                                        np.maximum(cb[None, :] - cb[:, None], 0.0)), -np.inf)
    best = np.full(n + 1, -np.inf)
    best[0] = 0.0
    args = []
    for _k in range(K):
        cand = best[:, None] + gain                 # cand[i, j]: last cluster = symbols [i, j)
        arg = np.argmax(cand, axis=0)
        best = cand[arg, np.arange(n + 1)]
        args.append(arg)
    bounds = [n]
    j = n
    for k in range(K - 1, -1, -1):
        j = int(args[k][j])
        bounds.append(j)
    return bounds[::-1]


def symmetric_mi_quantizer(p_xy: np.ndarray, cardinality_T: int):
    """p_xy: (Y, 2) joint pmf of the fine channel output y (increasing LLR order, mirror
    symmetric: p(y_i, x=0) = p(y_{Y-1-i}, x=1)) and the bit x.  Returns
    (p_t_given_y one-hot (Y,T), p_x_given_t (T,2), p_t (T,)), clusters ordered by increasing LLR."""
    Y = p_xy.shape[0]
    T = int(cardinality_T)
    if Y % 2 or T % 2:
        raise ValueError("cardinality_Y and cardinality_T must be even")
    half, K = Y // 2, T // 2
    pos = p_xy[half:, :]                          # y > 0 side, bins 0..half-1
    bounds = _dp_bounds(pos[:, 0], pos[:, 1], K)   # 0 = b_0 < b_1 < ... < b_K = half
    cluster = np.empty(Y, dtype=np.int64)
    for k in range(K):
        cluster[half + bounds[k]: half + bounds[k + 1]] = K + k
    cluster[:half] = T - 1 - cluster[half:][::-1]
    p_t_given_y = np.zeros((Y, T))
    p_t_given_y[np.arange(Y), cluster] = 1.0
    p_xt = p_t_given_y.T @ p_xy                   # (T, 2)
    p_t = p_xt.sum(1)
    p_x_given_t = p_xt / p_t[:, None]
    return p_t_given_y, p_x_given_t, p_t


# ---------------------------------------------------------------------------------------------
# Discrete density evolution for regular codes: designs the look-up tables of the IB decoder.
#
# Replaces the design chain of the reference (Discrete_LDPC_decoding/Discrete_Density_Evolution.py
# + ib_base's lin_sym_sIB, driven by Regular_LDPC_Decoding/BPSK/decoder_config_generation.py) with
# the deterministic max-MI quantizer above.  The table LAYOUT is the reference's
# (Discrete_Density_Evolution.py:299-344); the table CONTENTS differ from
# the authors' sequential-IB tables (random restarts, no artefact to pin against): parity unpinned.
# ---------------------------------------------------------------------------------------------
_P_MIN = 1e-15


def _guard(p):
    """numerical guard of the reference (Discrete_Density_Evolution.py:434-440)."""
    p = np.clip(p, _P_MIN, 0.5 - _P_MIN)
    return p / p.sum()


def quantize_joint(p_xy: np.ndarray, T: int):
    """Mutual-information-maximising, mirror-symmetric quantizer of an arbitrary-order joint pmf
    p_xy (n, 2) [columns x=0, x=1].  Returns (labels (n,), p_xt (T, 2)): cluster index per row,
    clusters contiguous in LLR, numbered by increasing LLR log p(x=0|t)/p(x=1|t)."""
    n = p_xy.shape[0]
    if n % 2 or T % 2:
        raise ValueError("even sizes expected")
    llr = np.log(p_xy[:, 0]) - np.log(p_xy[:, 1])
    order = np.argsort(llr, kind="stable")
    half, K = n // 2, T // 2
    pos = p_xy[order[half:], :]                       # upper half in increasing LLR
    if half <= K:                                     # fewer symbols than clusters: one each
        bounds = list(range(half + 1)) + [half] * (K - half)
    else:
        bounds = _dp_bounds(pos[:, 0], pos[:, 1], K)
    lab_sorted = np.empty(n, dtype=np.int64)
    for k in range(K):
        lab_sorted[half + bounds[k]: half + bounds[k + 1]] = K + k
    lab_sorted[:half] = T - 1 - lab_sorted[half:][::-1]
    labels = np.empty(n, dtype=np.int64)
    labels[order] = lab_sorted
    p_xt = np.zeros((T, 2))
    np.add.at(p_xt, labels, p_xy)
    return labels, p_xt


def _cn_joint(p_t, p_y):
    """p(x, [t, y]) of a partial check-node operation, x = b_t xor b_y; row index t*|Y| + y
    (Discrete_Density_Evolution.py:346-386)."""
    same = p_t[:, None, 0] * p_y[None, :, 0] + p_t[:, None, 1] * p_y[None, :, 1]
    diff = p_t[:, None, 0] * p_y[None, :, 1] + p_t[:, None, 1] * p_y[None, :, 0]
    return np.stack([same.reshape(-1), diff.reshape(-1)], axis=1)


def _vn_joint(p_t, p_y):
    """p(x, [t, y]) of a partial variable-node operation (both observe the same bit); row index
    t*|Y| + y (Discrete_Density_Evolution.py:390-430)."""
    j = 2.0 * p_t[:, None, :] * p_y[None, :, :]
    return j.reshape(-1, 2)


def _mi(p_xt):
    p = p_xt / p_xt.sum()
    px, pt = p.sum(0), p.sum(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.nansum(p * np.log2(p / (pt[:, None] * px[None, :]))))


def design_regular_ib_decoder(p_x_and_t_channel: np.ndarray, d_v: int, d_c: int, T: int, imax: int):
    """Discrete density evolution for a (d_v, d_c)-regular code.

    p_x_and_t_channel: (Tc, 2) joint pmf of the channel cluster and the bit, e.g.
    ``AWGN_Channel_Quantizer.p_x_and_t`` (Tc must equal T, as in every script of the reference).
    Returns (Trellis_checknodevector_a, Trellis_varnodevector_a, mi_vn_out[imax]) in the reference
    layout: CN ``[Tc^2 | (d_c-3) x Tc*T | (imax-1) x (d_c-2) x T^2]``, VN ``imax x [Tc*T | (d_v-1) x T^2]``."""
    p_ch = np.asarray(p_x_and_t_channel, dtype=np.float64)
    p_ch = p_ch / p_ch.sum()
    Tc = p_ch.shape[0]
    if Tc != T:
        raise ValueError("cardinality_T_channel must equal cardinality_T_decoder_ops")
    cn_vec, vn_vec, mi_hist = [], [], []
    vn_msg = p_ch
    for _it in range(imax):
        p_t = vn_msg
        for _l in range(d_c - 2):
            lab, p_t = quantize_joint(_guard(_cn_joint(p_t, vn_msg)), T)
            cn_vec.append(lab)
        cn_out = p_t / p_t.sum()
        p_t = p_ch
        nxt = None
        for l in range(d_v):
            lab, p_t = quantize_joint(_guard(_vn_joint(p_t, cn_out)), T)
            p_t = p_t / p_t.sum()
            vn_vec.append(lab)
            if l == d_v - 2:
                nxt = p_t                          # extrinsic message: channel + d_v-1 check messages
        vn_msg = nxt if nxt is not None else p_t
        mi_hist.append(_mi(vn_msg))
    cn = np.concatenate(cn_vec).astype(np.float64) if cn_vec else np.zeros(0)
    vn = np.concatenate(vn_vec).astype(np.float64)
    return cn, vn, np.array(mi_hist)


# ---------------------------------------------------------------------------------------------
# Irregular codes: degree-mixed density evolution with message alignment.
#
# Follows the structure of the reference's irregular design chain
# (Discrete_LDPC_decoding/Discrete_Density_Evolution_irreg.py:95-135,225-300 and
# Information_Matching.py:34-78): one chain of partial node operations is designed for the
# maximum degree, a node of degree d taps the chain after its own last stage, every tapped
# message is aligned to a reference meaning by z*(t) = argmin_z KL(p(x|t) || p_ref(x|z)), and the
# aligned densities are mixed with the edge-perspective degree distributions.  Deterministic;
# contents not pinned to the authors' tables (no artefact, ib_base absent).
# ---------------------------------------------------------------------------------------------
def edge_degree_distribution(node_degrees: np.ndarray) -> np.ndarray:
    """Edge-perspective distribution: entry d-1 = fraction of edges attached to degree-d nodes
    (Information_Matching.py:23-31)."""
    node_degrees = np.asarray(node_degrees, dtype=np.int64)
    hist = np.bincount(node_degrees, minlength=int(node_degrees.max()) + 1)[1:].astype(np.float64)
    w = hist * (np.arange(hist.size) + 1)
    return w / w.sum()


def _kl_rows(p_row, q_rows):
    p = np.clip(p_row, 1e-300, None)
    q = np.clip(q_rows, 1e-300, None)
    return np.sum(p[None, :] * (np.log(p[None, :]) - np.log(q)), axis=1)


def align_messages(p_xt: np.ndarray, p_ref: np.ndarray):
    """Message alignment (information matching): z*(t) = argmin_z KL(p(x|t) || p_ref(x|z)).
    Returns (z_star (T,), aligned joint pmf (T, 2))."""
    T = p_xt.shape[0]
    cond = p_xt / np.clip(p_xt.sum(1, keepdims=True), 1e-300, None)
    ref = p_ref / np.clip(p_ref.sum(1, keepdims=True), 1e-300, None)
    z = np.array([int(np.argmin(_kl_rows(cond[t], ref))) for t in range(T)], dtype=np.int64)
    out = np.zeros_like(p_xt)
    np.add.at(out, z, p_xt)
    return z, out


def _mean_abs_llr(p_xt):
    p = np.clip(p_xt, 1e-300, None)
    return float(np.sum(p.sum(1) * np.abs(np.log(p[:, 0]) - np.log(p[:, 1]))))


def design_irregular_ib_decoder(p_x_and_t_channel, lambda_edge, rho_edge, T: int, imax: int):
    """Design tables for an irregular ensemble.  lambda_edge / rho_edge: edge-perspective degree
    distributions (index d-1).  Returns (cn_vec, vn_vec, mc_vec, mv_vec, mi[imax]) in the reference
    layout with DC = len(rho_edge), DV = len(lambda_edge)."""
    p_ch = np.asarray(p_x_and_t_channel, dtype=np.float64)
    p_ch = p_ch / p_ch.sum()
    if p_ch.shape[0] != T:
        raise ValueError("cardinality_T_channel must equal cardinality_T_decoder_ops")
    lam = np.asarray(lambda_edge, dtype=np.float64)
    rho = np.asarray(rho_edge, dtype=np.float64)
    DV, DC = lam.size, rho.size
    ident = np.arange(T)
    cn_vec, vn_vec, mi_hist = [], [], []
    MC = np.tile(ident, (imax, DC, 1))
    MV = np.tile(ident, (imax, DV, 1))
    p_in = p_ch
    for g in range(imax):
        # ---- check nodes: chain of DC-2 stages, degree d taps after stage d-3
        taps = {2: p_in}
        p_t = p_in
        for l in range(DC - 2):
            lab, p_t = quantize_joint(_guard(_cn_joint(p_t, p_in)), T)
            p_t = p_t / p_t.sum()
            cn_vec.append(lab)
            taps[l + 3] = p_t
        active = [d for d in range(2, DC + 1) if rho[d - 1] > 0]
        ref_d = max(active, key=lambda d: _mean_abs_llr(taps[d]))       # most reliable = lowest degree
        p_c = np.zeros((T, 2))
        for d in active:
            if d == ref_d:
                q = taps[d]
            else:
                z, q = align_messages(taps[d], taps[ref_d])
                MC[g, d - 1, :] = z
            p_c += rho[d - 1] * q
        p_c = p_c / p_c.sum()
        # ---- variable nodes: chain of DV stages (last = decision stage), degree d taps after stage d-2
        vtaps = {1: p_ch}
        p_t = p_ch
        for l in range(DV):
            lab, p_t = quantize_joint(_guard(_vn_joint(p_t, p_c)), T)
            p_t = p_t / p_t.sum()
            vn_vec.append(lab)
            vtaps[l + 2] = p_t
        vactive = [d for d in range(1, DV + 1) if lam[d - 1] > 0]
        cand = [d for d in vactive if d >= 2]
        ref_d = max(cand, key=lambda d: _mean_abs_llr(vtaps[d])) if cand else 1
        p_v = np.zeros((T, 2))
        for d in vactive:
            if d == 1 or d == ref_d:          # degree-1 nodes forward the raw channel value unaligned
                q = vtaps[d]
            else:
                z, q = align_messages(vtaps[d], vtaps[ref_d])
                MV[g, d - 1, :] = z
            p_v += lam[d - 1] * q
        p_in = p_v / p_v.sum()
        mi_hist.append(_mi(p_in))
    cn = np.concatenate(cn_vec).astype(np.float64) if cn_vec else np.zeros(0)
    vn = np.concatenate(vn_vec).astype(np.float64)
    return cn, vn, MC.reshape(-1).astype(np.float64), MV.reshape(-1).astype(np.float64), np.array(mi_hist)
