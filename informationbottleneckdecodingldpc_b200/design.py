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
    ca = np.concatenate(([0.0], np.cumsum(pos[:, 0])))
    cb = np.concatenate(([0.0], np.cumsum(pos[:, 1])))
    # best[k][j]: best gain of splitting bins [0, j) into k clusters
    best = np.full((K + 1, half + 1), -np.inf)
    arg = np.zeros((K + 1, half + 1), dtype=np.int64)
    best[0, 0] = 0.0
    j_idx = np.arange(half + 1)
    for k in range(1, K + 1):
        for j in range(k, half + 1):
            i = j_idx[k - 1:j]                     # last cluster = bins [i, j)
            g = best[k - 1, i] + _pair_gain(ca[j] - ca[i], cb[j] - cb[i])
            m = int(np.argmax(g))
            best[k, j] = g[m]
            arg[k, j] = i[m]
    bounds = [half]
    j = half
    for k in range(K, 0, -1):
        j = int(arg[k, j])
        bounds.append(j)
    bounds = bounds[::-1]                          # 0 = b_0 < b_1 < ... < b_K = half
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
