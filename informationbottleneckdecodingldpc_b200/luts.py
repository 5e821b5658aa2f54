"""Look-up-table layout of the IB decoders and the decoder-config file format.

Layout facts, as produced by the reference's generator
(Discrete_LDPC_decoding/Discrete_Density_Evolution.py:92-95,120-122,299-344 and
Discrete_Density_Evolution_irreg.py:49-50,431-432) and as addressed by its kernels
(kernels_template_irreg.cl:60-96,125-177,205-245,277-300):

* check-node table  ``[iter0: Tc^2 | (DC-3) x Tc*T][iter g=1..imax-1: (DC-2) x T^2]``
* variable-node table ``imax x [Tc*T | (DV-1) x T^2]`` (last stage = decision stage)
* matching tables ``MC[imax][DC][T]``, ``MV[imax][DV][T]`` (second index = degree-1)

``DC``/``DV`` are the maximum degrees (the ``CN_DEGREE``/``VN_DEGREE`` macros).
"""
from __future__ import annotations

import pickle
from dataclasses import dataclass
from typing import Optional

import numpy as np


def cn_lut_len(Tc: int, T: int, dc: int, imax: int) -> int:
    return Tc * Tc + max(dc - 3, 0) * Tc * T + (imax - 1) * max(dc - 2, 0) * T * T


def vn_lut_len(Tc: int, T: int, dv: int, imax: int) -> int:
    return imax * (Tc * T + (dv - 1) * T * T)


def match_len(T: int, dmax: int, imax: int) -> int:
    return imax * dmax * T


@dataclass
class DecoderTables:
    """The six keys the BER drivers consume from ``decoder_config_*.pkl``
    (Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL.py:54-80)."""
    Trellis_checknodevector_a: np.ndarray
    Trellis_varnodevector_a: np.ndarray
    cardinality_T_decoder_ops: int
    imax: int
    matching_vector_checknode: Optional[np.ndarray] = None
    matching_vector_varnode: Optional[np.ndarray] = None

    def as_dict(self) -> dict:
        d = {
            "Trellis_checknodevector_a": self.Trellis_checknodevector_a,
            "Trellis_varnodevector_a": self.Trellis_varnodevector_a,
            "cardinality_T_decoder_ops": self.cardinality_T_decoder_ops,
            "imax": self.imax,
        }
        if self.matching_vector_checknode is not None:
            d["matching_vector_checknode"] = self.matching_vector_checknode
            d["matching_vector_varnode"] = self.matching_vector_varnode
        return d


def random_tables(T: int, dc: int, dv: int, imax: int, seed: int, Tc: Optional[int] = None,
                  matching: bool = False) -> DecoderTables:
    """Uniformly random tables in the reference layout: they have no decoding power but
    exercise every address the kernels can form, which is what bit-exactness needs.
    Stored as float arrays holding integers, like the generator's output
    (Discrete_Density_Evolution.py:301)."""
    Tc = T if Tc is None else Tc
    rng = np.random.Generator(np.random.PCG64(seed))
    cn = rng.integers(0, T, size=cn_lut_len(Tc, T, dc, imax)).astype(np.float64)
    vn = rng.integers(0, T, size=vn_lut_len(Tc, T, dv, imax)).astype(np.float64)
    mc = mv = None
    if matching:
        mc = rng.integers(0, T, size=match_len(T, dc, imax)).astype(np.float64)
        mv = rng.integers(0, T, size=match_len(T, dv, imax)).astype(np.float64)
    return DecoderTables(cn, vn, T, imax, mc, mv)


def save_config(tables: DecoderTables, filename: str, **extra) -> None:
    """Write a decoder-config file.  ``.pkl`` mirrors ``save_config`` of the reference
    (AWGN_Channel_Transmission/AWGN_Discrete_Density_Evolution.py:197-206: a pickled dict);
    ``.npz`` holds the same keys as plain arrays."""
    d = tables.as_dict()
    d.update(extra)
    if filename.endswith(".npz"):
        np.savez(filename, **{k: np.asarray(v) for k, v in d.items()})
    else:
        with open(filename, "wb") as fh:
            pickle.dump(d, fh, protocol=pickle.HIGHEST_PROTOCOL)


def load_config(filename: str) -> dict:
    """Read a decoder-config ``.pkl`` (reference format) or ``.npz``; returns the dict the
    drivers index (``generated_decoder['Trellis_checknodevector_a']`` ...)."""
    if filename.endswith(".npz"):
        z = np.load(filename, allow_pickle=False)
        d = {k: z[k] for k in z.files}
        for k in ("cardinality_T_decoder_ops", "imax"):
            if k in d:
                d[k] = int(d[k])
        return d
    with open(filename, "rb") as fh:
        return pickle.load(fh)


def as_int32(vec, name: str, T: int) -> np.ndarray:
    """``.astype(int)`` of the reference (discrete_LDPC_decoder.py:41-42) plus the range
    check the kernels rely on."""
    a = np.ascontiguousarray(np.asarray(vec).astype(np.int64).ravel())
    if a.size and (a.min() < 0 or a.max() >= T):
        raise ValueError(f"{name}: entries must lie in [0,{T}) (found {a.min()}..{a.max()})")
    return a.astype(np.int32)


def minsum_like_tables(T: int, dc: int, dv: int, imax: int) -> DecoderTables:
    """Deterministic tables with real decoding power, in the reference layout: cluster index t
    stands for the level ``t - (T-1)/2`` (t < T/2 <=> bit 1, as everywhere in the reference,
    e.g. kernels_template.cl:310); a check-node stage is sign*min on levels, a variable-node
    stage a saturating sum.  Not the authors' information-bottleneck tables (those need the
    absent ib_base package) -- used where a test or the benchmark needs frames
    that actually converge (early termination, BER sanity)."""
    half = (T - 1) / 2.0
    lev = np.arange(T) - half
    a, b = np.meshgrid(lev, lev, indexing="ij")          # a: running value t, b: next message
    cn = np.sign(a) * np.sign(b) * np.minimum(np.abs(a), np.abs(b))
    cn_idx = (cn + half).astype(np.int64)
    s = a + b                                             # integer-valued
    mag = np.minimum(np.abs(s), T // 2 - 1) + 0.5
    vn = np.where(s >= 0, mag, -mag)
    vn_idx = (vn + half).astype(np.int64)
    cn_stage = cn_idx.reshape(-1)                          # index t*T + m
    vn_stage = vn_idx.reshape(-1)
    cn_vec = np.concatenate([cn_stage] * (max(dc - 2, 1) + (imax - 1) * max(dc - 2, 0)))
    cn_vec = cn_vec[:cn_lut_len(T, T, dc, imax)]
    vn_vec = np.concatenate([vn_stage] * (imax * dv))
    ident_c = np.tile(np.arange(T), imax * dc)
    ident_v = np.tile(np.arange(T), imax * dv)
    return DecoderTables(cn_vec.astype(np.float64), vn_vec.astype(np.float64), T, imax,
                         ident_c.astype(np.float64), ident_v.astype(np.float64))
