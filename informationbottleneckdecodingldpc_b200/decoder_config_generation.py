"""Generate IB decoder-config files (regular codes) -- stand-in for the reference's
``Regular_LDPC_Decoding/BPSK/decoder_config_generation.py`` (:16-41), which runs discrete density
evolution through the absent ib_base package and pickles the result as
``decoder_config_EbN0_gen_<EbN0>_<T>.pkl``.  Same file name, same keys, same table layout; the
tables come from the deterministic design in design.py (contents not pinned to the authors' tables).

    python -m informationbottleneckdecodingldpc_b200.decoder_config_generation --ebn0 1.2 --dv 3 --dc 6
"""
from __future__ import annotations

import argparse

import numpy as np

from .AWGN_Channel_Transmission.AWGN_Quantizer_BPSK import AWGN_Channel_Quantizer
from .design import design_irregular_ib_decoder, design_regular_ib_decoder, edge_degree_distribution
from .luts import DecoderTables, save_config


def generate_regular_config(EbN0_dB: float, d_v: int = 3, d_c: int = 6, cardinality_T: int = 16, imax: int = 50,
                            AD_max_abs: float = 3.0, cardinality_Y_channel: int = 2000):
    """Design quantizer + decoder tables for one Eb/N0 (sigma_n2 = 10^(-EbN0/10) / (2 R_c), as in
    Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py:85).  Returns (DecoderTables, extras dict)."""
    R_c = 1.0 - d_v / d_c
    sigma_n2 = 10 ** (-EbN0_dB / 10) / (2 * R_c)
    quanti = AWGN_Channel_Quantizer(sigma_n2, AD_max_abs, cardinality_T, cardinality_Y_channel)
    cn, vn, mi = design_regular_ib_decoder(quanti.p_x_and_t, d_v, d_c, cardinality_T, imax)
    tables = DecoderTables(cn, vn, cardinality_T, imax)
    extras = dict(sigma_n2=sigma_n2, EbN0=EbN0_dB, d_v=d_v, d_c=d_c, AD_max_abs=AD_max_abs,
                  cardinality_Y_channel=cardinality_Y_channel, cardinality_T_channel=cardinality_T,
                  p_x_and_t_input=quanti.p_x_and_t, ext_mi_varnode_in_iter=np.asarray(mi))
    return tables, extras


def generate_irregular_config(EbN0_dB: float, H, cardinality_T: int = 16, imax: int = 50, AD_max_abs: float = 3.0,
                              cardinality_Y_channel: int = 2000):
    """Irregular codes (stand-in for Irregular_LDPC_Decoding/{WLAN,DVB-S2}/decoder_config_generation.py):
    degree distributions are read off the parity-check matrix H, tables and matching vectors come from
    design.design_irregular_ib_decoder.  Returns (DecoderTables incl. matching vectors, extras)."""
    from .graph import code_rate_from_degrees, edge_tables
    t = edge_tables(H)
    R_c = float(code_rate_from_degrees(H))
    sigma_n2 = 10 ** (-EbN0_dB / 10) / (2 * R_c)
    quanti = AWGN_Channel_Quantizer(sigma_n2, AD_max_abs, cardinality_T, cardinality_Y_channel)
    lam = edge_degree_distribution(t.degree_var)
    rho = edge_degree_distribution(t.degree_chk)
    cn, vn, mc, mv, mi = design_irregular_ib_decoder(quanti.p_x_and_t, lam, rho, cardinality_T, imax)
    tables = DecoderTables(cn, vn, cardinality_T, imax, mc, mv)
    extras = dict(sigma_n2=sigma_n2, EbN0=EbN0_dB, lambda_vec=lam, rho_vec=rho, R_c=R_c, match='true',
                  cardinality_T_channel=cardinality_T, p_x_and_t_input=quanti.p_x_and_t,
                  ext_mi_varnode_in_iter=np.asarray(mi))
    return tables, extras


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--ebn0", type=float, default=1.2)
    ap.add_argument("--dv", type=int, default=3)
    ap.add_argument("--dc", type=int, default=6)
    ap.add_argument("--T", type=int, default=16)
    ap.add_argument("--imax", type=int, default=50)
    ap.add_argument("--out", default=None)
    a = ap.parse_args(argv)
    tables, extras = generate_regular_config(a.ebn0, a.dv, a.dc, a.T, a.imax)
    out = a.out or f"decoder_config_EbN0_gen_{a.ebn0}_{a.T}.pkl"
    save_config(tables, out, **extras)
    print(f"wrote {out}: CN table {tables.Trellis_checknodevector_a.size}, VN table {tables.Trellis_varnodevector_a.size} entries; "
          f"I(X;T) after {a.imax} iterations = {extras['ext_mi_varnode_in_iter'][-1]:.4f} bit")


if __name__ == "__main__":
    main()
