#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the REFERENCE itself (run in the build container only).

TEST INFRASTRUCTURE ONLY.  Two sources of truth are frozen here, because the reference
ships no golden vectors of its own:

1. the reference's OpenCL kernels executed on the CPU (oracle/_ref, built by
   oracle/build_ref.py from the ``.cl`` files under /root/reference) -> decoder outputs and
   iteration counts for the cases listed in SURVEY.md 8c (i)-(vi);
2. the reference's importable Python host code (``map_node_connections``,
   ``set_code_parameters``, ``discrete_cn_operation`` / ``discrete_vn_operation``), imported
   with stub ``pyopencl`` / ``mako`` modules and ``np.int = int`` -> edge tables, R_c,
   data_len and per-node LUT results.

The committed .npz files let the CPU and GPU test-suites run where /root/reference does not
exist (the GPU box).  Re-run:  python oracle/make_golden.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT  # replace the script directory so that "oracle" is the package
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")

from informationbottleneckdecodingldpc_b200 import codes, graph, luts  # noqa: E402
from oracle import build_ref, oracle  # noqa: E402


def small_irregular(seed=7):
    """Synthetic irregular code containing degree-1 variable nodes; check degrees 3..5.
    (Degree-2 checks are left out of the *reference-generated* vectors because the reference
    kernel reads an uninitialised operand for them, kernels_template_irreg.cl:72.)"""
    dv = np.array([1, 1] + [2] * 20 + [3] * 16 + [4] * 10)       # 130 sockets
    dc = np.array([3] * 6 + [4] * 8 + [5] * 16)                  # 18 + 32 + 80 = 130
    rng = np.random.Generator(np.random.PCG64(seed))
    return codes.random_from_degrees(dv, rng.permutation(dc), seed)


def pack(H):
    H = H.tocsr()
    return dict(H_indptr=H.indptr.astype(np.int32), H_indices=H.indices.astype(np.int32),
                H_shape=np.array(H.shape, dtype=np.int64))


def ib_case(name, H, T, imax, B, seed, match, early, tables_kind="random", irregular=True, Tc=None):
    t = graph.edge_tables(H)
    DC, DV = t.d_c_max, t.d_v_max
    if tables_kind == "random":
        tb = luts.random_tables(T, DC, DV, imax, seed=seed, matching=match)
    else:
        tb = luts.minsum_like_tables(T, DC, DV, imax)
        if not match:
            tb.matching_vector_checknode = tb.matching_vector_varnode = None
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    if tables_kind == "random":
        ch = rng.integers(0, T, size=(t.n_var, B))
    else:  # mostly-correct channel so that the min-sum-like tables converge (all-zero codeword)
        ch = np.clip(np.round(rng.normal(T * 0.70, T * 0.15, size=(t.n_var, B))), 0, T - 1).astype(np.int64)
    kw = dict(T=T, imax=imax, cn_lut=tb.Trellis_checknodevector_a, vn_lut=tb.Trellis_varnodevector_a,
              cn_match=tb.matching_vector_checknode, vn_match=tb.matching_vector_varnode, early=early)
    out, i_num = oracle.ib_decode(t, ch, backend="reference", irregular=irregular, **kw)
    out_p, i_p = oracle.ib_decode(t, ch, backend="port", **kw)
    assert np.array_equal(out, out_p) and i_num == i_p, f"{name}: port != reference"
    d = pack(H)
    d.update(T=T, imax=imax, match=int(match), early=int(early), irregular=int(irregular),
             cn_lut=np.asarray(tb.Trellis_checknodevector_a, dtype=np.uint8),
             vn_lut=np.asarray(tb.Trellis_varnodevector_a, dtype=np.uint8),
             ch=ch.astype(np.uint8), out=out.astype(np.uint8), i_num=i_num)
    if match:
        d.update(cn_match=np.asarray(tb.matching_vector_checknode, dtype=np.uint8),
                 vn_match=np.asarray(tb.matching_vector_varnode, dtype=np.uint8))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **d)
    print(f"{name}: N={t.n_var} B={B} imax={imax} i_num={i_num} errors={(out < T // 2).sum()}")


def llr_case(name, H, imax, B, seed, early, levels=16):
    t = graph.edge_tables(H)
    rng = np.random.Generator(np.random.PCG64(seed))
    pos = np.sort(np.abs(rng.normal(0, 4, size=levels // 2)))
    llr_values = np.concatenate([-pos[::-1], pos])       # symmetric, increasing with the cluster index
    idx = np.clip(np.round(rng.normal(levels * 0.66, levels * 0.2, size=(t.n_var, B))), 0, levels - 1).astype(int)
    ch = llr_values[idx]
    d = pack(H)
    d.update(imax=imax, early=int(early), ch=ch)
    for algo in ("minsum", "bp"):
        out, i_num = oracle.llr_decode(t, ch, algo=algo, imax=imax, early=early, backend="reference")
        out_p, i_p = oracle.llr_decode(t, ch, algo=algo, imax=imax, early=early, backend="port")
        assert i_num == i_p
        if algo == "minsum":
            assert np.array_equal(out, out_p), f"{name}: min-sum port != reference"
        else:
            assert np.allclose(out, out_p, rtol=1e-12, atol=1e-12), f"{name}: BP port != reference"
        d[f"out_{algo}"] = out
        d[f"i_num_{algo}"] = i_num
        print(f"{name}/{algo}: i_num={i_num} neg={(out < 0).sum()}")
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **d)


def quantizer_case():
    T = 16
    limits = np.concatenate([[-3.0], np.sort(np.random.Generator(np.random.PCG64(3)).uniform(-2.5, 2.5, T - 1))])
    limits[T // 2] = 0.0
    limits = np.sort(limits)
    llr = np.linspace(-9, 9, T + 1)
    rng = np.random.Generator(np.random.PCG64(4))
    x = rng.normal(0, 1.5, size=(40, 7))
    # boundary cases: exactly on a limit, below limits[1], above limits[T-1]
    x[0, :] = limits[[1, 2, 5, 8, 9, 14, 15]]
    x[1, :] = [-10, -3, -2.9999999, 10, 2.5, 0.0, -0.0]
    x[2, :] = np.nextafter(limits[[1, 2, 5, 8, 9, 14, 15]], np.inf)
    cl = oracle.quantize(x, limits, T, backend="reference")
    assert np.array_equal(cl, oracle.quantize(x, limits, T, backend="port"))
    cdf = np.concatenate([[0.0], np.cumsum(np.full(T, 1.0 / T))])
    u = rng.random(size=(40, 7))
    u[0, :3] = [0.0, cdf[3], cdf[T - 1]]
    cl_direct = oracle.quantize(u, cdf, T + 1, backend="reference")
    assert np.array_equal(cl_direct, oracle.quantize(u, cdf, T + 1, backend="port"))
    ll = oracle.quantize_llr(u, cdf, T + 1, llr, backend="reference")
    assert np.array_equal(ll, oracle.quantize_llr(u, cdf, T + 1, llr, backend="port"))
    np.savez_compressed(os.path.join(GOLD, "quantizer.npz"), T=T, limits=limits, x=x, clusters=cl,
                        cdf=cdf, u=u, clusters_direct=cl_direct, llr_values=llr, llrs_direct=ll)
    print("quantizer: ok", cl.min(), cl.max(), cl_direct.min(), cl_direct.max())


def reference_host_tables():
    """Import the reference's Python decoder classes (host side only) and freeze what they compute."""
    for mod in ("pyopencl", "pyopencl.array", "pyopencl.reduction", "pyopencl.tools", "pyopencl.clrandom",
                "mako", "mako.template"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    sys.modules["pyopencl.reduction"].get_sum_kernel = lambda *a, **k: None
    sys.modules["mako.template"].Template = object
    sys.modules["pyopencl"].array = sys.modules["pyopencl.array"]
    np.int = int  # removed from numpy >= 1.24; every reference module uses it
    sys.path.insert(0, REF)
    from Discrete_LDPC_decoding.discrete_LDPC_decoder import Discrete_LDPC_Decoder_class
    from Discrete_LDPC_decoding.discrete_LDPC_decoder_irreg import Discrete_LDPC_Decoder_class_irregular
    import tempfile
    tmp = tempfile.mkdtemp()
    out = {}
    # regular class, alist input (dense map_node_connections, discrete_LDPC_decoder.py:88-130)
    Hr = codes.regular_random(96, 3, 6, seed=11)
    alist = os.path.join(tmp, "reg.alist")
    graph.write_alist(Hr, alist)
    tb = luts.random_tables(16, 6, 3, 4, seed=5)
    dec = Discrete_LDPC_Decoder_class(alist, 4, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, 2)
    out.update(reg_H_indptr=Hr.indptr, reg_H_indices=Hr.indices, reg_shape=np.array(Hr.shape),
               reg_sc=dec.inbox_memory_start_checknodes, reg_sv=dec.inbox_memory_start_varnodes,
               reg_tc=dec.target_memory_cells_checknodes, reg_tv=dec.target_memory_cells_varnodes)
    # node-op formulas (discrete_LDPC_decoder.py:302-355) on random inputs
    rng = np.random.Generator(np.random.PCG64(9))
    yc = rng.integers(0, 16, size=(50, 5))
    yv = rng.integers(0, 16, size=(50, 3))
    out.update(reg_cn_lut=tb.Trellis_checknodevector_a, reg_vn_lut=tb.Trellis_varnodevector_a, yc=yc, yv=yv,
               cn_op_iter0=dec.discrete_cn_operation(yc, 0), cn_op_iter2=dec.discrete_cn_operation(yc, 2),
               vn_op_iter1=dec.discrete_vn_operation(yv, 1))
    # irregular class on the WLAN matrix (.npy input, sparse map_node_connections, _irreg.py:121-170)
    Hw = codes.wlan_80211n(54)
    npy = os.path.join(tmp, "WLAN_H.npy")
    np.save(npy, Hw.toarray().astype(np.float64))
    tbw = luts.random_tables(16, 8, 11, 3, seed=6, matching=True)
    decw = Discrete_LDPC_Decoder_class_irregular(npy, 3, 16, 16, tbw.Trellis_checknodevector_a,
                                                 tbw.Trellis_varnodevector_a, tbw.matching_vector_checknode,
                                                 tbw.matching_vector_varnode, 2)
    out.update(wlan_sc=decw.inbox_memory_start_checknodes, wlan_sv=decw.inbox_memory_start_varnodes,
               wlan_tc=decw.target_memory_cells_checknodes, wlan_tv=decw.target_memory_cells_varnodes,
               wlan_R_c=float(decw.R_c), wlan_data_len=int(decw.data_len), wlan_d_c_max=int(decw.d_c_max),
               wlan_d_v_max=int(decw.d_v_max))
    # DVB-S2-like (.npz CSR input) -- R_c / data_len arithmetic
    Hd = codes.dvbs2_like_half_rate(6480, q_groups=36)
    npz = os.path.join(tmp, "dvb.npz")
    codes.save_csr_npz(Hd, npz)
    tbd = luts.random_tables(16, 7, 8, 2, seed=8, matching=True)
    decd = Discrete_LDPC_Decoder_class_irregular(npz, 2, 16, 16, tbd.Trellis_checknodevector_a,
                                                 tbd.Trellis_varnodevector_a, tbd.matching_vector_checknode,
                                                 tbd.matching_vector_varnode, 2)
    td = graph.edge_tables(Hd)
    assert np.array_equal(decd.target_memory_cells_checknodes, td.target_cells_chk)
    assert np.array_equal(decd.target_memory_cells_varnodes, td.target_cells_var)
    out.update(dvb_small_R_c=float(decd.R_c), dvb_small_data_len=int(decd.data_len))
    np.savez_compressed(os.path.join(GOLD, "reference_host_tables.npz"), **out)
    print("reference host tables: WLAN R_c", decw.R_c, "data_len", decw.data_len,
          "| DVB-like(6480) R_c", decd.R_c, "data_len", decd.data_len)


def main():
    if not os.path.isdir(REF):
        sys.exit("the reference tree is needed to (re)generate golden vectors")
    os.makedirs(GOLD, exist_ok=True)
    build_ref.build(REF, quiet=True)
    reference_host_tables()
    # (i) toy regular codes
    ib_case("ib_toy_3_6_n24", codes.regular_random(24, 3, 6, seed=1), 16, 5, 4, 100, False, True, irregular=False)
    ib_case("ib_toy_2_4_n32_T8", codes.regular_random(32, 2, 4, seed=2), 8, 6, 5, 101, False, False, irregular=False)
    ib_case("ib_toy_3_4_n36", codes.regular_random(36, 3, 4, seed=3), 16, 4, 19, 102, False, True, irregular=False)
    ib_case("ib_toy_3_5_n40_irrkernels", codes.regular_random(40, 3, 5, seed=4), 16, 7, 33, 103, False, True)
    # (ii) C1: (3,6) n=8000, random tables, B=32, imax 5/50, ET on/off; converging tables with ET
    Hc1 = codes.regular_random(8000, 3, 6)
    ib_case("ib_c1_rand_imax5_et", Hc1, 16, 5, 32, 200, False, True, irregular=False)
    ib_case("ib_c1_rand_imax50_fixed", Hc1, 16, 50, 32, 201, False, False, irregular=False)
    ib_case("ib_c1_minsumlut_imax50_et", Hc1, 16, 50, 32, 202, False, True, tables_kind="minsum", irregular=False)
    ib_case("ib_c1_minsumlut_imax50_fixed", Hc1, 16, 50, 24, 203, False, False, tables_kind="minsum", irregular=False)
    # (iii) WLAN n=1296: T=16/32, match true/false, ragged batch sizes
    Hw = codes.wlan_80211n(54)
    ib_case("ib_wlan_T16_match", Hw, 16, 8, 5, 300, True, True)
    ib_case("ib_wlan_T16_nomatch", Hw, 16, 8, 17, 301, False, False)
    ib_case("ib_wlan_T32_match", Hw, 32, 4, 3, 302, True, True)
    ib_case("ib_wlan_T16_minsumlut_et", Hw, 16, 30, 48, 303, True, True, tables_kind="minsum")
    # (iv) corner cases: degree-1 VN; DVB-S2-like scaled instance (degree-1 VN, d_v 8, d_c 6/7)
    ib_case("ib_irreg_deg1vn", small_irregular(), 16, 6, 9, 400, True, True)
    ib_case("ib_dvb_small_match", codes.dvbs2_like_half_rate(6480, q_groups=36), 16, 6, 20, 401, True, False)
    # (v) min-sum / BP float64
    llr_case("llr_toy_3_6_n24", codes.regular_random(24, 3, 6, seed=1), 6, 5, 500, True)
    llr_case("llr_wlan", Hw, 12, 6, 501, True)
    llr_case("llr_c1", Hc1, 10, 8, 502, False)
    # (vi) quantizer
    quantizer_case()


if __name__ == "__main__":
    main()
