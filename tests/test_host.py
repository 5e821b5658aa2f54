"""CPU suite: host-side logic (graph tables, generators, table layout, config files, quantizer design)
and the C-ABI library surface.  No GPU compute here."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from informationbottleneckdecodingldpc_b200 import codes, graph, luts


def test_edge_tables_match_reference_classes():
    """Tables frozen from the reference's map_node_connections (dense and sparse variants)."""
    g = load_golden("reference_host_tables")
    import scipy.sparse as sp
    Hr = sp.csr_matrix((np.ones(g["reg_H_indices"].size), g["reg_H_indices"], g["reg_H_indptr"]), shape=tuple(g["reg_shape"]))
    t = graph.edge_tables(Hr)
    assert np.array_equal(t.inbox_start_chk, g["reg_sc"]) and np.array_equal(t.inbox_start_var, g["reg_sv"])
    assert np.array_equal(t.target_cells_chk, g["reg_tc"]) and np.array_equal(t.target_cells_var, g["reg_tv"])
    tw = graph.edge_tables(codes.wlan_80211n(54))
    assert np.array_equal(tw.inbox_start_chk, g["wlan_sc"]) and np.array_equal(tw.inbox_start_var, g["wlan_sv"])
    assert np.array_equal(tw.target_cells_chk, g["wlan_tc"]) and np.array_equal(tw.target_cells_var, g["wlan_tv"])
    assert (tw.d_c_max, tw.d_v_max) == (int(g["wlan_d_c_max"]), int(g["wlan_d_v_max"]))


def test_rate_and_data_len_match_reference():
    g = load_golden("reference_host_tables")
    Hw = codes.wlan_80211n(54)
    r = graph.code_rate_from_degrees(Hw)
    assert r == float(g["wlan_R_c"]) and int(r * 1296) == int(g["wlan_data_len"]) == 648
    Hd = codes.dvbs2_like_half_rate(6480, q_groups=36)
    rd = graph.code_rate_from_degrees(Hd)
    assert rd == float(g["dvb_small_R_c"]) and int(rd * 6480) == int(g["dvb_small_data_len"])


def test_tables_are_inverse_permutations():
    for H in (codes.regular_random(96, 3, 6, seed=5), codes.wlan_80211n(54)):
        t = graph.edge_tables(H)
        assert np.array_equal(t.target_cells_var[t.target_cells_chk], np.arange(t.n_edge))
        # slot k of check c holds its k-th neighbour in ascending variable order
        for c in (0, t.n_chk - 1):
            s, d = t.inbox_start_chk[c], t.degree_chk[c]
            assert np.all(np.diff(t.var_of_chk_slot[s:s + d]) > 0)


def test_alist_docstring_example_and_roundtrip(tmp_path):
    # the one executable example of the reference (discrete_LDPC_decoder.py:64-67)
    H = graph.alist_to_csr([[3, 2], [2, 2], [1, 1, 2], [2, 2], [1], [2], [1, 2], [1, 2, 3, 4]]).toarray()
    assert np.array_equal(H, [[1, 0, 1], [0, 1, 1]])
    Hr = codes.regular_random(60, 3, 6, seed=2)
    f = str(tmp_path / "c.alist")
    graph.write_alist(Hr, f)
    assert (graph.load_check_matrix(f) != Hr).nnz == 0
    f2 = str(tmp_path / "c.npz")
    codes.save_csr_npz(Hr, f2)
    assert (graph.load_check_matrix(f2) != Hr).nnz == 0
    f3 = str(tmp_path / "c.npy")
    np.save(f3, Hr.toarray().astype(float))
    assert (graph.load_check_matrix(f3) != Hr).nnz == 0


def test_code_generators_have_the_reference_degree_profiles():
    t = graph.edge_tables(codes.regular_random(8000, 3, 6))
    assert (t.n_var, t.n_chk, t.n_edge) == (8000, 4000, 24000)
    assert set(t.degree_chk) == {6} and set(t.degree_var) == {3}
    tw = graph.edge_tables(codes.wlan_80211n(54))
    assert (tw.n_var, tw.n_chk, tw.n_edge) == (1296, 648, 4644)
    assert dict(zip(*np.unique(tw.degree_var, return_counts=True))) == {2: 594, 3: 486, 4: 54, 11: 162}
    assert dict(zip(*np.unique(tw.degree_chk, return_counts=True))) == {7: 540, 8: 108}
    tw2 = graph.edge_tables(codes.wlan_80211n(81))
    assert tw2.n_var == 1944 and set(tw2.degree_chk) == {7, 8}
    td = graph.edge_tables(codes.dvbs2_like_half_rate())
    assert (td.n_var, td.n_chk, td.n_edge) == (64800, 32400, 226799)
    assert dict(zip(*np.unique(td.degree_var, return_counts=True))) == {1: 1, 2: 32399, 3: 19440, 8: 12960}
    assert dict(zip(*np.unique(td.degree_chk, return_counts=True))) == {6: 1, 7: 32399}


def test_lut_lengths_match_appendix_b():
    assert luts.cn_lut_len(16, 16, 6, 50) == 51200 and luts.vn_lut_len(16, 16, 3, 50) == 38400
    assert luts.cn_lut_len(16, 16, 8, 50) == 76800 and luts.vn_lut_len(16, 16, 11, 50) == 140800
    assert luts.match_len(16, 8, 50) == 6400 and luts.match_len(16, 11, 50) == 8800
    assert luts.cn_lut_len(32, 32, 8, 50) == 307200 and luts.vn_lut_len(32, 32, 11, 50) == 563200


def test_config_file_roundtrip(tmp_path):
    tb = luts.random_tables(16, 8, 11, 5, seed=1, matching=True)
    for ext in (".pkl", ".npz"):
        f = str(tmp_path / ("decoder_config_EbN0_gen_0.9_16" + ext))
        luts.save_config(tb, f, sigma_n2=0.5)
        d = luts.load_config(f)
        assert int(d["cardinality_T_decoder_ops"]) == 16 and int(d["imax"]) == 5
        for k in ("Trellis_checknodevector_a", "Trellis_varnodevector_a", "matching_vector_checknode",
                  "matching_vector_varnode"):
            assert np.array_equal(d[k], getattr(tb, k))
    # a pickle written the way the reference does (a plain dict of the generator's attributes)
    f = str(tmp_path / "ref_style.pkl")
    with open(f, "wb") as fh:
        pickle.dump(dict(tb.as_dict(), EbN0=0.9, nror=5), fh)
    assert int(luts.load_config(f)["imax"]) == 5


def test_node_op_helpers_match_reference():
    g = load_golden("reference_host_tables")
    import scipy.sparse as sp
    from informationbottleneckdecodingldpc_b200 import Discrete_LDPC_Decoder_class
    Hr = sp.csr_matrix((np.ones(g["reg_H_indices"].size), g["reg_H_indices"], g["reg_H_indptr"]), shape=tuple(g["reg_shape"]))
    dec = Discrete_LDPC_Decoder_class(Hr, 4, 16, 16, g["reg_cn_lut"], g["reg_vn_lut"], 2)
    assert np.array_equal(dec.discrete_cn_operation(g["yc"], 0), g["cn_op_iter0"])
    assert np.array_equal(dec.discrete_cn_operation(g["yc"], 2), g["cn_op_iter2"])
    assert np.array_equal(dec.discrete_vn_operation(g["yv"], 1), g["vn_op_iter1"])


def test_constructor_validation():
    from informationbottleneckdecodingldpc_b200 import Discrete_LDPC_Decoder_class
    H = codes.regular_random(24, 3, 6, seed=1)
    tb = luts.random_tables(16, 6, 3, 4, seed=1)
    bad = tb.Trellis_checknodevector_a.copy()
    bad[3] = 16
    with pytest.raises(ValueError):
        Discrete_LDPC_Decoder_class(H, 4, 16, 16, bad, tb.Trellis_varnodevector_a, 2)
    with pytest.raises(ValueError):   # irregular H in the regular class
        Discrete_LDPC_Decoder_class(codes.wlan_80211n(54), 4, 16, 16, tb.Trellis_checknodevector_a,
                                    tb.Trellis_varnodevector_a, 2)


def test_quantizer_design_properties():
    from informationbottleneckdecodingldpc_b200 import AWGN_Channel_Quantizer
    q = AWGN_Channel_Quantizer(10 ** (-1.2 / 10) / (2 * 0.5), 3, 16, 2000)
    assert np.all(np.diff(q.limits) > 0) and q.limits[8] == 0
    assert np.allclose(q.output_LLRs, -q.output_LLRs[::-1]) and np.all(np.diff(q.output_LLRs) > 0)
    assert q.cdf_t_given_x_equals_zero[0] == 0 and abs(q.cdf_t_given_x_equals_zero[-1] - 1) < 1e-12
    assert np.array_equal(q.p_t_given_y.sum(1), np.ones(2000))
    x = np.array([[-5.0, -1e-9, 0.0, 1e-9, 5.0]])
    assert np.array_equal(q.quantize_on_host(x)[0], [0, 7, 7, 8, 15])


def test_cabi_library_exports_every_declared_symbol():
    from informationbottleneckdecodingldpc_b200 import _lib
    header = open(os.path.join(ROOT, "include", "ibldpc.h")).read()
    declared = set(re.findall(r"\b(ibldpc_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(dll, name), name
    _lib.lib()


def test_launch_planner_geometry_invariants():
    """ibldpc_plan_geometry (host logic, no GPU call): every choice covers all tiles and all nodes, keeps at least
    one node per CTA step, never plans more CTAs per tile group than node steps, and for the shapes measured on
    B200 picks the footprint that fills the resident slots."""
    from informationbottleneckdecodingldpc_b200 import _lib
    L = _lib.lib()
    out = (ctypes.c_int32 * 3)()

    def plan(slots, warps, tiles, nodes):
        _lib.check(L.ibldpc_plan_geometry(slots, warps, tiles, nodes, out))
        return out[0], out[1], out[2]

    rng = np.random.default_rng(5)
    for _ in range(2000):
        slots = int(rng.integers(1, 8)) * 148
        warps = int(rng.choice([8, 16, 24, 32]))      # 24: the 768-thread tail-pair variable-node kernels
        tiles = int(rng.integers(1, 400))
        nodes = int(rng.integers(1, 70000))
        tpc, tg, gx = plan(slots, warps, tiles, nodes)
        assert 0 <= tpc <= 5 and (warps >> tpc) >= 1
        assert warps % (1 << tpc) == 0                              # whole groups of 2^tpc warps: no (node, tile) is visited twice
        assert tg == -(-tiles // (1 << tpc))                     # all tiles covered, no empty tile group
        assert tpc == 0 or (1 << (tpc - 1)) < tiles              # footprint not wider than the row needs
        nsteps = -(-nodes // (warps >> tpc))
        assert 1 <= gx <= nsteps
        assert gx * tg <= max(slots, tg)                          # one wave unless there are more tile groups than slots
    # 802.11n n=1296 d_v=11 class, B=100096 (196 tiles of 256 B), 148 slots of 16 warps: 49 groups x 3 CTAs = 147 of 148
    assert plan(148, 16, 196, 162) == (2, 49, 3)
    # C1, B=65536, one 1024-thread CTA per SM: all 32 warps of a CTA on consecutive tiles of one node fill all 148 slots
    # (8 tiles per CTA would leave 4 SMs idle: 16 tile groups x 9 CTAs)
    assert plan(148, 32, 128, 4000) == (5, 4, 37) and plan(148, 32, 64, 8000) == (5, 2, 74)
    assert plan(296, 16, 128, 4000) == (4, 8, 37)
    assert plan(148, 24, 10, 6)[0] <= 3                             # few nodes, 10 tiles: a tie must not go to 16 tiles per 24-warp CTA
    assert L.ibldpc_plan_geometry(0, 16, 1, 1, out) != 0 and L.ibldpc_plan_geometry(148, 12, 1, 1, out) != 0


def test_host_chunk_schedule_properties():
    """ibldpc_host_chunk_schedule (host logic): chunks are positive and sum to B; early termination is never
    chunked; the automatic schedule stays under ~256 MiB of channel values per chunk, has quarter-size first and
    last chunks once B >= 8192, and an explicit chunk size gives equal chunks."""
    from informationbottleneckdecodingldpc_b200 import _lib
    L = _lib.lib()
    buf = (ctypes.c_int64 * 4096)()

    def sched(B, n_var, chunk=0, early=0):
        n = L.ibldpc_host_chunk_schedule(B, n_var, chunk, early, buf, 4096)
        assert 1 <= n <= 4096
        return [int(buf[i]) for i in range(n)]

    rng = np.random.default_rng(6)
    for _ in range(500):
        B = int(rng.integers(1, 300000))
        n_var = int(rng.choice([24, 1296, 8000, 64800]))
        w = sched(B, n_var)
        assert all(x > 0 for x in w) and sum(w) == B
        assert max(w) * n_var <= (256 << 20) + 512 * n_var
        target = max(512, ((256 << 20) // n_var) // 512 * 512)
        if B >= 8192 or B > target:
            assert len(w) >= 3 and w[0] <= max(w) and w[-1] <= max(w)
            if B >= 8192:
                assert len(w) >= 4 and w[0] <= max(w) // 2 and w[-1] <= max(w) // 2
        else:
            assert w == [B]
        assert sched(B, n_var, early=1) == [B]
        c = int(rng.integers(16, 40000))
        w = sched(B, n_var, chunk=c)
        assert sum(w) == B and all(x == c // 16 * 16 for x in w[:-1]) and 0 < w[-1] <= c // 16 * 16
    assert sched(65536, 8000) == [8192, 24576, 24576, 8192]
    assert L.ibldpc_host_chunk_schedule(0, 8000, 0, 0, buf, 16) < 0


def test_bench_roofline_object_handles_both_launch_schemes():
    """bench.py's roofline object: per-phase launches (large batches) and the single cooperative launch (small
    batches, no per-phase times) -- pure host arithmetic, checked here without a GPU."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    N, E, B = 8000, 24000, 65536
    r = bench.roofline_from_phase_times([23.3, 19.3, 0.6], [49, 49, 4], N, E, B, 50, True, "c1")
    assert r["kernel"].startswith("ib_cn_n4") and r["algorithmic_bytes_per_launch"] == 2 * E * B
    assert "look-up" in r["bound"] and r["denominator"] == "hbm" and abs(r["frac_stored"] * 2 - r["frac"]) < 1e-12
    assert r["stored_bytes_per_launch"] * 2 == r["algorithmic_bytes_per_launch"]
    assert abs(r["frac"] - r["cn_frac"]) < 1e-12 and r["vn_frac"] > r["cn_frac"]
    assert abs(r["achieved"] - 2 * E * B / (23.3e-3 / 49) / 1e9) < 1e-6
    assert r["whole_decode"]["bytes_per_frame"] == 49 * (4 * E + N) + 2 * E + 3 * N == 5168000
    r = bench.roofline_from_phase_times([0.0, 0.0, 1.0], [0, 0, 2], N, E, 512, 50, True, "c1")
    assert r["kernel"].startswith("ib_decode_coop_kernel") and r["cn_frac"] is None and r["cn_avg_ms"] is None
    assert abs(r["achieved"] - 5168000 * 512 / 1e-3 / 1e9) < 1e-6 and r["traffic"] is None
    r = bench.roofline_from_phase_times([10.0, 12.0, 1.0], [49, 49, 4], N, E, B, 50, False, "c1")
    assert r["kernel"].startswith("ib_vn_fast") and r["stored_bytes_per_launch"] == (2 * E + N) * B
    assert r["whole_decode"]["frac_stored"] == r["whole_decode"]["frac"]


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from informationbottleneckdecodingldpc_b200 import Discrete_LDPC_Decoder_class
    H = codes.regular_random(24, 3, 6, seed=1)
    tb = luts.random_tables(16, 6, 3, 4, seed=1)
    dec = Discrete_LDPC_Decoder_class(H, 4, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, 2)
    with pytest.raises(RuntimeError):
        dec.decode_OpenCL(np.zeros((24, 2), dtype=np.int32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "informationbottleneckdecodingldpc_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "ldpc_oracle" not in txt and "libref_kernels" not in txt, f


def test_ib_design_tool_layout_and_convergence():
    """In-repo replacement of the reference's design chain (decoder_config_generation.py): table
    lengths of SURVEY Appendix B, entries in [0,T), mutual information non-decreasing over the
    iterations and reaching ~1 bit above the decoder's threshold; a mirror-symmetric design
    (t -> T-1-t under x -> 1-x), which every hard decision of the code base relies on."""
    from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_regular_config
    tb, ex = generate_regular_config(1.4, 3, 6, 16, 30)
    cn, vn = tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a
    assert cn.size == luts.cn_lut_len(16, 16, 6, 30) and vn.size == luts.vn_lut_len(16, 16, 3, 30)
    assert cn.min() >= 0 and cn.max() < 16 and vn.min() >= 0 and vn.max() < 16
    mi = ex["ext_mi_varnode_in_iter"]
    assert np.all(np.diff(mi) > -1e-9) and mi[-1] > 0.999
    T = 16
    st = vn[:T * T].reshape(T, T).astype(int)                  # VN stage 0 of iteration 0: [ch][m]
    assert np.array_equal(st, T - 1 - st[::-1, ::-1])
    c0 = cn[:T * T].reshape(T, T).astype(int)                  # CN stage 0: flipping one input flips the output
    assert np.array_equal(c0, T - 1 - c0[::-1, :])
    # below threshold the evolution stalls (no free lunch)
    _, ex2 = generate_regular_config(0.6, 3, 6, 16, 30)
    assert ex2["ext_mi_varnode_in_iter"][-1] < 0.9


def test_decoder_config_cli_writes_reference_named_pickle(tmp_path, monkeypatch):
    from informationbottleneckdecodingldpc_b200 import decoder_config_generation as g
    monkeypatch.chdir(tmp_path)
    g.main(["--ebn0", "1.2", "--imax", "3"])
    d = luts.load_config(str(tmp_path / "decoder_config_EbN0_gen_1.2_16.pkl"))
    assert int(d["imax"]) == 3 and int(d["cardinality_T_decoder_ops"]) == 16
    assert d["Trellis_checknodevector_a"].size == luts.cn_lut_len(16, 16, 6, 3)


def test_irregular_design_tool_with_message_alignment():
    """Irregular stand-in for the reference's WLAN / DVB-S2 decoder_config_generation: table and
    matching-vector lengths of SURVEY Appendix B, identity rows for the reference degree and for unused
    degrees, mirror-symmetric alignment maps, mutual information growing to ~1 bit."""
    from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_irregular_config
    H = codes.wlan_80211n(54)
    tb, ex = generate_irregular_config(1.4, H, 16, 20)
    assert tb.Trellis_checknodevector_a.size == luts.cn_lut_len(16, 16, 8, 20)
    assert tb.Trellis_varnodevector_a.size == luts.vn_lut_len(16, 16, 11, 20)
    mc = tb.matching_vector_checknode.reshape(20, 8, 16).astype(int)
    mv = tb.matching_vector_varnode.reshape(20, 11, 16).astype(int)
    assert mc.min() >= 0 and mc.max() < 16 and mv.min() >= 0 and mv.max() < 16
    ident = np.arange(16)
    for d in (1, 2, 3, 4, 5, 6):                       # no checks of these degrees in the 802.11n code
        assert np.array_equal(mc[:, d - 1], np.tile(ident, (20, 1)))
    assert np.array_equal(mc[:, 6], np.tile(ident, (20, 1)))     # degree 7 = most reliable = reference meaning
    assert np.array_equal(mv, 15 - mv[:, :, ::-1])               # z*(T-1-t) = T-1-z*(t)
    assert np.all(np.diff(mv, axis=2) >= 0)                      # alignment keeps the LLR order
    mi = ex["ext_mi_varnode_in_iter"]
    assert mi[-1] > 0.95 and mi[-1] > mi[0] + 0.25 and np.all(np.diff(mi) > -1e-6)
    assert abs(ex["lambda_vec"].sum() - 1) < 1e-12 and abs(ex["rho_vec"].sum() - 1) < 1e-12


def test_philox_substreams_are_disjoint_per_rank():
    """every rank of a multi-process BER run must draw its own channel realisations."""
    from informationbottleneckdecodingldpc_b200 import AWGN_Channel_Quantizer, rng
    keys = {rng.stream_key(20181001, r) for r in range(64)}
    assert len(keys) == 64 and rng.stream_key(20181001, 0) == 20181001
    assert all(0 <= k < 2 ** 64 for k in keys)
    for seed in (0, 1, 2 ** 63, 2 ** 64 - 1):
        assert rng.stream_key(seed, 0) == seed and rng.stream_key(seed, 5) != seed
    a = AWGN_Channel_Quantizer(0.8, 3, 16, 2000, dont_calc=True)
    b = AWGN_Channel_Quantizer(0.8, 3, 16, 2000, dont_calc=True)
    a.set_stream(0)
    b.set_stream(1)
    assert a._philox_key() != b._philox_key() and a._offset == b._offset == 0
    c = AWGN_Channel_Quantizer(0.8, 3, 16, 2000, dont_calc=True)
    assert c._stream is None and c.stream == 0          # no process group here: stream 0


def test_run_driver_overrides_and_shadowing():
    from informationbottleneckdecodingldpc_b200 import run_driver
    src = "a = 1\nif True:\n    min_errors = 7000\n    x = min_errors == 3\nmin_errors = 9\n"
    out = run_driver.apply_overrides(src, {"min_errors": 200})
    assert out == "a = 1\nif True:\n    min_errors = 200\n    x = min_errors == 3\nmin_errors = 9\n"
    with pytest.raises(KeyError):
        run_driver.apply_overrides(src, {"nope": 1})
    run_driver.install_shadow()
    import importlib
    m = importlib.import_module("Discrete_LDPC_decoding.discrete_LDPC_decoder_irreg")
    assert m.Discrete_LDPC_Decoder_class_irregular.__module__.startswith("informationbottleneckdecodingldpc_b200")
    from AWGN_Channel_Transmission.LDPC_Transmitter import LDPC_BPSK_Transmitter  # noqa: F401
    from Continous_LDPC_Decoding.min_sum_decoder_irreg import Min_Sum_Decoder_class_irregular  # noqa: F401
