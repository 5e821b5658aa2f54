"""GPU suite (-m gpu): the CUDA path (through the class API, i.e. the ctypes C ABI) against the
golden vectors frozen from the reference and against the CPU oracle on seeded inputs; full-size
runs are checked through size-independent properties."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import IB_CASES, LLR_CASES, load_golden
from informationbottleneckdecodingldpc_b200 import codes, graph, luts

pytestmark = pytest.mark.gpu


def _mk_ib(g_or_H, T, imax, cn, vn, mc=None, mv=None, irregular=True, B=1):
    import informationbottleneckdecodingldpc_b200 as pkg
    if irregular:
        return pkg.Discrete_LDPC_Decoder_class_irregular(g_or_H, imax, T, T, cn, vn, mc, mv, B,
                                                         match='true' if mc is not None else 'false')
    return pkg.Discrete_LDPC_Decoder_class(g_or_H, imax, T, T, cn, vn, B)


def _dev(a):
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    return pkg.DeviceArray(torch.from_numpy(np.ascontiguousarray(a)).cuda())


# Kernel families of the IB fast path (|T| <= 16), selected when the tables are uploaded:
#   n4       packed-nibble messages (default), per-degree default vector widths
#   n4_vn4 / n4_vn2  packed nibbles with the variable-node vector width forced to 4 (degree <= 6) / 2 words
#   u8       one byte per message (IBLDPC_NO_NIBBLE=1), incl. the tail-pair check-node kernels
#   n4_pair4 / n4_nopair  tail-pair check-node kernels from degree 4 on / no tail-pair kernels at all
#   n4_vpair3 / n4_vpair3_256  tail-pair variable-node kernels from degree 3 on, 512- / 256-thread CTAs
#   n4_small_ctas  512- / 256-thread CTAs instead of the default 1024-thread ones (per-phase launches)
#   n4_nocoop  per-phase launches also for small batches (default: one cooperative whole-decode kernel up to 4096 frames);
#              for the instantiated degree sets these are the fused per-phase kernels of ib_phase_n4.cuh (one launch per
#              phase over all degree classes, TMA-staged table image), n4_nophase = one launch per degree class instead
#              (degree-3 variable nodes and check nodes of degree 6..8 through the three-input tables of ib_triple_n4.cuh;
#              n4_notriple = without them, n4_cn_tri6 = check nodes up to degree 6 only)
IB_VARIANTS = {"n4": {}, "n4_coop": {"IBLDPC_COOP_MAX_B": "4096", "IBLDPC_NO_COOP_PHASE": "1"},   # table-restaging cooperative kernels of ib_coop_n4.cuh up to 4096 frames (default: see the batch-size policy at the end of ibldpc_set_luts)
               "n4_vn4": {"IBLDPC_VN_VEC": "4"}, "n4_vn2": {"IBLDPC_VN_VEC": "2"}, "n4_pair4": {"IBLDPC_PAIR_MIN_DEGREE": "4"},
               "n4_vpair3": {"IBLDPC_VN_PAIR_MIN_DEGREE": "3"},
               "n4_vpair3_256": {"IBLDPC_VN_PAIR_MIN_DEGREE": "3", "IBLDPC_VN_PAIR_THREADS": "256"},
               "n4_nocoop": {"IBLDPC_COOP_MAX_B": "0", "IBLDPC_PHASE": "1"},   # small batches too through the fused per-phase kernels (every instantiated degree set)
               "n4_nophase": {"IBLDPC_COOP_MAX_B": "0", "IBLDPC_NO_PHASE": "1"},   # one launch per degree class (round-1 default)
               "n4_small_ctas": {"IBLDPC_COOP_MAX_B": "0", "IBLDPC_CN_THREADS": "512", "IBLDPC_VN_THREADS": "256"},
               "n4_notriple": {"IBLDPC_COOP_MAX_B": "0", "IBLDPC_NO_PHASE": "1", "IBLDPC_NO_TRIPLE": "1"},   # degree-3 variable nodes through the two-input stage tables
               "n4_cn_tri6": {"IBLDPC_COOP_MAX_B": "0", "IBLDPC_NO_PHASE": "1", "IBLDPC_CN_TRI_MAX_DEGREE": "6"},   # check nodes of degree 7, 8 through the plain tail-pair kernels (default: three-input table up to degree 8)
               "n4_nopair": {"IBLDPC_NO_PAIR": "1"}, "u8": {"IBLDPC_NO_NIBBLE": "1"}}


@pytest.fixture(params=list(IB_VARIANTS))
def ib_variant(request, monkeypatch):
    for k, v in IB_VARIANTS[request.param].items():
        monkeypatch.setenv(k, v)
    return 1 if request.param == "u8" else 2     # expected ibldpc_info()[0] of the fast path


@pytest.mark.parametrize("force_generic", [False, True])
@pytest.mark.parametrize("case", IB_CASES)
def test_ib_golden_device_buffers(gpu, case, force_generic, monkeypatch, ib_variant):
    g = load_golden(case)
    if force_generic:
        if ib_variant != 2 or any(os.environ.get(k) for k in ("IBLDPC_VN_VEC", "IBLDPC_PAIR_MIN_DEGREE", "IBLDPC_NO_PAIR", "IBLDPC_VN_PAIR_MIN_DEGREE", "IBLDPC_COOP_MAX_B", "IBLDPC_CN_THREADS", "IBLDPC_NO_PHASE")):
            pytest.skip("the generic path has one variant")
        monkeypatch.setenv("IBLDPC_FORCE_GENERIC", "1")
    T, imax = int(g["T"]), int(g["imax"])
    match = bool(int(g["match"]))
    irregular = bool(int(g["irregular"])) or match
    dec = _mk_ib(g["H"], T, imax, g["cn_lut"], g["vn_lut"], g["cn_match"] if match else None,
                 g["vn_match"] if match else None, irregular, g["ch"].shape[1])
    dec.init_OpenCL_decoding(g["ch"].shape[1])
    dec.early_termination = bool(int(g["early"]))
    out = dec.decode_OpenCL(_dev(g["ch"]), buffer_in=True, return_buffer=True)
    fast = dec.info()[0]
    assert fast == (0 if force_generic else 3 if T > 16 else ib_variant)
    assert np.array_equal(out.get(), g["out"]), case
    if dec.early_termination:
        assert dec.last_i_num == int(g["i_num"])
    # error counter: all-zero codeword, rows = N (regular) or data_len (irregular)
    rows = dec.N_v if not irregular else int(dec.data_len)
    assert dec.return_errors_all_zero(out) == int((g["out"][:rows] < T // 2).sum())


@pytest.mark.parametrize("case", ["ib_toy_3_6_n24", "ib_wlan_T16_match", "ib_c1_minsumlut_imax50_et", "ib_irreg_deg1vn"])
def test_ib_golden_host_buffers(gpu, case):
    """numpy in -> numpy int32 out (decode_OpenCL(buffer_in=False, return_buffer=False))."""
    g = load_golden(case)
    T, imax = int(g["T"]), int(g["imax"])
    match = bool(int(g["match"]))
    irregular = bool(int(g["irregular"])) or match
    dec = _mk_ib(g["H"], T, imax, g["cn_lut"], g["vn_lut"], g["cn_match"] if match else None,
                 g["vn_match"] if match else None, irregular)
    dec.init_OpenCL_decoding(g["ch"].shape[1])
    dec.early_termination = bool(int(g["early"]))
    out = dec.decode_OpenCL(g["ch"].astype(np.int32), buffer_in=False, return_buffer=False)
    assert out.dtype == np.int32 and np.array_equal(out, g["out"])
    assert dec.last_i_num == int(g["i_num"])
    # decode_on_host: one frame, no early termination
    if not int(g["early"]):
        v = dec.decode_on_host(g["ch"][:, 0])
        assert np.array_equal(v.astype(np.uint8), g["out"][:, 0])


def _oracle_ib(t, ch, T, imax, tb, early):
    from oracle import oracle
    return oracle.ib_decode(t, ch, T=T, imax=imax, cn_lut=tb.Trellis_checknodevector_a,
                            vn_lut=tb.Trellis_varnodevector_a, cn_match=tb.matching_vector_checknode,
                            vn_match=tb.matching_vector_varnode, early=early)


@pytest.mark.parametrize("B", [1, 15, 16, 31, 33, 100, 513, 1040, 2049])
def test_ib_c1_vs_oracle_ragged_batches(gpu, B, ib_variant):
    """(3,6) n=8000, random tables: every batch-size class (sub-vector, unaligned, multi-tile)."""
    H = codes.regular_random(8000, 3, 6)
    t = graph.edge_tables(H)
    T, imax = 16, 6
    tb = luts.random_tables(T, 6, 3, imax, seed=B)
    ch = np.random.Generator(np.random.PCG64(B)).integers(0, T, size=(8000, B)).astype(np.uint8)
    dec = _mk_ib(H, T, imax, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, irregular=False)
    dec.early_termination = False
    got = dec.decode_OpenCL(_dev(ch), buffer_in=True, return_buffer=True).get()
    ref, _ = _oracle_ib(t, ch, T, imax, tb, False)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("name,H,T", [
    ("wlan1296", lambda: codes.wlan_80211n(54), 16),
    ("wlan1944", lambda: codes.wlan_80211n(81), 16),
    ("wlan1296_T8", lambda: codes.wlan_80211n(54), 8),
    ("dvb6480", lambda: codes.dvbs2_like_half_rate(6480, q_groups=36), 16),
    ("deg2checks", lambda: codes.random_from_degrees([1, 1] + [2] * 30 + [3] * 20 + [5] * 4,
                                                     [2] * 10 + [3] * 10 + [4] * 11 + [5] * 4 + [6] * 3 + [10], seed=9), 16),
    ("wlan1296_T12", lambda: codes.wlan_80211n(54), 12),
])
@pytest.mark.parametrize("match", [True, False])
def test_ib_irregular_vs_oracle(gpu, name, H, T, match, ib_variant):
    H = H()
    t = graph.edge_tables(H)
    imax, B = 7, 77
    tb = luts.random_tables(T, t.d_c_max, t.d_v_max, imax, seed=11, matching=match)
    ch = np.random.Generator(np.random.PCG64(12)).integers(0, T, size=(t.n_var, B)).astype(np.uint8)
    dec = _mk_ib(H, T, imax, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                 tb.matching_vector_checknode, tb.matching_vector_varnode)
    got = dec.decode_OpenCL(_dev(ch), buffer_in=True, return_buffer=True).get()
    ref, i_num = _oracle_ib(t, ch, T, imax, tb, True)
    assert dec.info()[0] == ib_variant
    assert np.array_equal(got, ref) and dec.last_i_num == i_num


@pytest.mark.parametrize("family", ["n4", "n4_notriple", "u8"])
def test_ib_tail_pair_variant_all_degrees(gpu, monkeypatch, family):
    """The composed tail-pair check-node kernels (cn_word_pair / cn_word_n4_pair) for every degree 4..10,
    forced on with IBLDPC_PAIR_MIN_DEGREE=4, with and without message alignment, against the oracle.  In the packed
    family the degrees 6..8 run the three-input-table kernels (ib_cn_n4_tri_kernel); n4_notriple keeps them on the
    plain tail-pair kernels."""
    monkeypatch.setenv("IBLDPC_PAIR_MIN_DEGREE", "4")
    if family == "u8":
        monkeypatch.setenv("IBLDPC_NO_NIBBLE", "1")
    if family == "n4_notriple":
        monkeypatch.setenv("IBLDPC_NO_TRIPLE", "1")
    H = codes.random_from_degrees([2] * 40 + [3] * 40 + [4] * 16, [4] * 16 + [5] * 8 + [6] * 8 + [7] * 6 + [8] * 4 + [9] * 2 + [10] * 2, seed=4)
    t = graph.edge_tables(H)
    assert sorted(set(t.degree_chk)) == [4, 5, 6, 7, 8, 9, 10]
    for T in (16, 8):
        for match in (True, False):
            imax, B = 6, 50
            tb = luts.random_tables(T, t.d_c_max, t.d_v_max, imax, seed=31, matching=match)
            ch = np.random.Generator(np.random.PCG64(32)).integers(0, T, size=(t.n_var, B)).astype(np.uint8)
            dec = _mk_ib(H, T, imax, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                         tb.matching_vector_checknode, tb.matching_vector_varnode)
            got = dec.decode_OpenCL(_dev(ch), buffer_in=True, return_buffer=True).get()
            ref, i_num = _oracle_ib(t, ch, T, imax, tb, True)
            assert np.array_equal(got, ref) and dec.last_i_num == i_num


@pytest.mark.parametrize("threads", ["256", "512", "768"])
def test_ib_vn_tail_pair_variant_all_degrees(gpu, monkeypatch, threads):
    """The composed tail-pair variable-node kernels (vn_word_n4_pair) for every degree 3..12, forced on with
    IBLDPC_VN_PAIR_MIN_DEGREE=3, both CTA sizes, with and without message alignment, against the oracle."""
    monkeypatch.setenv("IBLDPC_VN_PAIR_MIN_DEGREE", "3")
    monkeypatch.setenv("IBLDPC_VN_PAIR_THREADS", threads)
    H = codes.random_from_degrees([d for d in range(3, 13) for _ in range(4)], [6] * 50, seed=6)
    t = graph.edge_tables(H)
    assert sorted(set(t.degree_var)) == list(range(3, 13))
    for T in (16, 8):
        for match in (True, False):
            imax = 5
            tb = luts.random_tables(T, t.d_c_max, t.d_v_max, imax, seed=41, matching=match)
            # 5000 frames = 10 tiles per row with four nodes per class: the launch planner must not put a 24-warp CTA
            # (768 threads) on 16 tiles of one node (warps 16..23 would visit the next node's tiles a second time)
            for B in (70, 5000):
                ch = np.random.Generator(np.random.PCG64(42)).integers(0, T, size=(t.n_var, B)).astype(np.uint8)
                dec = _mk_ib(H, T, imax, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                             tb.matching_vector_checknode, tb.matching_vector_varnode)
                got = dec.decode_OpenCL(_dev(ch), buffer_in=True, return_buffer=True).get()
                ref, i_num = _oracle_ib(t, ch, T, imax, tb, True)
                assert dec.info()[0] == 2
                assert np.array_equal(got, ref) and dec.last_i_num == i_num


@pytest.mark.parametrize("B,chunk", [(9001, 0), (8200, 0), (5000, 1008), (777, 0)])
def test_ib_host_pipeline_chunk_schedules(gpu, B, chunk):
    """ibldpc_decode_ib_host (pinned-host two-slot pipeline): the automatic ramped chunk schedule and an explicit
    chunk size, ragged batch sizes, must give exactly what one device-buffer decode of the whole batch gives."""
    import informationbottleneckdecodingldpc_b200 as pkg
    from informationbottleneckdecodingldpc_b200 import _lib
    H = codes.wlan_80211n(54)
    t = graph.edge_tables(H)
    T, imax = 16, 5
    tb = luts.random_tables(T, t.d_c_max, t.d_v_max, imax, seed=51, matching=True)
    ch = np.random.Generator(np.random.PCG64(52)).integers(0, T, size=(t.n_var, B)).astype(np.uint8)
    dec = _mk_ib(H, T, imax, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                 tb.matching_vector_checknode, tb.matching_vector_varnode)
    dec.early_termination = False
    dec.host_output_dtype = np.uint8
    full = dec.decode_OpenCL(_dev(ch), buffer_in=True, return_buffer=True).get()
    _lib.check(_lib.lib().ibldpc_set_host_chunk(dec._ensure_handle(), chunk))
    host_in = pkg.pinned_empty(ch.shape, np.uint8)
    host_in[:] = ch
    host = dec.decode_OpenCL(host_in, buffer_in=False, return_buffer=False)
    assert np.array_equal(host, full)
    ref, _ = _oracle_ib(t, ch[:, -40:], T, imax, tb, False)
    assert np.array_equal(host[:, -40:], ref)


def test_ib_dvbs2_full_size_vs_oracle(gpu):
    """DVB-S2-like n=64800 (degree-1 VN, d_v 8, d_c 6/7, matching), a few frames, 4 iterations."""
    H = codes.dvbs2_like_half_rate()
    t = graph.edge_tables(H)
    T, imax, B = 16, 4, 20
    tb = luts.random_tables(T, 7, 8, imax, seed=21, matching=True)
    ch = np.random.Generator(np.random.PCG64(22)).integers(0, T, size=(t.n_var, B)).astype(np.uint8)
    dec = _mk_ib(H, T, imax, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                 tb.matching_vector_checknode, tb.matching_vector_varnode)
    assert int(dec.data_len) == 32399          # the reference's float arithmetic gives int(0.4999..*64800)
    got = dec.decode_OpenCL(_dev(ch), buffer_in=True, return_buffer=True).get()
    ref, _ = _oracle_ib(t, ch, T, imax, tb, True)
    assert np.array_equal(got, ref)


def test_ib_generic_path_T32_and_Tc_ne_T(gpu):
    from oracle import oracle
    H = codes.wlan_80211n(54)
    t = graph.edge_tables(H)
    imax, B = 4, 9
    for T, Tc in ((32, 32), (16, 8), (20, 20)):
        tb = luts.random_tables(T, 8, 11, imax, seed=5, Tc=Tc, matching=True)
        ch = np.random.Generator(np.random.PCG64(6)).integers(0, Tc, size=(t.n_var, B)).astype(np.uint8)
        import informationbottleneckdecodingldpc_b200 as pkg
        dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, imax, Tc, T, tb.Trellis_checknodevector_a,
                                                        tb.Trellis_varnodevector_a, tb.matching_vector_checknode,
                                                        tb.matching_vector_varnode, B)
        got = dec.decode_OpenCL(_dev(ch), buffer_in=True, return_buffer=True).get()
        ref, i_num = oracle.ib_decode(t, ch, T=T, Tc=Tc, imax=imax, cn_lut=tb.Trellis_checknodevector_a,
                                      vn_lut=tb.Trellis_varnodevector_a, cn_match=tb.matching_vector_checknode,
                                      vn_match=tb.matching_vector_varnode, early=True)
        assert dec.info()[0] == (3 if (T > 16 and Tc == T) else 0)      # |T| <= 32 family / generic path
        assert np.array_equal(got, ref) and dec.last_i_num == i_num


def test_ib_full_size_properties(gpu):
    """BASELINE size (3,6) n=8000, i_max=50, B=16384 (too large for the oracle): determinism,
    frame independence (sub-batches and a column permutation give the same per-frame answer), the
    host-buffer pipeline equals the device path, and the all-zero codeword is decoded error-free
    from a converging input."""
    import torch
    H = codes.regular_random(8000, 3, 6)
    T, imax, B = 16, 50, 16384
    tb = luts.minsum_like_tables(T, 6, 3, imax)
    rng = np.random.Generator(np.random.PCG64(77))
    ch = np.clip(np.round(rng.normal(T * 0.70, T * 0.15, size=(8000, B))), 0, T - 1).astype(np.uint8)
    dec = _mk_ib(H, T, imax, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, irregular=False)
    dec.early_termination = False
    d_ch = torch.from_numpy(ch).cuda()
    out1 = dec.decode_OpenCL(d_ch, buffer_in=True, return_buffer=True).tensor
    out2 = dec.decode_OpenCL(d_ch, buffer_in=True, return_buffer=True).tensor
    assert torch.equal(out1, out2)
    assert dec.return_errors_all_zero(out1) == 0
    bit, frame = dec.count_errors(out1)
    assert (bit, frame) == (0, 0)
    perm = torch.randperm(B, device="cuda")
    outp = dec.decode_OpenCL(d_ch[:, perm].contiguous(), buffer_in=True, return_buffer=True).tensor
    assert torch.equal(outp, out1[:, perm])
    sub = dec.decode_OpenCL(d_ch[:, 1000:1777].contiguous(), buffer_in=True, return_buffer=True).tensor
    assert torch.equal(sub, out1[:, 1000:1777])
    # random tables: outputs are "random" but must still be frame-independent and equal to the host pipeline
    tbr = luts.random_tables(T, 6, 3, imax, seed=3)
    decr = _mk_ib(H, T, imax, tbr.Trellis_checknodevector_a, tbr.Trellis_varnodevector_a, irregular=False)
    decr.early_termination = False
    decr.host_output_dtype = np.uint8
    chr_ = rng.integers(0, T, size=(8000, B)).astype(np.uint8)
    full = decr.decode_OpenCL(torch.from_numpy(chr_).cuda(), buffer_in=True, return_buffer=True).get()
    host = decr.decode_OpenCL(chr_, buffer_in=False, return_buffer=False)
    assert np.array_equal(full, host)
    from oracle import oracle
    t = graph.edge_tables(H)
    ref, _ = oracle.ib_decode(t, chr_[:, 5000:5016], T=T, imax=imax, cn_lut=tbr.Trellis_checknodevector_a,
                              vn_lut=tbr.Trellis_varnodevector_a, early=False)
    assert np.array_equal(full[:, 5000:5016], ref)


def test_ib_designed_tables_decode_the_all_zero_codeword(gpu):
    """End-to-end sanity with tables from the in-repo design tool: quantizer -> direct sampling on the
    device -> IB decoder; at Eb/N0 = 1.7 dB the (3,6) n=8000 code must decode every frame, stop early,
    and agree bit for bit (and in i_num) with the oracle on the same device-drawn inputs."""
    import informationbottleneckdecodingldpc_b200 as pkg
    from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_regular_config
    from oracle import oracle
    H = codes.regular_random(8000, 3, 6)
    tb, _ = generate_regular_config(1.2, 3, 6, 16, 50)
    q = pkg.AWGN_Channel_Quantizer(10 ** (-1.7 / 10) / (2 * 0.5), 3, 16, 2000)
    q.init_OpenCL_quanti(8000, 48, return_buffer_only=True)
    dec = pkg.Discrete_LDPC_Decoder_class(H, 50, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, 48)
    dec.init_OpenCL_decoding(48, q.context)
    rec = q.quantize_direct_OpenCL(8000, 48)
    out = dec.decode_OpenCL(rec, buffer_in=True, return_buffer=True)
    assert dec.return_errors_all_zero(out) == 0
    assert 5 < dec.last_i_num < 50
    ref, i_num = oracle.ib_decode(graph.edge_tables(H), rec.get(), T=16, imax=50, cn_lut=tb.Trellis_checknodevector_a,
                                  vn_lut=tb.Trellis_varnodevector_a, early=True)
    assert np.array_equal(out.get(), ref) and dec.last_i_num == i_num


def test_ib_designed_irregular_tables_with_alignment(gpu):
    """802.11n code with tables + matching vectors from the in-repo irregular design: at 2.2 dB every
    frame is decoded when message alignment is on, and the result equals the oracle's bit for bit;
    with match='false' the same tables leave errors (alignment matters, as in the reference's papers)."""
    import informationbottleneckdecodingldpc_b200 as pkg
    from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_irregular_config
    from oracle import oracle
    H = codes.wlan_80211n(54)
    tb, _ = generate_irregular_config(1.0, H, 16, 50)
    q = pkg.AWGN_Channel_Quantizer(10 ** (-2.2 / 10) / (2 * 0.5), 3, 16, 2000)
    q.init_OpenCL_quanti(1296, 400, return_buffer_only=True)
    rec = q.quantize_direct_OpenCL(1296, 400)
    errs = {}
    for match in ('true', 'false'):
        dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, 50, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                        tb.matching_vector_checknode, tb.matching_vector_varnode, 400, match=match)
        dec.init_OpenCL_decoding(400, q.context)
        dec.early_termination = False
        out = dec.decode_OpenCL(rec, buffer_in=True, return_buffer=True)
        errs[match] = dec.return_errors_all_zero(out)
        ref, _ = oracle.ib_decode(graph.edge_tables(H), rec.get(), T=16, imax=50, cn_lut=tb.Trellis_checknodevector_a,
                                  vn_lut=tb.Trellis_varnodevector_a,
                                  cn_match=tb.matching_vector_checknode if match == 'true' else None,
                                  vn_match=tb.matching_vector_varnode if match == 'true' else None, early=False)
        assert np.array_equal(out.get(), ref)
    assert errs['true'] == 0 and errs['false'] > 0


def test_early_termination_is_batch_granular(gpu):
    """One noisy frame keeps the whole batch iterating (reference stop rule, decoder.py:273)."""
    g = load_golden("ib_c1_minsumlut_imax50_et")
    T, imax = int(g["T"]), int(g["imax"])
    dec = _mk_ib(g["H"], T, imax, g["cn_lut"], g["vn_lut"], irregular=False)
    dec.decode_OpenCL(_dev(g["ch"]), buffer_in=True, return_buffer=True)
    assert dec.last_i_num == int(g["i_num"]) < imax
    ch = g["ch"].copy()
    ch[:, 3] = np.random.Generator(np.random.PCG64(1)).integers(0, T, size=ch.shape[0])
    from oracle import oracle
    t = graph.edge_tables(g["H"])
    ref, i_ref = oracle.ib_decode(t, ch, T=T, imax=imax, cn_lut=g["cn_lut"], vn_lut=g["vn_lut"], early=True)
    got = dec.decode_OpenCL(_dev(ch), buffer_in=True, return_buffer=True).get()
    assert dec.last_i_num == i_ref == imax
    assert np.array_equal(got, ref)


# ---------------------------------------------------------------------------------- min-sum / BP
@pytest.mark.parametrize("bp_order", ["sequential", "forward_backward", "forward_backward_log"])
@pytest.mark.parametrize("case", LLR_CASES)
def test_llr_golden_float64_exact(gpu, case, bp_order, monkeypatch):
    """float64 messages: min-sum bit-identical to the reference; BP in the reference's operation order
    (IBLDPC_BP_SEQUENTIAL=1) to 1e-9 (exp/log rounding), in the default forward/backward order (likelihood-ratio domain;
    IBLDPC_BP_LOGDOMAIN=1: the reference box-plus expression per operation) to 1e-6 with identical hard decisions and
    stop iteration."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    if bp_order == "sequential":
        monkeypatch.setenv("IBLDPC_BP_SEQUENTIAL", "1")
    if bp_order == "forward_backward_log":
        monkeypatch.setenv("IBLDPC_BP_LOGDOMAIN", "1")
    g = load_golden(case)
    imax = int(g["imax"])
    for cls, algo, meth in ((pkg.Min_Sum_Decoder_class_irregular, "minsum", "decode_OpenCL_min_sum"),
                            (pkg.BeliefPropagationDecoderClassIrregular, "bp", "decode_OpenCL_belief_propagation")):
        dec = cls(g["H"], imax, 16, g["ch"].shape[1])
        dec.init_OpenCL_decoding(g["ch"].shape[1])
        dec.early_termination = bool(int(g["early"]))
        out = getattr(dec, meth)(_dev(g["ch"]), buffer_in=True, return_buffer=True)
        assert out.tensor.dtype == torch.float64
        assert dec.last_i_num == int(g[f"i_num_{algo}"])
        if algo == "minsum":
            assert np.array_equal(out.get(), g["out_minsum"])
        elif bp_order == "sequential":
            assert np.allclose(out.get(), g["out_bp"], rtol=1e-9, atol=1e-9)
        else:
            assert np.allclose(out.get(), g["out_bp"], rtol=1e-6, atol=1e-6)
            assert np.array_equal(out.get() < 0, g["out_bp"] < 0)
        assert dec.return_errors_all_zero(out) == int((g[f"out_{algo}"][:int(dec.data_len)] < 0).sum())


@pytest.mark.parametrize("case", LLR_CASES)
def test_llr_golden_float32_tolerance(gpu, case):
    """fp32 messages (the opt-in fast path) against the float64 golden LLRs: |error| <= 1e-3*(1+|LLR|)
    and identical hard decisions wherever the reference LLR is not a rounding-level tie (|LLR| > 1e-3)."""
    import informationbottleneckdecodingldpc_b200 as pkg
    g = load_golden(case)
    imax = int(g["imax"])
    for cls, algo, meth in ((pkg.Min_Sum_Decoder_class_irregular, "minsum", "decode_OpenCL_min_sum"),
                            (pkg.BeliefPropagationDecoderClassIrregular, "bp", "decode_OpenCL_belief_propagation")):
        dec = cls(g["H"], imax, 16, g["ch"].shape[1])
        dec.precision = 'f32'
        dec.early_termination = bool(int(g["early"]))
        out = getattr(dec, meth)(g["ch"], buffer_in=False, return_buffer=False)   # numpy in, f32 on device
        ref = g[f"out_{algo}"]
        if dec.last_i_num != int(g[f"i_num_{algo}"]):
            pytest.skip("fp32 rounding moved the batch-wide stop by one pass")
        assert np.all(np.abs(out - ref) <= 1e-3 * (1 + np.abs(ref)))
        safe = np.abs(ref) > 1e-3
        assert np.array_equal((out < 0)[safe], (ref < 0)[safe])


def _wlan_llr_batch(B, ebn0_db, seed):
    import informationbottleneckdecodingldpc_b200 as pkg
    H = codes.wlan_80211n(54)
    q = pkg.AWGN_Channel_Quantizer(10 ** (-ebn0_db / 10) / (2 * 0.5), 3, 16, 2000)
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.random(size=(1296, B))
    cl = ((u[:, :, None] - q.cdf_t_given_x_equals_zero) > 0).sum(2) - 1
    return H, q.output_LLRs[cl]


@pytest.mark.parametrize("bp_order", ["forward_backward", "sequential"])
def test_llr_float64_frame_agreement_large_batch(gpu, monkeypatch, bp_order):
    """The stated tolerance of BASELINE.md section 5 on 20000 frames of the WLAN code (Eb/N0 = 2 dB,
    20 iterations, 16-level channel LLRs), float64 GPU path vs the float64 oracle: identical hard
    decisions on >= 99.99 % of frames (min-sum: on all of them, the LLRs are bit-identical) and BER
    inside the 95 % Monte-Carlo confidence interval of the oracle's BER."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    from oracle import oracle
    # float64 BP runs the forward/backward box-plus recursion in the likelihood-ratio domain by default (what bench.py
    # times); IBLDPC_BP_SEQUENTIAL=1 evaluates the reference's operation order.  Both must meet the bar.
    if bp_order == "sequential":
        monkeypatch.setenv("IBLDPC_BP_SEQUENTIAL", "1")
    B, imax = 20000, 20
    H, ch = _wlan_llr_batch(B, 2.0, 5)
    t = graph.edge_tables(H)
    for cls, algo, meth in ((pkg.Min_Sum_Decoder_class_irregular, "minsum", "decode_OpenCL_min_sum"),
                            (pkg.BeliefPropagationDecoderClassIrregular, "bp", "decode_OpenCL_belief_propagation")):
        if algo == "minsum" and bp_order == "sequential":
            continue
        dec = cls(H, imax, 16, B)
        dec.early_termination = False
        got = getattr(dec, meth)(torch.from_numpy(ch).cuda(), buffer_in=True, return_buffer=True).get()
        ref, _ = oracle.llr_decode(t, ch, algo=algo, imax=imax, early=False)
        if algo == "minsum":
            assert np.array_equal(got, ref)
        else:
            print("BP float64", bp_order, "max |LLR - oracle|", float(np.abs(got - ref).max()))
        frames_equal = np.all((got < 0) == (ref < 0), axis=0)
        assert frames_equal.mean() >= 0.9999, (algo, frames_equal.mean())
        ber_ref, ber_got = (ref[:648] < 0).mean(), (got[:648] < 0).mean()
        ci = 1.96 * np.sqrt(max(ber_ref, 1e-9) * (1 - ber_ref) / (648 * B)) + 1e-7
        assert abs(ber_got - ber_ref) <= ci, (algo, ber_got, ber_ref)


def test_llr_float32_fast_path_statistics(gpu):
    """fp32 fast path on the same batch.  With 16-level channel LLRs many a-posteriori sums are exact
    ties in real arithmetic, so their sign is decided by rounding noise in ANY precision (float64
    included); fp32 therefore cannot reproduce float64 decisions frame by frame.  What must hold:
    BER within the Monte-Carlo interval of the float64 oracle, and >= 97 % of frames identical."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    from oracle import oracle
    B, imax = 20000, 20
    H, ch = _wlan_llr_batch(B, 2.0, 5)
    t = graph.edge_tables(H)
    for cls, algo, meth in ((pkg.Min_Sum_Decoder_class_irregular, "minsum", "decode_OpenCL_min_sum"),
                            (pkg.BeliefPropagationDecoderClassIrregular, "bp", "decode_OpenCL_belief_propagation")):
        dec = cls(H, imax, 16, B)
        dec.early_termination = False
        got = getattr(dec, meth)(torch.from_numpy(ch.astype(np.float32)).cuda(), buffer_in=True, return_buffer=True).get()
        ref, _ = oracle.llr_decode(t, ch, algo=algo, imax=imax, early=False)
        frames_equal = np.all((got < 0) == (ref < 0), axis=0)
        assert frames_equal.mean() >= 0.97, (algo, frames_equal.mean())
        ber_ref, ber_got = (ref[:648] < 0).mean(), (got[:648] < 0).mean()
        ci = 3 * np.sqrt(max(ber_ref, 1e-9) * (1 - ber_ref) / (648 * B)) + 1e-6
        assert abs(ber_got - ber_ref) <= ci, (algo, ber_got, ber_ref)
        print(algo, "fp32 vs f64 frame agreement", frames_equal.mean(), "BER", ber_got, ber_ref)


@pytest.mark.parametrize("B", [1, 2, 5, 8])
def test_minsum_low_batch_warp_shuffle_path(gpu, B, monkeypatch):
    """msg_at_time = 2 (the reference's DVB-S2 / WLAN min-sum drivers): the edge-per-lane warp-shuffle
    check-node kernel must give the same float64 LLRs, bit for bit, as the oracle and as the
    frame-per-lane kernel."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    from oracle import oracle
    for H in (codes.wlan_80211n(54), codes.dvbs2_like_half_rate(6480, q_groups=36), codes.regular_random(96, 2, 4, seed=3)):
        t = graph.edge_tables(H)
        rng = np.random.Generator(np.random.PCG64(B))
        vals = np.concatenate([-np.sort(np.abs(rng.normal(0, 3, 8)))[::-1], np.sort(np.abs(rng.normal(0, 3, 8)))])
        ch = vals[rng.integers(0, 16, size=(t.n_var, B))]
        ch[3, 0] = 0.0
        ref, i_ref = oracle.llr_decode(t, ch, algo="minsum", imax=9, early=True)
        for no_shfl in (False, True):
            if no_shfl:
                monkeypatch.setenv("IBLDPC_NO_SHFL", "1")
            else:
                monkeypatch.delenv("IBLDPC_NO_SHFL", raising=False)
            dec = pkg.Min_Sum_Decoder_class_irregular(H, 9, 16, B)
            got = dec.decode_OpenCL_min_sum(torch.from_numpy(ch).cuda(), buffer_in=True, return_buffer=True).get()
            assert np.array_equal(got, ref) and dec.last_i_num == i_ref


# ---------------------------------------------------------------------------------- quantizer
def test_quantizer_golden(gpu):
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    g = load_golden("quantizer")
    q = pkg.AWGN_Channel_Quantizer(0.5, 3, int(g["T"]), 2000, dont_calc=True)
    q.limits = g["limits"]
    q.cdf_t_given_x_equals_zero = g["cdf"]
    q.output_LLRs = g["llr_values"][:-1]
    q.init_OpenCL_quanti(40, 7)
    assert np.array_equal(q.quantize_OpenCL(g["x"]), g["clusters"])
    q.return_buffer_only = True
    assert np.array_equal(q.quantize_OpenCL(torch.from_numpy(g["x"]).cuda()).get(), g["clusters"])


def test_direct_sampling_matches_inversion_of_its_own_uniforms(gpu):
    """quantize_direct_OpenCL draws u on the device; ibldpc_uniform exposes the same u, and the
    oracle's quantize kernel applied to it must give the same clusters / LLRs."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    from informationbottleneckdecodingldpc_b200 import _lib
    from oracle import oracle
    q = pkg.AWGN_Channel_Quantizer(10 ** (-1.2 / 10) / (2 * 0.5), 3, 16, 2000)
    N, B = 300, 40
    q.init_OpenCL_quanti(N, B, return_buffer_only=True)
    q.llr_dtype = np.float64
    cl = q.quantize_direct_OpenCL(N, B).get()
    llr = q.quantize_direct_OpenCL_LLR(N, B).get()
    u = torch.empty(2 * N * B, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().ibldpc_uniform(0, q.seed, 0, 2 * N * B, C.c_void_p(u.data_ptr()), None))
    torch.cuda.synchronize()
    u = u.cpu().numpy()
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    assert np.array_equal(cl, oracle.quantize(u[:N * B].reshape(N, B), q.cdf_t_given_x_equals_zero, 17))
    ref_llr = oracle.quantize_llr(u[N * B:].reshape(N, B), q.cdf_t_given_x_equals_zero, 17,
                                  np.append(q.output_LLRs, q.output_LLRs[-1]))
    assert np.array_equal(llr, ref_llr)
    # empirical distribution ~ p(t | x=0)
    big = q.quantize_direct_OpenCL(2000, 500).get()
    emp = np.bincount(big.ravel(), minlength=16) / big.size
    assert np.allclose(emp, np.diff(q.cdf_t_given_x_equals_zero), atol=3e-3)


def test_error_counters(gpu):
    import torch
    from informationbottleneckdecodingldpc_b200.engine import count_errors
    rng = np.random.Generator(np.random.PCG64(8))
    out = rng.integers(0, 16, size=(333, 77)).astype(np.uint8)
    out[:, 5] = 15
    bits = rng.integers(0, 2, size=(333, 77)).astype(np.uint8)
    for rows in (333, 100, 1):
        b, f = count_errors(torch.from_numpy(out).cuda(), rows, 8)
        assert b == int((out[:rows] < 8).sum()) and f == int(((out[:rows] < 8).sum(0) > 0).sum())
        b, f = count_errors(torch.from_numpy(out).cuda(), rows, 8, torch.from_numpy(bits).cuda())
        e = (out[:rows] < 8) != (bits[:rows] != 0)
        assert b == int(e.sum()) and f == int((e.sum(0) > 0).sum())
    llr = rng.normal(size=(50, 9))
    b, f = count_errors(torch.from_numpy(llr).cuda(), 20, None)
    assert b == int((llr[:20] < 0).sum())


def test_decoder_object_reuse(gpu):
    """One decoder object across what the BER drivers do to it: re-init per Eb/N0 point with another
    batch size, update_trellis_vectors (discrete_LDPC_decoder.py:53), unpinned / int64 / uint8 host
    inputs, alternating device- and host-buffer calls."""
    from oracle import oracle
    H = codes.regular_random(600, 3, 6, seed=8)
    t = graph.edge_tables(H)
    T, imax = 16, 6
    tb1 = luts.random_tables(T, 6, 3, imax, seed=1)
    tb2 = luts.random_tables(T, 6, 3, imax, seed=2)
    dec = _mk_ib(H, T, imax, tb1.Trellis_checknodevector_a, tb1.Trellis_varnodevector_a, irregular=False)
    rng = np.random.Generator(np.random.PCG64(4))
    for B, tb in ((40, tb1), (7, tb1), (530, tb2), (40, tb2), (16, tb1)):
        dec.update_trellis_vectors(tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a)
        dec.init_OpenCL_decoding(B)
        ch = rng.integers(0, T, size=(600, B))
        ref, i_ref = oracle.ib_decode(t, ch, T=T, imax=imax, cn_lut=tb.Trellis_checknodevector_a,
                                      vn_lut=tb.Trellis_varnodevector_a, early=True)
        got_dev = dec.decode_OpenCL(_dev(ch.astype(np.uint8)), buffer_in=True, return_buffer=True).get()
        got_i64 = dec.decode_OpenCL(ch.astype(np.int64), buffer_in=False, return_buffer=False)
        got_f = dec.decode_OpenCL(np.asfortranarray(ch.astype(np.uint8)), buffer_in=False, return_buffer=True).get()
        assert np.array_equal(got_dev, ref) and np.array_equal(got_i64, ref) and np.array_equal(got_f, ref)
        assert dec.last_i_num == i_ref


def test_cabi_argument_errors(gpu):
    import informationbottleneckdecodingldpc_b200 as pkg
    H = codes.regular_random(24, 3, 6, seed=1)
    tb = luts.random_tables(16, 6, 3, 4, seed=1)
    dec = pkg.Discrete_LDPC_Decoder_class(H, 4, 16, 16, tb.Trellis_checknodevector_a[:100], tb.Trellis_varnodevector_a, 2)
    with pytest.raises(RuntimeError, match="too short"):
        dec.init_OpenCL_decoding(2)
    dec = pkg.Discrete_LDPC_Decoder_class(H, 4, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, 2)
    with pytest.raises(ValueError):
        dec.decode_OpenCL(np.full((24, 2), 16, dtype=np.int32))
    with pytest.raises(ValueError):
        dec.decode_OpenCL(np.zeros((23, 2), dtype=np.int32))
