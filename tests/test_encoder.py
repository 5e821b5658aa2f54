"""Transmit side: systematic LDPC encoder, transmitter, AWGN channel.

CPU part: the oracle and the product's host analysis against the golden codewords frozen from the reference's
LDPCEncoder (oracle/make_golden_encoder.py).  GPU part (-m gpu): the CUDA encoder through the class API against the
same goldens and the oracle, size-independent properties at full size, and the encode -> channel -> quantize ->
decode round trip of the reference's ``_enc`` drivers."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ENC_CASES, load_golden
from informationbottleneckdecodingldpc_b200 import codes
from informationbottleneckdecodingldpc_b200.Discrete_LDPC_decoding.LDPC_encoder import (EncoderPlan, gf2_inverse_packed,
                                                                                          is_full_diag_triangular)


def _evaluate_plan(plan, x):
    """numpy evaluation of an EncoderPlan for ONE frame: checks the host analysis (schedule / dense inverse)
    without a GPU.  Test helper only -- the product encodes through ibldpc_encode."""
    x = np.asarray(x).astype(np.uint8).ravel() & 1
    s = np.array([x[plan.a_col[plan.a_rowptr[r]:plan.a_rowptr[r + 1]]].sum() & 1 for r in range(plan.M)], dtype=np.uint8)
    p = np.zeros(plan.M, dtype=np.uint8)
    if plan.method == 1:
        for t in range(plan.M):
            p[plan.var[t]] = (s[plan.eq[t]] + p[plan.oth[plan.oth_ptr[t]:plan.oth_ptr[t + 1]]].sum()) & 1
    else:
        G = np.unpackbits(plan.dense_inverse.view(np.uint8), axis=1, bitorder="little")[:, :plan.M]
        p = (G.astype(np.int64) @ s.astype(np.int64) & 1).astype(np.uint8)
    return np.concatenate([x, p])


def _syndrome_free(H, cw):
    return not (sp.csr_matrix(H).astype(np.int64) @ np.asarray(cw, dtype=np.int64) % 2).any()


# ------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("case", ENC_CASES)
def test_oracle_encoder_matches_reference_codewords(case):
    from oracle import oracle
    g = load_golden(case)
    cw = oracle.encode(g["H"], g["bits"])
    assert np.array_equal(cw, g["codeword"])
    assert _syndrome_free(g["H"], cw) and np.array_equal(cw[:g["bits"].shape[0]], g["bits"])


@pytest.mark.parametrize("case", ENC_CASES)
def test_host_analysis_matches_reference(case):
    """Algorithm classification (LDPC_encoder.py:196-262) and the substitution schedule / dense inverse, evaluated
    in numpy for two frames."""
    g = load_golden(case)
    plan = EncoderPlan(g["H"])
    if int(g["reference_valid"]):
        assert plan.EncodingAlgorithm == str(g["algorithm"])
        assert (plan.RowOrder[0] >= 0) == bool(int(g["row_order_reversed"]))
    else:   # the reference's backward branch: same classification, but only we produce codewords
        assert plan.EncodingAlgorithm == "Backward Substitution" == str(g["algorithm"])
    for b in range(2):
        assert np.array_equal(_evaluate_plan(plan, g["bits"][:, b]), g["codeword"][:, b])
    if plan.method == 1:   # schedule is a permutation and only reads solved bits
        assert sorted(plan.eq) == sorted(plan.var) == list(range(plan.M))
        solved = np.full(plan.M, -1)
        for t in range(plan.M):
            assert all(solved[j] >= 0 for j in plan.oth[plan.oth_ptr[t]:plan.oth_ptr[t + 1]])
            solved[plan.var[t]] = t


def test_gf2_inverse_and_shape_detection():
    rng = np.random.default_rng(3)
    for n in (1, 5, 33, 64, 100):
        while True:
            X = (rng.random((n, n)) < 0.4).astype(np.uint8)
            try:
                G = gf2_inverse_packed(X)
                break
            except ValueError:
                continue
        Gd = np.unpackbits(G.view(np.uint8), axis=1, bitorder="little")[:, :n]
        assert np.array_equal((Gd.astype(np.int64) @ X.astype(np.int64)) % 2, np.eye(n, dtype=np.int64))
    with pytest.raises(ValueError):
        gf2_inverse_packed(np.array([[1, 1], [1, 1]]))
    L = np.tril(np.ones((6, 6), dtype=int))
    assert is_full_diag_triangular(sp.csr_matrix(L)) == 1
    assert is_full_diag_triangular(sp.csr_matrix(L.T)) == -1
    assert is_full_diag_triangular(sp.csr_matrix(np.eye(4, dtype=int))) == 1
    assert is_full_diag_triangular(sp.csr_matrix(L[::-1])) == 0
    with pytest.raises(ValueError):
        EncoderPlan(np.ones((4, 3)))
    with pytest.raises(ValueError):       # singular last part
        EncoderPlan(np.array([[1, 0, 1, 1], [0, 1, 1, 1]]))


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", ENC_CASES)
def test_gpu_encoder_golden(gpu, case):
    import informationbottleneckdecodingldpc_b200 as pkg
    g = load_golden(case)
    enc = pkg.LDPCEncoder(g["H"])
    out = enc.encode_batch(g["bits"])
    assert out.shape == g["codeword"].shape and np.array_equal(out.get(), g["codeword"])
    one = enc.encode_c(g["bits"][:, 1].astype(np.int32))
    assert one.shape == (g["H"].shape[1],) and np.array_equal(one, g["codeword"][:, 1])
    assert np.array_equal(enc.encode(g["bits"][:, 0]), g["codeword"][:, 0])
    assert enc.NumInfoBits + enc.NumParityBits == enc.BlockLength == g["H"].shape[1]


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 31, 33, 128, 1000, 4097])
def test_gpu_encoder_ragged_batches_vs_oracle(gpu, B):
    import informationbottleneckdecodingldpc_b200 as pkg
    from oracle import oracle
    for H in (codes.wlan_80211n(54), codes.dvbs2_like_half_rate(6480, q_groups=36)):
        K = H.shape[1] - H.shape[0]
        bits = np.random.default_rng(B).integers(0, 2, size=(K, B)).astype(np.uint8)
        enc = pkg.LDPCEncoder(H)
        got = enc.encode_batch(bits).get()
        assert np.array_equal(got, oracle.encode(H, bits))


@pytest.mark.gpu
def test_gpu_encoder_full_size_properties(gpu):
    """DVB-S2-like n=64800 (32400-step substitution) and 802.11n n=1944, thousands of frames: every codeword
    satisfies H c = 0, is systematic, encoding is linear (c(a^b) = c(a)^c(b)) and frame-independent."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    for H, B in ((codes.dvbs2_like_half_rate(), 2048), (codes.wlan_80211n(81), 20000)):
        K = H.shape[1] - H.shape[0]
        tr = pkg.LDPC_BPSK_Transmitter(H, B)
        tr.return_buffer_only = True
        assert tr.data_len == K and tr.codeword_len == H.shape[1]
        a, b = tr.random_bits(), tr.random_bits()
        assert 0.49 < float(a.float().mean()) < 0.51 and not torch.equal(a, b)
        ca, cb, cab = (tr.encoder.encode_batch(x).tensor for x in (a, b, a ^ b))
        assert torch.equal(ca[:K], a) and torch.equal(ca ^ cb, cab)
        Hs = torch.sparse_csr_tensor(torch.from_numpy(H.indptr.astype(np.int64)), torch.from_numpy(H.indices.astype(np.int64)),
                                     torch.ones(H.nnz, dtype=torch.float32), size=H.shape).cuda()
        assert int((Hs @ ca.float()).remainder(2).sum()) == 0
        sub = tr.encoder.encode_batch(a[:, 100:177].contiguous()).tensor
        assert torch.equal(sub, ca[:, 100:177])


@pytest.mark.gpu
def test_gpu_channel_and_bits_are_deterministic_and_gaussian(gpu):
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    ch = pkg.AWGN_channel(0.5)
    x = torch.zeros((1000, 2000), dtype=torch.float64, device="cuda")
    n1 = ch.transmission(x).tensor
    assert abs(float(n1.mean())) < 5e-3 and abs(float(n1.var()) - 0.5) < 5e-3
    assert abs(float((n1 ** 4).mean()) / 0.25 - 3.0) < 0.05            # Gaussian kurtosis
    ch2 = pkg.AWGN_channel(0.5)
    assert torch.equal(ch2.transmission(x).tensor, n1)                 # same seed and offset -> same noise
    assert not torch.equal(ch2.transmission(x).tensor, n1)             # the offset advances
    bits = (torch.rand((64, 500), device="cuda") < 0.5).to(torch.uint8)
    ca, cb = pkg.AWGN_channel(0.1), pkg.AWGN_channel(0.1)
    y1 = ca.transmission_bits(bits).tensor
    y2 = cb.transmission(pkg.DeviceArray(1.0 - 2.0 * bits.double())).tensor
    assert torch.equal(y1, y2)
    host = pkg.AWGN_channel(0.1).transmission(np.ones((64, 500)))
    assert isinstance(host, np.ndarray) and abs(host.mean() - 1.0) < 0.01


@pytest.mark.gpu
def test_enc_driver_round_trip(gpu):
    """The loop of Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL_enc.py:118-134 on the device: transmit ->
    AWGN channel -> quantize -> IB decode with message alignment -> compare with last_transmitted_bits.  At
    3.5 dB every frame must come back exactly; at 0 dB there must be errors (the comparison is not vacuous)."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_irregular_config
    from informationbottleneckdecodingldpc_b200.engine import count_errors
    H = codes.wlan_80211n(54)
    B, T = 512, 16
    cfg, _ = generate_irregular_config(1.0, H, T, 50)
    transi = pkg.LDPC_BPSK_Transmitter(H, B)
    transi.return_buffer_only = True
    decodi = pkg.Discrete_LDPC_Decoder_class_irregular(H, 50, T, T, cfg.Trellis_checknodevector_a, cfg.Trellis_varnodevector_a,
                                                       cfg.matching_vector_checknode, cfg.matching_vector_varnode, B)
    errs = {}
    for ebn0 in (3.5, 0.0):
        sigma_n2 = 10 ** (-ebn0 / 10) / (2 * transi.R_c)
        chani = pkg.AWGN_channel(sigma_n2)
        quanti = pkg.AWGN_Channel_Quantizer(sigma_n2, 3, T, 2000)
        quanti.init_OpenCL_quanti(H.shape[1], B, return_buffer_only=True)
        decodi.init_OpenCL_decoding(B, quanti.context)
        send = transi.transmit()
        rec = quanti.quantize_OpenCL(chani.transmission(send))
        out = decodi.decode_OpenCL(rec, buffer_in=True, return_buffer=True)
        sent = transi.last_transmitted_bits.tensor
        hard = (out.tensor[:transi.data_len] < T // 2).to(torch.uint8)
        errs[ebn0] = int((hard != sent).sum())
        bit, frame = count_errors(out, transi.data_len, T // 2, ref_bits=sent)
        assert bit == errs[ebn0]
        # the fused BPSK + noise kernel gives the same received values as mapping + channel
        c1, c2 = pkg.AWGN_channel(sigma_n2), pkg.AWGN_channel(sigma_n2)
        coded = transi.transmit_bits()
        assert torch.equal(c1.transmission_bits(coded).tensor, c2.transmission(transi.BPSK_mapping(coded)).tensor)
    assert errs[3.5] == 0 and errs[0.0] > 0
