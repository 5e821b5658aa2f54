"""GPU suite, round 2 additions: packed / int32 host contracts, lazy i_num, input-range reporting, asynchronous LLR
error counters, the C-ABI counter all-reduce, the pipelined multi-GPU BER loop and the timed bench geometry."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from informationbottleneckdecodingldpc_b200 import codes, graph, luts

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_ib(t, ch, T, imax, tb, early):
    from oracle import oracle
    return oracle.ib_decode(t, ch, T=T, imax=imax, cn_lut=tb.Trellis_checknodevector_a, vn_lut=tb.Trellis_varnodevector_a,
                            cn_match=tb.matching_vector_checknode, vn_match=tb.matching_vector_varnode, early=early)


def _wlan_decoder(T=16, imax=6, B=64, seed=3):
    import informationbottleneckdecodingldpc_b200 as pkg
    H = codes.wlan_80211n(54)
    t = graph.edge_tables(H)
    tb = luts.random_tables(T, t.d_c_max, t.d_v_max, imax, seed=seed, matching=True)
    dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                    tb.matching_vector_checknode, tb.matching_vector_varnode, B)
    dec.init_OpenCL_decoding(B)
    return H, t, tb, dec


@pytest.mark.parametrize("B,chunk", [(100, 0), (777, 0), (9001, 0), (5000, 1008), (64, 0)])
def test_packed_host_contract_matches_oracle(gpu, B, chunk):
    """nibble-packed channel values in, bit-packed hard decisions of the first data_len rows out."""
    from informationbottleneckdecodingldpc_b200 import _lib
    H, t, tb, dec = _wlan_decoder(B=B)
    dec.early_termination = False
    rng = np.random.Generator(np.random.PCG64(B))
    ch = rng.integers(0, 16, size=(t.n_var, B)).astype(np.uint8)
    if chunk:
        _lib.check(_lib.lib().ibldpc_set_host_chunk(dec._ensure_handle(), chunk))
    packed = dec.pack_channel_values(ch)
    assert packed.shape == (t.n_var, (B + 1) // 2)
    bits = dec.unpack_bits(dec.decode_packed(packed, B), B)
    sel = np.r_[0:min(B, 8), max(B - 8, 0):B]
    ref, _ = _oracle_ib(t, np.ascontiguousarray(ch[:, sel]), 16, 6, tb, False)
    rows = int(dec.data_len)
    assert bits.shape == (rows, B)
    assert np.array_equal(bits[:, sel], (ref[:rows] < 8).astype(np.uint8))
    # and against the device-buffer path on every frame
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    out = dec.decode_OpenCL(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True).get()
    assert np.array_equal(bits, (out[:rows] < 8).astype(np.uint8))


def test_packed_host_contract_early_termination_and_uint8_family(gpu, monkeypatch):
    H, t, tb, dec = _wlan_decoder(B=48)
    tbm = luts.minsum_like_tables(16, t.d_c_max, t.d_v_max, 6)
    dec.update_trellis_vectors(tbm.Trellis_checknodevector_a, tbm.Trellis_varnodevector_a)
    dec.matching_vector_checknode, dec.matching_vector_varnode = tbm.matching_vector_checknode, tbm.matching_vector_varnode
    ch = np.full((t.n_var, 48), 15, dtype=np.uint8)       # strongly "bit 0": converges at once
    bits = dec.unpack_bits(dec.decode_packed(dec.pack_channel_values(ch), 48), 48)
    ref, inum = _oracle_ib(t, ch, 16, 6, tbm, True)
    assert dec.last_i_num == inum and not bits.any()
    monkeypatch.setenv("IBLDPC_NO_NIBBLE", "1")            # uint8 family: the library unpacks on the device
    H, t, tb, dec2 = _wlan_decoder(B=48)
    rng = np.random.Generator(np.random.PCG64(9))
    ch = rng.integers(0, 16, size=(t.n_var, 48)).astype(np.uint8)
    bits = dec2.unpack_bits(dec2.decode_packed(dec2.pack_channel_values(ch), 48), 48)
    ref, _ = _oracle_ib(t, ch, 16, 6, tb, True)
    assert dec2.info()[0] == 1
    assert np.array_equal(bits, (ref[:int(dec2.data_len)] < 8).astype(np.uint8))


@pytest.mark.parametrize("B", [1, 33, 1000, 9001])
def test_int32_host_contract(gpu, B):
    """The reference's own contract: int numpy in -> int32 numpy out (discrete_LDPC_decoder.py:207-209, :292-295)."""
    H, t, tb, dec = _wlan_decoder(B=B)
    dec.early_termination = True
    rng = np.random.Generator(np.random.PCG64(B))
    ch = rng.integers(0, 16, size=(t.n_var, B)).astype(np.int32)
    out = dec.decode_OpenCL(ch)
    assert out.dtype == np.int32 and out.shape == ch.shape
    sel = np.arange(B) if B <= 64 else np.r_[0:8, B - 8:B]
    if B <= 64:
        ref, inum = _oracle_ib(t, ch, 16, 6, tb, True)
        assert dec.last_i_num == inum
        assert np.array_equal(out, ref)
    else:
        dec.early_termination = False
        out = dec.decode_OpenCL(ch.astype(np.int64))        # int64 (quantize_on_host output) takes the same path
        ref, _ = _oracle_ib(t, np.ascontiguousarray(ch[:, sel]), 16, 6, tb, False)
        assert np.array_equal(out[:, sel], ref)
    bad = ch.copy()
    bad[5, B // 2] = 16
    with pytest.raises(ValueError):
        dec.decode_OpenCL(bad)
    bad[5, B // 2] = -1
    with pytest.raises(ValueError):
        dec.decode_OpenCL(bad)


def test_out_of_range_cluster_indices_are_reported_not_dereferenced(gpu, monkeypatch):
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    for env in ({}, {"IBLDPC_NO_NIBBLE": "1"}, {"IBLDPC_FORCE_GENERIC": "1"}):
        for k in ("IBLDPC_NO_NIBBLE", "IBLDPC_FORCE_GENERIC"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        H, t, tb, dec = _wlan_decoder(B=40)
        ch = np.zeros((t.n_var, 40), dtype=np.uint8)
        ch[7, 3] = 200
        with pytest.raises(ValueError):
            dec.decode_OpenCL(ch)                                        # host uint8 path
        out = dec.decode_OpenCL(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True)
        with pytest.raises(RuntimeError):
            dec.last_i_num                                               # device path: reported at the lazy read-back
        torch.cuda.synchronize()                                         # and the context is still healthy
        ch[7, 3] = 1
        dec.decode_OpenCL(ch)
        with pytest.raises(ValueError):                                  # int32 device buffers are checked before the cast
            dec.decode_OpenCL(pkg.DeviceArray(torch.full((t.n_var, 8), 300, dtype=torch.int32, device="cuda")), buffer_in=True)
        del out


def test_device_buffer_decode_is_asynchronous_and_i_num_is_lazy(gpu):
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    H = codes.regular_random(8000, 3, 6)
    t = graph.edge_tables(H)
    tb = luts.minsum_like_tables(16, 6, 3, 50)
    dec = pkg.Discrete_LDPC_Decoder_class(H, 50, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, 8192)
    dec.init_OpenCL_decoding(8192)
    rng = np.random.Generator(np.random.PCG64(1))
    ch = pkg.DeviceArray(torch.from_numpy(rng.integers(0, 16, size=(8000, 8192)).astype(np.uint8)).cuda())
    dec.decode_OpenCL(ch, buffer_in=True, return_buffer=True)            # warm-up (allocations)
    torch.cuda.synchronize()
    ev = torch.cuda.Event()
    dec.decode_OpenCL(ch, buffer_in=True, return_buffer=True)            # early termination on (class default)
    ev.record()
    assert not ev.query(), "decode_OpenCL(buffer_in=True, return_buffer=True) must not synchronise the stream"
    assert dec._inum_pending
    assert 2 <= dec.last_i_num <= 50 and not dec._inum_pending


def test_llr_async_error_counters(gpu):
    import torch
    from informationbottleneckdecodingldpc_b200.engine import count_errors, count_errors_async
    rng = np.random.Generator(np.random.PCG64(2))
    for dt in (np.float32, np.float64):
        x = rng.normal(1.0, 1.0, size=(100, 333)).astype(dt)
        t = torch.from_numpy(x).cuda()
        c = torch.tensor([5, 7, 0, 0], dtype=torch.int64, device="cuda")
        count_errors_async(t, 60, 0, c)
        count_errors_async(t, 60, 0, c)
        bit, frame = int((x[:60] < 0).sum()), int((x[:60] < 0).any(axis=0).sum())
        assert c.tolist() == [5 + 2 * bit, 7 + 2 * frame, 0, 0]
        assert count_errors(t, 60) == (bit, frame)


def test_cabi_counter_allreduce_single_rank(gpu):
    """ibldpc_nccl_unique_id / _init / ibldpc_allreduce_counters with a one-rank communicator (the 2-rank case is
    covered by test_ber_point_two_ranks when two GPUs are visible)."""
    import torch
    from informationbottleneckdecodingldpc_b200 import _lib
    from informationbottleneckdecodingldpc_b200.parallel import nccl_library_path
    p = nccl_library_path()
    if p:
        os.environ.setdefault("IBLDPC_NCCL_LIB", p)
    H, t, tb, dec = _wlan_decoder(B=16)
    L = _lib.lib()
    ident = (C.c_uint8 * 128)()
    _lib.check(L.ibldpc_nccl_unique_id(ident))
    assert any(ident)
    h = dec._ensure_handle()
    c = torch.tensor([3, 1, 16, 96], dtype=torch.int64, device="cuda")
    assert L.ibldpc_allreduce_counters(h, C.c_void_p(c.data_ptr()), 4, None) < 0      # before init: state error
    _lib.check(L.ibldpc_nccl_init(h, ident, 0, 1))
    _lib.check(L.ibldpc_allreduce_counters(h, C.c_void_p(c.data_ptr()), 4, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert c.tolist() == [3, 1, 16, 96]
    _lib.check(L.ibldpc_nccl_finalize(h))


def test_philox_substreams_differ_on_the_device(gpu):
    import informationbottleneckdecodingldpc_b200 as pkg
    qs = []
    for stream in (0, 1, 2):
        q = pkg.AWGN_Channel_Quantizer(0.8, 3, 16, 2000)
        q.set_stream(stream)
        q.init_OpenCL_quanti(64, 256, return_buffer_only=True)
        qs.append(q.quantize_direct_OpenCL(64, 256).get())
    assert not np.array_equal(qs[0], qs[1]) and not np.array_equal(qs[1], qs[2])
    assert abs(float((qs[0] == qs[1]).mean()) - float((qs[0] == np.roll(qs[0], 1, axis=1)).mean())) < 0.03   # independent draws
    q = pkg.AWGN_Channel_Quantizer(0.8, 3, 16, 2000)
    q.init_OpenCL_quanti(64, 256, return_buffer_only=True)           # no process group: stream 0 = the seed itself
    assert np.array_equal(q.quantize_direct_OpenCL(64, 256).get(), qs[0])


def _ber_setup(llr):
    import informationbottleneckdecodingldpc_b200 as pkg
    from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_irregular_config
    H = codes.wlan_80211n(54)
    if llr:
        dec = pkg.Min_Sum_Decoder_class_irregular(H, 20, 16, 512)
    else:
        tb, _ = generate_irregular_config(1.0, H, 16, 20)
        dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, 20, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                        tb.matching_vector_checknode, tb.matching_vector_varnode, 512)
    q = pkg.AWGN_Channel_Quantizer(10 ** (-1.5 / 10) / (2 * 0.5), 3, 16, 2000)
    q.init_OpenCL_quanti(1296, 512, return_buffer_only=True)
    dec.init_OpenCL_decoding(512, q.context)
    return dec, q


@pytest.mark.parametrize("llr", [False, True])
def test_ber_point_equals_the_synchronous_reference_loop(gpu, llr):
    """simulation.ber_point (pipelined, counters one batch late) against the plain loop of the reference drivers
    (WLAN/BER_simulation_OpenCL.py:124-131) on the same Philox stream: the pipelined loop decodes the same batches
    plus at most two more."""
    from informationbottleneckdecodingldpc_b200.simulation import ber_point
    dec, q = _ber_setup(llr)
    min_errors = 400
    res = ber_point(dec, q, 512, min_errors=min_errors, llr=llr)
    q.set_stream(0)                                       # rewind the same sub-stream
    errors, fer, batches, per_batch = 0, 0, 0, []
    while batches * 512 < res["frames"]:
        rec = q.quantize_direct_OpenCL_LLR(1296, 512) if llr else q.quantize_direct_OpenCL(1296, 512)
        out = dec.decode(rec, buffer_in=True, return_buffer=True)
        b, f = dec.count_errors(out)
        assert b == dec.return_errors_all_zero(out)
        errors += b
        fer += f
        batches += 1
        per_batch.append(b)
    assert (res["bit_errors"], res["frame_errors"], res["frames"]) == (errors, fer, batches * 512)
    need = next(i + 1 for i in range(batches) if sum(per_batch[:i + 1]) >= min_errors)
    assert need <= batches <= need + 2, (need, batches)
    assert abs(res["ber"] - errors / (batches * 512 * int(dec.data_len))) < 1e-12


def test_ber_point_two_ranks(gpu):
    """2-rank torchrun: disjoint sub-streams per rank, NCCL all-reduced counters (once through torch.distributed, once
    through the library's own communicator), identical totals on both ranks."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = os.path.join(ROOT, "tests", "_ber_two_ranks.py")
    for abi in ("0", "1"):
        env = dict(os.environ, IBLDPC_ABI_ALLREDUCE=abi)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                            "--master-port", "29631", script], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "TWO_RANK_OK" in r.stdout


def test_timed_bench_geometry_is_bit_exact(gpu):
    """The launch plan bench.py times (C1, B = 65536: 1024-thread CTAs, 37 x 4 / 74 x 2 grids) against the oracle on 32
    columns, and frame independence across the whole batch (every column equals its copy at another position)."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_regular_config
    H = codes.regular_random(8000, 3, 6)
    t = graph.edge_tables(H)
    tb, _ = generate_regular_config(1.2, 3, 6, 16, 50)
    B = 65536
    dec = pkg.Discrete_LDPC_Decoder_class(H, 50, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, B)
    dec.init_OpenCL_decoding(B)
    dec.early_termination = False
    q = pkg.AWGN_Channel_Quantizer(10 ** (-1.6 / 10) / (2 * 0.5), 3, 16, 2000)
    q.init_OpenCL_quanti(8000, B, return_buffer_only=True)
    ch = q.quantize_direct_OpenCL(8000, B)
    # plant copies: columns of the second half repeat those of the first half in reversed order
    ch.tensor[:, B // 2:] = ch.tensor[:, :B // 2].flip(1)
    out = dec.decode_OpenCL(ch, buffer_in=True, return_buffer=True)
    assert torch.equal(out.tensor[:, B // 2:], out.tensor[:, :B // 2].flip(1))
    cols = np.r_[0:8, 8191:8199, 32760:32776]
    ref, _ = _oracle_ib(t, ch.tensor[:, torch.from_numpy(cols).cuda()].cpu().numpy(), 16, 50, tb, False)
    assert np.array_equal(out.tensor[:, torch.from_numpy(cols).cuda()].cpu().numpy(), ref.astype(np.uint8))


@pytest.mark.parametrize("code", ["wlan1296", "dvb6480", "reg36"])
@pytest.mark.parametrize("B", [1, 77, 513, 1040, 4097, 20011])
@pytest.mark.parametrize("match", [True, False])
def test_fused_phase_kernels_vs_oracle(gpu, monkeypatch, code, B, match):
    """ib_phase_n4.cuh: one launch per phase over all degree classes (TMA-staged table image, dynamic item
    distribution) against the oracle for ragged batch sizes -- one tile, partial tiles, more items than warps."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    monkeypatch.setenv("IBLDPC_COOP_MAX_B", "0")
    monkeypatch.setenv("IBLDPC_PHASE", "1")          # also for the sets where the per-class launches stay the default
    H = {"wlan1296": lambda: codes.wlan_80211n(54), "dvb6480": lambda: codes.dvbs2_like_half_rate(6480, q_groups=36),
         "reg36": lambda: codes.regular_random(2000, 3, 6, seed=5)}[code]()
    if code == "reg36" and match:
        pytest.skip("regular decoder has no message alignment")
    t = graph.edge_tables(H)
    T, imax = 16, 5
    tb = luts.random_tables(T, t.d_c_max, t.d_v_max, imax, seed=B, matching=match)
    rng = np.random.Generator(np.random.PCG64(B + 1))
    ch = rng.integers(0, T, size=(t.n_var, B)).astype(np.uint8)
    if code == "reg36":
        dec = pkg.Discrete_LDPC_Decoder_class(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, B)
    else:
        dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                        tb.matching_vector_checknode, tb.matching_vector_varnode, B,
                                                        match='true' if match else 'false')
    dec.init_OpenCL_decoding(B)
    out = dec.decode_OpenCL(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True).get()
    assert dec.info()[0] == 2 and dec.info()[1] == 1 + 2 * imax, dec.info()      # pack + 1 + 2 (imax - 1) + 1 launches
    sel = np.arange(B) if B <= 1100 else np.r_[0:24, B // 2:B // 2 + 16, B - 24:B]
    ref, i_num = _oracle_ib(t, np.ascontiguousarray(ch[:, sel]), T, imax, tb, B <= 1100)
    assert np.array_equal(out[:, sel], ref)
    if B <= 1100:
        assert dec.last_i_num == i_num


def test_fused_phase_kernels_early_termination(gpu, monkeypatch):
    """Tables with decoding power and mostly reliable channel values: the batch converges after 9 passes and the fused
    kernels must stop exactly where the reference's batch-granular rule stops (i_num and outputs equal to the oracle's)."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    monkeypatch.setenv("IBLDPC_COOP_MAX_B", "0")
    H = codes.wlan_80211n(54)
    t = graph.edge_tables(H)
    tb = luts.minsum_like_tables(16, t.d_c_max, t.d_v_max, 12)
    dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, 12, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                    tb.matching_vector_checknode, tb.matching_vector_varnode, 300)
    dec.init_OpenCL_decoding(300)
    rng = np.random.Generator(np.random.PCG64(3))
    ch = np.where(rng.random((t.n_var, 300)) < 0.97, rng.integers(10, 16, size=(t.n_var, 300)),
                  rng.integers(5, 8, size=(t.n_var, 300))).astype(np.uint8)
    out = dec.decode_OpenCL(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True).get()
    ref, i_num = _oracle_ib(t, ch, 16, 12, tb, True)
    assert dec.info()[1] == 1 + 2 * 12          # fused kernels (every launch is issued; converged passes return at once)
    assert np.array_equal(out, ref) and dec.last_i_num == i_num and 2 < i_num < 12, (dec.last_i_num, i_num)


# ---------------------------------------------------------------------------------- |T| <= 32 family (ib_kernels_t32.cuh)
@pytest.mark.parametrize("T", [32, 24, 18])
@pytest.mark.parametrize("match", [True, False])
def test_t32_family_all_degrees(gpu, T, match):
    """Every instantiated degree (checks 3..10, variables 1..12) of the |T| <= 32 family against the oracle, ragged batch."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    deg_v = [1, 1] + [d for d in range(2, 13) for _ in range(6)]
    E = sum(deg_v)
    deg_c = [3] * 9 + [4] * 8 + [5] * 8 + [6] * 8 + [7] * 6 + [8] * 6 + [9] * 4 + [10] * 4
    deg_c += [4] * ((E - sum(deg_c)) // 4)
    rest = E - sum(deg_c)
    assert 0 <= rest < 4
    if rest:
        deg_c[0:rest] = [d + 1 for d in deg_c[0:rest]]
    H = codes.random_from_degrees(deg_v, deg_c, seed=8)
    t = graph.edge_tables(H)
    assert set(t.degree_chk) >= set(range(3, 11)) and set(t.degree_var) == set(range(1, 13))
    imax, B = 5, 333
    tb = luts.random_tables(T, t.d_c_max, t.d_v_max, imax, seed=T, matching=match)
    ch = np.random.Generator(np.random.PCG64(T)).integers(0, T, size=(t.n_var, B)).astype(np.uint8)
    dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                    tb.matching_vector_checknode, tb.matching_vector_varnode, B,
                                                    match='true' if match else 'false')
    dec.init_OpenCL_decoding(B)
    out = dec.decode_OpenCL(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True).get()
    assert dec.info()[0] == 3
    ref, i_num = _oracle_ib(t, ch, T, imax, tb, True)
    assert np.array_equal(out, ref) and dec.last_i_num == i_num


@pytest.mark.parametrize("group", [0, 1, 2])
@pytest.mark.parametrize("T", [32, 24])
def test_t32_fused_phase_kernels_all_degrees(gpu, monkeypatch, group, T):
    """ib_t32_coop_kernel (whole decode in one cooperative launch, small batches) and ib_t32_phase_kernel (one launch per
    phase; the degree classes one after the other inside every CTA, image reloaded by TMA between the classes) for degree sets
    of four classes covering every degree -- checks 3..10, variables 1..12 -- against the oracle (outputs, i_num with the
    batch-granular stop) and against one launch per class (IBLDPC_T32_NO_PHASE=1)."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    vset = ([1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12])[group]
    cset = ([3, 4, 5, 6], [7, 8, 9, 10], [4, 6, 8, 10])[group]
    deg_v = [d for d in vset for _ in range(24)]
    E = sum(deg_v)
    deg_c = [d for d in cset for _ in range(6)]
    while sum(deg_c) + cset[0] <= E:
        deg_c.append(cset[len(deg_c) % 4] if sum(deg_c) + cset[len(deg_c) % 4] <= E else cset[0])
    rest = E - sum(deg_c)
    # spread what is left over checks that stay inside the set's degrees: bump checks to the next degree of the set
    i = 0
    while rest > 0 and i < len(deg_c):
        nxt = [d for d in cset if d > deg_c[i]]
        if nxt and nxt[0] - deg_c[i] <= rest:
            rest -= nxt[0] - deg_c[i]
            deg_c[i] = nxt[0]
        i += 1
    if rest:
        pytest.skip("degree bookkeeping left a remainder")
    H = codes.random_from_degrees(deg_v, deg_c, seed=80 + group)
    t = graph.edge_tables(H)
    assert len(set(t.degree_chk)) <= 4 and set(t.degree_var) == set(vset)
    imax = 6
    tb = luts.random_tables(T, t.d_c_max, t.d_v_max, imax, seed=T + group, matching=True)
    for B in (77, 1500):
        ch = np.random.Generator(np.random.PCG64(T + B)).integers(0, T, size=(t.n_var, B)).astype(np.uint8)

        def run():
            dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                            tb.matching_vector_checknode, tb.matching_vector_varnode, B)
            dec.init_OpenCL_decoding(B)
            out = dec.decode_OpenCL(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True).get()
            return dec, out

        dec, out = run()
        assert dec.info()[0] == 3 and dec.info()[1] <= 2 * imax + 2, dec.info()      # one launch per phase (+ clamp / pad kernels)
        sel = np.arange(B) if B <= 100 else np.r_[0:16, B - 16:B]
        ref, i_num = _oracle_ib(t, np.ascontiguousarray(ch[:, sel]), T, imax, tb, B <= 100)
        assert np.array_equal(out[:, sel], ref)
        if B <= 100:
            assert dec.last_i_num == i_num
        assert dec.info()[1] <= 3, dec.info()                 # small batch: the whole decode in one cooperative launch
        monkeypatch.setenv("IBLDPC_T32_COOP_MAX_B", "0")      # one launch per phase
        dec1, out1 = run()
        monkeypatch.delenv("IBLDPC_T32_COOP_MAX_B")
        assert 2 * imax <= dec1.info()[1] <= 2 * imax + 2 and np.array_equal(out1, out) and dec1.last_i_num == dec.last_i_num
        monkeypatch.setenv("IBLDPC_T32_NO_PHASE", "1")         # one launch per phase and degree class
        dec2, out2 = run()
        monkeypatch.delenv("IBLDPC_T32_NO_PHASE")
        assert dec2.info()[1] > dec1.info()[1] and np.array_equal(out2, out) and dec2.last_i_num == dec.last_i_num


@pytest.mark.parametrize("B", [1, 100, 2049, 20011])
def test_t32_family_wlan_vs_oracle_and_generic(gpu, monkeypatch, B):
    """802.11n with the reference's cardinality 32: oracle on a sample of frames, the generic path on all of them;
    host-buffer contract (int numpy in / out) through the same kernels."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    H = codes.wlan_80211n(54)
    t = graph.edge_tables(H)
    T, imax = 32, 6
    tb = luts.random_tables(T, t.d_c_max, t.d_v_max, imax, seed=B, matching=True)
    ch = np.random.Generator(np.random.PCG64(B)).integers(0, T, size=(t.n_var, B)).astype(np.uint8)

    def run():
        dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                        tb.matching_vector_checknode, tb.matching_vector_varnode, B)
        dec.init_OpenCL_decoding(B)
        dec.early_termination = False
        return dec, dec.decode_OpenCL(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True).get()

    dec, out = run()
    assert dec.info()[0] == 3
    sel = np.arange(B) if B <= 100 else np.r_[0:12, B - 12:B]
    ref, _ = _oracle_ib(t, np.ascontiguousarray(ch[:, sel]), T, imax, tb, False)
    assert np.array_equal(out[:, sel], ref)
    host = dec.decode_OpenCL(ch.astype(np.int32))
    assert host.dtype == np.int32 and np.array_equal(host, out)
    monkeypatch.setenv("IBLDPC_NO_T32", "1")
    dec_g, out_g = run()
    assert dec_g.info()[0] == 0 and np.array_equal(out_g, out)
    bad = ch.copy()
    bad[3, 0] = 40
    with pytest.raises(ValueError):
        dec.decode_OpenCL(bad)


# ---------------------------------------------------------------------------------- per-frame early termination (ib_perframe.cu)
def _per_frame_oracle(t, ch, T, imax, tb):
    """The reference result of decoding every frame on its own (msg_at_time = 1)."""
    outs, inums = [], []
    for f in range(ch.shape[1]):
        o, i = _oracle_ib(t, np.ascontiguousarray(ch[:, f:f + 1]), T, imax, tb, True)
        outs.append(o[:, 0])
        inums.append(i)
    return np.stack(outs, axis=1), np.asarray(inums)


@pytest.mark.parametrize("code,B,ebn0", [("wlan1296", 300, 2.2), ("wlan1296", 4099, 2.2), ("wlan1296", 1, 2.2), ("reg36", 333, 2.0),
                                         ("reg36", 2500, 1.8), ("reg36_notriple", 333, 2.0), ("dvb6480", 130, 1.6)])
def test_per_frame_early_termination_equals_single_frame_reference(gpu, code, B, ebn0, monkeypatch):
    """early_termination='frame': outputs AND per-frame i_num equal to the oracle run with one frame per call, with
    frames converging at different passes (designed tables, channel draws near the waterfall), frames that never
    converge, and several compactions on the way.  The (3,6) set runs on the phase images with the three-input tables
    (ib_phase_reg36_tri.cu); reg36_notriple = IBLDPC_NO_PF_TRIPLE=1, the plain tail-pair images."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_irregular_config, generate_regular_config
    T, imax = 16, 25
    if code == "reg36_notriple":
        monkeypatch.setenv("IBLDPC_NO_PF_TRIPLE", "1")
        code = "reg36"
    if code == "reg36":
        H = codes.regular_random(2000, 3, 6, seed=5)
        tb, _ = generate_regular_config(1.2, 3, 6, T, imax)
        dec = pkg.Discrete_LDPC_Decoder_class(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, B)
    else:
        H = codes.wlan_80211n(54) if code == "wlan1296" else codes.dvbs2_like_half_rate(6480, q_groups=36)
        tb, _ = generate_irregular_config(1.0, H, T, imax)
        dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                        tb.matching_vector_checknode, tb.matching_vector_varnode, B)
    t = graph.edge_tables(H)
    dec.init_OpenCL_decoding(B)
    q = pkg.AWGN_Channel_Quantizer(10 ** (-ebn0 / 10) / (2 * 0.5), 3, T, 2000)
    u = np.random.Generator(np.random.PCG64(B)).random(size=(t.n_var, B))
    ch = (((u[:, :, None] - q.cdf_t_given_x_equals_zero) > 0).sum(2) - 1).astype(np.uint8)
    dec.early_termination = 'frame'
    out = dec.decode_OpenCL(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True).get()
    inum = dec.last_i_num_per_frame.get()
    sel = np.arange(B) if B <= 400 else np.r_[0:100, B // 2:B // 2 + 100, B - 100:B]
    ref, ref_inum = _per_frame_oracle(t, ch[:, sel], T, imax, tb)
    assert np.array_equal(inum[sel], ref_inum), (inum[sel][:20], ref_inum[:20])
    assert np.array_equal(out[:, sel], ref)
    assert dec.last_i_num == int(inum.max())
    if code == "wlan1296" and B >= 300:
        assert len(set(ref_inum.tolist())) >= 6 and ref_inum.max() == imax     # a real spread, incl. frames that never converge
    # host-array contract of the same mode
    if B <= 400:
        host = dec.decode_OpenCL(ch.astype(np.int32))
        assert host.dtype == np.int32 and np.array_equal(host, out)
    # and the batch-granular default is untouched
    dec.early_termination = True
    out_b = dec.decode_OpenCL(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True).get()
    assert dec.last_i_num == int(ref_inum.max()) or B > 400
    conv = ref_inum[: len(sel)] == dec.last_i_num
    assert np.array_equal(out_b[:, sel][:, conv], ref[:, conv])


def test_per_frame_early_termination_needs_an_instantiated_degree_set(gpu):
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    H = codes.random_from_degrees([2] * 40 + [3] * 40 + [4] * 16, [4] * 16 + [5] * 8 + [6] * 8 + [7] * 6 + [8] * 4 + [9] * 2 + [10] * 2, seed=4)
    t = graph.edge_tables(H)
    tb = luts.random_tables(16, t.d_c_max, t.d_v_max, 4, seed=1, matching=True)
    dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, 4, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                    tb.matching_vector_checknode, tb.matching_vector_varnode, 8)
    dec.early_termination = 'frame'
    with pytest.raises(RuntimeError, match="degree set"):
        dec.decode_OpenCL(pkg.DeviceArray(torch.zeros((t.n_var, 8), dtype=torch.uint8, device="cuda")), buffer_in=True, return_buffer=True)


# ---------------------------------------------------------------------------------- layered LLR schedule (llr_layered.cu)
def _greedy_layers(t):
    """The colouring of csrc/llr_layered.cu: checks in index order take the lowest colour no neighbour check has."""
    sc, dc, tc = np.asarray(t.inbox_start_chk), np.asarray(t.degree_chk), np.asarray(t.target_cells_chk)
    sv, dv = np.asarray(t.inbox_start_var), np.asarray(t.degree_var)
    var_of_vrow = np.repeat(np.arange(t.n_var), dv)            # variable of every VN-major row
    vidx = var_of_vrow[tc]                                     # variable of every CN-major slot
    chk_of_crow = np.repeat(np.arange(t.n_chk), dc)
    checks_of_var = [[] for _ in range(t.n_var)]
    for e in range(t.n_edge):
        checks_of_var[vidx[e]].append(chk_of_crow[e])
    colour = -np.ones(t.n_chk, dtype=int)
    for c in range(t.n_chk):
        used = {colour[o] for k in range(dc[c]) for o in checks_of_var[vidx[sc[c] + k]] if colour[o] >= 0}
        col = 0
        while col in used:
            col += 1
        colour[c] = col
    return colour, vidx


def _clip150(x):
    return np.where(x > 150, 150.0, np.where(x < -150, -150.0, np.where(x == 0, 0.0, x)))


def _boxplus(a, b):
    with np.errstate(over="ignore"):
        return _clip150(np.log((1.0 + np.exp(a + b)) / (np.exp(a) + np.exp(b))))


def _layered_reference(t, ch, imax, algo, early):
    """numpy restatement of the layered schedule (header of csrc/llr_layered.cu), float64."""
    colour, vidx = _greedy_layers(t)
    sc, dc = np.asarray(t.inbox_start_chk), np.asarray(t.degree_chk)
    L = ch.astype(np.float64).copy()
    R = np.zeros((t.n_edge, ch.shape[1]))
    order = [c for l in range(colour.max() + 1) for c in np.nonzero(colour == l)[0]]
    passes = 0
    for it in range(imax - 1):
        for c in order:
            s, d = sc[c], dc[c]
            v = vidx[s:s + d]
            x = L[v] - R[s:s + d]
            q = _clip150(x)
            o = np.empty_like(q)
            if algo == "minsum":
                if d == 2:
                    o[0], o[1] = q[1], q[0]
                else:
                    aq, neg = np.abs(q), q < 0
                    for k in range(d):
                        mag = np.delete(aq, k, axis=0).min(axis=0)
                        sneg = np.logical_xor.reduce(np.delete(neg, k, axis=0), axis=0)
                        o[k] = np.where(mag == 0, 0.0, np.where(sneg, -mag, mag))
            else:
                if d >= 4:   # forward/backward box-plus, the order of llr_cn_compute<ALGO 2>
                    fw, bw = [q[0]], {d - 1: q[d - 1]}
                    for k in range(1, d - 1):
                        fw.append(_boxplus(q[k], fw[k - 1]))
                    for k in range(d - 2, 0, -1):
                        bw[k] = _boxplus(q[k], bw[k + 1])
                    o[0], o[d - 1] = _clip150(bw[1]), _clip150(fw[d - 2])
                    for k in range(1, d - 1):
                        o[k] = _clip150(_boxplus(fw[k - 1], bw[k + 1]))
                else:        # sequential chains in slot order
                    for k in range(d):
                        others = [j for j in range(d) if j != k]
                        tt = q[others[0]]
                        for j in others[1:]:
                            tt = _boxplus(q[j], tt)
                        o[k] = _clip150(tt)
            R[s:s + d] = o
            L[v] = x + o
        passes = it + 1
        if early:
            hard = (L < 0).astype(np.int64)
            bad = False
            for c in range(t.n_chk):
                if (hard[vidx[sc[c]:sc[c] + dc[c]]].sum(axis=0) & 1).any():
                    bad = True
                    break
            if not bad:
                break
    return L, passes + 1


@pytest.mark.parametrize("algo", ["minsum", "bp"])
@pytest.mark.parametrize("code,B", [("wlan648", 37), ("reg36", 64), ("deg12", 5)])
def test_layered_schedule_matches_its_numpy_restatement(gpu, algo, code, B):
    """schedule='layered' (opt-in): a-posteriori LLRs and i_num against the numpy restatement of the
    documented schedule -- min-sum bit for bit, box-plus within 1e-9 and with identical hard decisions -- with early
    termination off and on, unaligned batch sizes, regular / quasi-cyclic / mixed-degree codes."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    if code == "wlan648":     # the 802.11n prototype (d_c {7,8}, d_v {2,3,4,11}) expanded with a small lifting size
        proto = codes.wlan_prototype(54)
        H = codes.qc_expand(np.where(proto >= 0, proto % 27, -1), 27)
    elif code == "reg36":
        H = codes.regular_random(240, 3, 6, seed=9)
    else:
        H = codes.random_from_degrees([1] * 2 + [2] * 30 + [3] * 40 + [5] * 12, [2] * 4 + [4] * 16 + [6] * 16 + [8] * 5 + [11] * 2 + [12] * 1, seed=2)
    t = graph.edge_tables(H)
    cls = pkg.Min_Sum_Decoder_class_irregular if algo == "minsum" else pkg.BeliefPropagationDecoderClassIrregular
    imax = 7
    dec = cls(H, imax, 16, B)
    dec.init_OpenCL_decoding(B)
    dec.schedule = 'layered'
    colour, _ = _greedy_layers(t)
    assert dec.layer_count() == colour.max() + 1 >= int(np.asarray(t.degree_var).max())
    rng = np.random.Generator(np.random.PCG64(11))
    sigma2 = 10 ** (-2.5 / 10)
    ch = 2.0 / sigma2 * (1.0 + np.sqrt(sigma2) * rng.standard_normal((t.n_var, B)))
    ch[3, 0] = 0.0                                                # a zero input (sign(0) = 0 rule)
    for early in (False, True):
        dec.early_termination = early
        out = dec.decode(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True).get()
        ref, ref_inum = _layered_reference(t, ch, imax, algo, early)
        assert dec.last_i_num == ref_inum
        if algo == "minsum":
            assert np.array_equal(out, ref)
        else:
            assert np.allclose(out, ref, rtol=0, atol=1e-9) and np.array_equal(out < 0, ref < 0)
    # host-array contract and the float32 kernels run the same schedule
    host = dec.decode(ch)
    assert np.array_equal(host, out) if algo == "minsum" else np.allclose(host, out, atol=1e-9)
    dec.precision = 'f32'
    out32 = dec.decode(ch)
    assert ((out32 < 0) == (out < 0)).mean() > 0.98
    dec.schedule = 'row'
    with pytest.raises(ValueError):
        dec.decode(ch)


def test_layered_schedule_needs_fewer_passes_than_flooding(gpu):
    """The point of the schedule: at the same Eb/N0 the layered decoder stops (syndrome zero for the whole batch) after
    clearly fewer passes than the flooding decoder, and what it returns are codewords."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    H = codes.regular_random(2000, 3, 6, seed=5)
    t = graph.edge_tables(H)
    B, imax = 256, 60
    rng = np.random.Generator(np.random.PCG64(3))
    sigma2 = 10 ** (-3.0 / 10)
    ch = torch.from_numpy(2.0 / sigma2 * (1.0 + np.sqrt(sigma2) * rng.standard_normal((t.n_var, B)))).cuda()
    inum = {}
    for sched in ("flooding", "layered"):
        dec = pkg.Min_Sum_Decoder_class_irregular(H, imax, 16, B)
        dec.init_OpenCL_decoding(B)
        dec.schedule = sched
        out = dec.decode(pkg.DeviceArray(ch), buffer_in=True, return_buffer=True).get()
        inum[sched] = dec.last_i_num
        assert inum[sched] < imax                               # the batch converged
        hard = (out < 0).astype(np.int64)
        Hd = H.toarray() if hasattr(H, "toarray") else np.asarray(H)
        assert not ((Hd @ hard) & 1).any()
    assert inum["layered"] <= 0.7 * inum["flooding"], inum


@pytest.mark.parametrize("case", ["global_syndrome_flags", "all_converge_early", "none_converges", "imax2", "imax1"])
def test_per_frame_early_termination_corner_cases(gpu, case):
    """Paths of ib_perframe.cu the main test does not reach: a batch whose syndrome accumulator does not fit behind the
    table image (flags OR-ed into global memory, FS = 2), a batch that is done after a few passes (`done` raised early),
    one in which no frame ever converges (no compaction, everything decided by the last group), and the shortest
    schedules (i_max = 2: one pass; i_max = 1: no pass at all, group 0)."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_irregular_config
    T = 16
    imax = {"imax2": 2, "imax1": 1}.get(case, 12)
    B = 140000 if case == "global_syndrome_flags" else 1500
    ebn0 = {"all_converge_early": 6.0, "none_converges": -3.0}.get(case, 2.2)
    H = codes.wlan_80211n(54)
    tb, _ = generate_irregular_config(1.0, H, T, imax)
    dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                    tb.matching_vector_checknode, tb.matching_vector_varnode, B)
    t = graph.edge_tables(H)
    dec.init_OpenCL_decoding(B)
    q = pkg.AWGN_Channel_Quantizer(10 ** (-ebn0 / 10) / (2 * 0.5), 3, T, 2000)
    q.init_OpenCL_quanti(t.n_var, B, return_buffer_only=True)
    ch_dev = q.quantize_direct_OpenCL(t.n_var, B)
    if case == "all_converge_early":
        # a nearly noiseless channel: the most reliable "bit 0" cluster everywhere except 0.2 % unreliable symbols
        rng = np.random.Generator(np.random.PCG64(5))
        chn = np.full((t.n_var, B), T - 1, dtype=np.uint8)
        hit = rng.random((t.n_var, B)) < 0.002
        chn[hit] = rng.integers(T // 2 - 2, T // 2 + 2, size=int(hit.sum())).astype(np.uint8)   # unreliable, not confidently wrong
        ch_dev = pkg.DeviceArray(torch.from_numpy(chn).cuda())
    dec.early_termination = 'frame'
    out = dec.decode_OpenCL(ch_dev, buffer_in=True, return_buffer=True).get()
    inum = dec.last_i_num_per_frame.get()
    sel = np.r_[0:40, B // 2:B // 2 + 40, B - 40:B]
    ch = ch_dev.get()[:, sel]
    ref, ref_inum = _per_frame_oracle(t, np.ascontiguousarray(ch), T, imax, tb)
    assert np.array_equal(inum[sel], ref_inum), (inum[sel][:20], ref_inum[:20])
    assert np.array_equal(out[:, sel], ref)
    assert dec.last_i_num == int(inum.max())
    if case == "all_converge_early":
        assert inum.max() < imax
    if case == "none_converges":
        assert inum.min() == imax


@pytest.mark.parametrize("kernel", ["phase_images", "restaging"])
@pytest.mark.parametrize("code", ["reg36", "wlan1296", "dvb6480"])
@pytest.mark.parametrize("B", [1, 2, 9, 33, 100, 256, 257, 520])
def test_small_batch_cooperative_kernels_lane_mode(gpu, code, B, kernel, monkeypatch):
    """Batches of up to 256 frames run the whole-decode cooperative kernels with one LANE per (node, word) pair
    (cn_lanes_n4 / vn_lanes_n4: the reference's DVB-S2 drivers decode msg_at_time = 2 frames per call); 257 and 520 take the
    warp-per-(node, tile) bodies of the same kernels.  Outputs and i_num against the oracle, early termination off and on.
    kernel = phase_images: ib_coop_phase_kernel (TMA-staged images of whole phases, the default up to 256 frames);
    restaging: the kernels of ib_coop_n4.cuh (IBLDPC_NO_COOP_PHASE=1).  IBLDPC_COOP_MAX_B pins the cooperative kernels for
    257 and 520 frames too: by default batches above 256 frames of the instantiated sets run the fused per-phase kernels
    (test_batch_size_policy)."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    monkeypatch.setenv("IBLDPC_COOP_MAX_B", "4096")
    if kernel == "restaging":
        monkeypatch.setenv("IBLDPC_NO_COOP_PHASE", "1")
    T, imax = 16, 7
    if code == "reg36":
        H = codes.regular_random(2000, 3, 6, seed=5)
    else:
        H = codes.wlan_80211n(54) if code == "wlan1296" else codes.dvbs2_like_half_rate(6480, q_groups=36)
    t = graph.edge_tables(H)
    tb = luts.random_tables(T, t.d_c_max, t.d_v_max, imax, seed=B, matching=code != "reg36")
    if code == "reg36":
        dec = pkg.Discrete_LDPC_Decoder_class(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, B)
    else:
        dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                        tb.matching_vector_checknode, tb.matching_vector_varnode, B)
    dec.init_OpenCL_decoding(B)
    rng = np.random.Generator(np.random.PCG64(100 + B))
    ch = rng.integers(0, T, size=(t.n_var, B)).astype(np.uint8)
    ch[:, 0] = T - 1                                  # one frame that converges at once: the batch stop must wait for the others
    for early in (False, True):
        dec.early_termination = early
        out = dec.decode_OpenCL(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True).get()
        assert dec.info()[1] <= 3, "expected the single cooperative launch (+ pack / pad kernels)"
        ref, ref_inum = _oracle_ib(t, ch, T, imax, tb, early)
        assert dec.last_i_num == ref_inum
        assert np.array_equal(out, ref.astype(np.uint8))


@pytest.mark.parametrize("code,B,launches", [("reg36", 100, 3), ("reg36", 1000, 3), ("reg36", 3000, 2 * 7 + 1), ("reg36", 5000, 2 * 7 + 1),
                                             ("wlan1296", 2, 3), ("wlan1296", 256, 3), ("wlan1296", 2000, 3), ("wlan1296", 3000, 2 * 7 + 1),
                                             ("dvb6480", 2, 3), ("dvb6480", 600, 3), ("dvb6480", 3000, 2 * 7 + 1),
                                             ("dvb16200", 256, 3), ("dvb16200", 2048, 2 * 7 + 1)])   # 56699 edges x 1 KB > 32 MB
def test_batch_size_policy(gpu, code, B, launches):
    """Default dispatch by batch size (end of ibldpc_set_luts) for the instantiated degree sets: one cooperative launch over
    the phase images up to 2048 frames (above 256 only while the packed messages stay below 32 MB); above, 802.11n sets ->
    fused per-phase kernels always, (3,6) and DVB-S2 sets -> fused per-phase kernels up to 4096 frames, one launch per
    degree class above ((3,6): one class per phase, so the launch count is the same).  Results against the oracle."""
    import torch
    import informationbottleneckdecodingldpc_b200 as pkg
    T, imax = 16, 7
    H = codes.regular_random(2000, 3, 6, seed=5) if code == "reg36" else codes.wlan_80211n(54) if code == "wlan1296" \
        else codes.dvbs2_like_half_rate(6480, q_groups=36) if code == "dvb6480" else codes.dvbs2_like_half_rate(16200, q_groups=90)
    t = graph.edge_tables(H)
    tb = luts.random_tables(T, t.d_c_max, t.d_v_max, imax, seed=B, matching=code != "reg36")
    if code == "reg36":
        dec = pkg.Discrete_LDPC_Decoder_class(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, B)
    else:
        dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, imax, T, T, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                        tb.matching_vector_checknode, tb.matching_vector_varnode, B)
    dec.init_OpenCL_decoding(B)
    ch = np.random.Generator(np.random.PCG64(7 + B)).integers(0, T, size=(t.n_var, B)).astype(np.uint8)
    for early in (False, True):
        dec.early_termination = early
        out = dec.decode_OpenCL(pkg.DeviceArray(torch.from_numpy(ch).cuda()), buffer_in=True, return_buffer=True).get()
        assert dec.info()[1] <= launches + 1, (dec.info(), launches)      # + pack / pad kernel
        if launches > 3:
            assert dec.info()[1] >= launches
        ref, ref_inum = _oracle_ib(t, ch, T, imax, tb, early)
        assert dec.last_i_num == ref_inum
        assert np.array_equal(out, ref.astype(np.uint8))
