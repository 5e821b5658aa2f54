"""Replay of the reference's own BER driver scripts on the B200 engine (north_star: the classes "drop in under the
BER_simulation_OpenCL_* drivers").

The UNMODIFIED driver sources (staged from the reference tree into the git-ignored oracle/_ref/drivers/ by
oracle/stage_drivers.py -- they must not enter the repository and /root/reference does not exist on the GPU box) are
executed through informationbottleneckdecodingldpc_b200.run_driver: module-path shadowing, matplotlib stubs, the
script's own working directory with a generated LDPC_codes/ file and decoder_config_*.pkl next to it.  Only the two
Monte-Carlo constants min_errors and EbN0_dB_max_value are overridden so that a run ends after two Eb/N0 points.
The resulting BER is compared with the CPU oracle decoding independent channel draws of the same quantizer and tables.
"""
import os
import re
import shutil

import numpy as np
import pytest

from informationbottleneckdecodingldpc_b200 import codes, graph, luts
from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_irregular_config, generate_regular_config

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = 16


def _staged(rel):
    from oracle import stage_drivers
    p = stage_drivers.staged(rel)
    if p is None and os.path.isdir(stage_drivers.DEFAULT_REF):
        stage_drivers.stage()
        p = stage_drivers.staged(rel)
    if p is None:
        pytest.skip("reference drivers are not staged (run __graft_entry__.build() where /root/reference is mounted)")
    return p


def _prepare(tmp_path, rel, H, configs):
    """Lay out <tmp>/<rel> with the code file where the script's `filepath` points and the pickles next to it."""
    src = _staged(rel)
    dst = tmp_path / rel
    dst.parent.mkdir(parents=True, exist_ok=True)
    shutil.copyfile(src, dst)
    text = open(src).read()
    m = re.search(r'^filepath\s*=\s*"([^"]+)"', text, re.M)
    assert m, "driver has no filepath assignment"
    code_file = (dst.parent / m.group(1)).resolve()
    code_file.parent.mkdir(parents=True, exist_ok=True)
    if str(code_file).endswith(".npy"):
        np.save(code_file, np.asarray(H.toarray(), dtype=np.int8))
    elif str(code_file).endswith(".npz"):
        codes.save_csr_npz(H, str(code_file))
    else:
        graph.write_alist(H, str(code_file))
    for name, (tables, extras) in configs.items():
        luts.save_config(tables, str(dst.parent / name), **extras)
    return str(dst)


def _oracle_ber(H, sigma_n2, rows, frames, ib=None, llr_algo=None, imax=50, seed=5):
    """BER of the CPU oracle on `frames` independent all-zero-codeword draws (early termination like the drivers)."""
    import informationbottleneckdecodingldpc_b200 as pkg
    from oracle import oracle
    t = graph.edge_tables(H)
    q = pkg.AWGN_Channel_Quantizer(sigma_n2, 3, T, 2000)
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.random(size=(t.n_var, frames))
    cl = ((u[:, :, None] - q.cdf_t_given_x_equals_zero) > 0).sum(2) - 1
    if ib is not None:
        out, _ = oracle.ib_decode(t, cl, T=T, imax=imax, cn_lut=ib.Trellis_checknodevector_a, vn_lut=ib.Trellis_varnodevector_a,
                                  cn_match=ib.matching_vector_checknode, vn_match=ib.matching_vector_varnode, early=True)
        bits = out[:rows] < T // 2
    else:
        out, _ = oracle.llr_decode(t, q.output_LLRs[cl], algo=llr_algo, imax=imax, early=True)
        bits = out[:rows] < 0
    per_frame = bits.mean(axis=0)
    return float(per_frame.mean()), float(per_frame.std(ddof=1) / np.sqrt(frames))


def _check(ns, H, rows, R_c, frames=192, denominator_rows=None, **kw):
    """The driver's BER at its first Eb/N0 point (0 dB) against the oracle's Monte-Carlo estimate."""
    ebn0 = np.asarray(ns["EbN0_dB_vector"], dtype=float)
    ber = np.asarray(ns["BER_vector"], dtype=float)
    assert ebn0.size == ber.size >= 2 and ebn0[0] == 0.0          # two points were simulated: 0.0 and 0.1 dB
    assert np.all(ber > 0) and np.all(ber < 0.5)
    sigma_n2 = 10 ** (-ebn0[0] / 10) / (2 * R_c)
    ref, se = _oracle_ber(H, sigma_n2, rows, frames, **kw)
    got = ber[0] * (denominator_rows / rows if denominator_rows else 1.0)   # drivers normalise by N_var or R_c*N_var
    # both are Monte-Carlo estimates (driver: >= 100 frames): 6 standard errors of the oracle's estimate + 5 % slack
    assert abs(got - ref) <= 6 * se * np.sqrt(2) + 0.05 * ref, (got, ref, se)


def test_regular_ib_driver(tmp_path):
    rel = "Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py"
    H = codes.regular_random(8000, 3, 6)
    tb, ex = generate_regular_config(1.05, 3, 6, T, 50)
    path = _prepare(tmp_path, rel, H, {"decoder_config_EbN0_gen_1.05_16.pkl": (tb, ex)})
    from informationbottleneckdecodingldpc_b200 import run_driver
    ns = run_driver.run(path, {"min_errors": 2000, "EbN0_dB_max_value": 0.05})
    assert ns["decodi"].__class__.__module__.startswith("informationbottleneckdecodingldpc_b200")
    _check(ns, H, 8000, 0.5, ib=tb)


def test_regular_minsum_driver(tmp_path):
    rel = "Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL_min_sum.py"
    H = codes.regular_random(8000, 3, 6)
    path = _prepare(tmp_path, rel, H, {})
    from informationbottleneckdecodingldpc_b200 import run_driver
    ns = run_driver.run(path, {"min_errors": 2000, "EbN0_dB_max_value": 0.05})
    dl = int(ns["decodi"].data_len)
    # return_errors_all_zero counts rows [:data_len]; the driver divides by R_c * N_var = data_len bits per frame
    _check(ns, H, dl, 0.5, frames=96, llr_algo="minsum")


@pytest.fixture(scope="module")
def wlan():
    H = codes.wlan_80211n(54)
    return H, generate_irregular_config(0.9, H, T, 50)


def test_wlan_ib_driver(tmp_path, wlan):
    rel = "Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL.py"
    H, (tb, ex) = wlan
    path = _prepare(tmp_path, rel, H, {"decoder_config_EbN0_gen_0.9_16adapt71.pkl": (tb, ex)})
    from informationbottleneckdecodingldpc_b200 import run_driver
    # the script's last statement notifies the author's phone through a Pushbullet client `pb` it never defines
    ns = run_driver.run(path, {"min_errors": 2000, "EbN0_dB_max_value": 0.05}, inject={"pb": run_driver.Silent()})
    dl = int(ns["decodi"].data_len)
    _check(ns, H, dl, float(ns["transi"].R_c), ib=tb)


def test_wlan_ib_enc_driver(tmp_path, wlan):
    """Transmitter + encoder + real AWGN channel + quantize_on_host + host-buffer decode (int numpy in/out)."""
    rel = "Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL_enc.py"
    H, (tb, ex) = wlan
    path = _prepare(tmp_path, rel, H, {"decoder_config_EbN0_gen_0.7_16cas.pkl": (tb, ex),
                                       "decoder_config_EbN0_gen_0.8_16cas.pkl": (tb, ex)})
    from informationbottleneckdecodingldpc_b200 import run_driver
    ns = run_driver.run(path, {"min_errors": 2000, "EbN0_dB_max_value": 0.05})
    dl = int(ns["transi"].data_len)
    # this driver normalises by R_c * N_var = data_len information bits per frame; random codewords through the real
    # channel have the same error statistics as the all-zero codeword through the symmetric quantizer
    _check(ns, H, dl, float(ns["transi"].R_c), ib=tb)


def test_wlan_quant_bp_driver(tmp_path, wlan):
    rel = "Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL_quant_BP.py"
    H, _ = wlan
    path = _prepare(tmp_path, rel, H, {})
    from informationbottleneckdecodingldpc_b200 import run_driver
    ns = run_driver.run(path, {"min_errors": 2000, "EbN0_dB_max_value": 0.05})
    dl = int(ns["decodi"].data_len)
    _check(ns, H, dl, float(ns["transi"].R_c), frames=96, llr_algo="bp")


def test_dvbs2_ib_driver(tmp_path):
    """Irregular_LDPC_Decoding/DVB-S2/BER_simulation_OpenCL.py (msg_at_time = 2: the cooperative small-batch kernel of the
    DVB-S2 degree sets) on a DVB-S2-like rate-1/2 code of length 6480 (the script reads the length from the code file)."""
    rel = "Irregular_LDPC_Decoding/DVB-S2/BER_simulation_OpenCL.py"
    H = codes.dvbs2_like_half_rate(6480, q_groups=36)
    tb, ex = generate_irregular_config(0.6, H, T, 50)
    path = _prepare(tmp_path, rel, H, {"decoder_config_EbN0_gen_0.6_16adapt71.pkl": (tb, ex)})
    from informationbottleneckdecodingldpc_b200 import run_driver
    ns = run_driver.run(path, {"min_errors": 3000, "EbN0_dB_max_value": 0.05}, inject={"pb": run_driver.Silent()})
    dl = int(ns["decodi"].data_len)
    assert ns["decodi"].info()[0] == 2 and ns["msg_at_time"] == 2
    _check(ns, H, dl, float(ns["transi"].R_c), frames=64, ib=tb)
