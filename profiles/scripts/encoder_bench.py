"""Throughput of the transmit-side kernels (scratch): encoder, random bits, BPSK+AWGN, quantizer."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import informationbottleneckdecodingldpc_b200 as pkg
from informationbottleneckdecodingldpc_b200 import codes
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
for name, H, B in (("wlan1296", codes.wlan_80211n(54), 100096), ("wlan1944", codes.wlan_80211n(81), 65536), ("dvbs2", codes.dvbs2_like_half_rate(), 8192)):
    tr = pkg.LDPC_BPSK_Transmitter(H, B); tr.return_buffer_only = True
    K, N = tr.data_len, tr.codeword_len
    bits = tr.random_bits()
    t_bits = timeit(tr.random_bits)
    t_enc = timeit(lambda: tr.encoder.encode_batch(bits))
    coded = tr.encoder.encode_batch(bits)
    ch = pkg.AWGN_channel(0.5)
    t_ch = timeit(lambda: ch.transmission_bits(coded))
    q = pkg.AWGN_Channel_Quantizer(0.5, 3, 16, 2000); q.init_OpenCL_quanti(N, B, return_buffer_only=True)
    y = ch.transmission_bits(coded)
    t_q = timeit(lambda: q.quantize_OpenCL(y))
    print(f"{name}: B={B} '{tr.encoder.EncodingAlgorithm}' bits {t_bits*1e3:.2f} ms, encode {t_enc*1e3:.2f} ms "
          f"({K*B/t_enc/1e9:.1f} Gbit/s info), BPSK+AWGN {t_ch*1e3:.2f} ms ({N*B/t_ch/1e9:.1f} Gsym/s), quantize {t_q*1e3:.2f} ms", flush=True)
