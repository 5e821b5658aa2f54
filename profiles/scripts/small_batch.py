"""Decode time vs batch size (launch-bound regime), C1 tables, ET off and on (scratch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import informationbottleneckdecodingldpc_b200 as pkg
from informationbottleneckdecodingldpc_b200 import codes
from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_regular_config
H = codes.regular_random(8000, 3, 6, seed=20181001)
tb, _ = generate_regular_config(1.2, 3, 6, 16, 50)
q = pkg.AWGN_Channel_Quantizer(10 ** (-1.6 / 10) / (2 * 0.5), 3, 16, 2000)
for B in (2, 16, 64, 100, 256, 512, 2048):
    q.init_OpenCL_quanti(8000, B, return_buffer_only=True)
    dec = pkg.Discrete_LDPC_Decoder_class(H, 50, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, B)
    dec.init_OpenCL_decoding(B, q.context)
    rec = q.quantize_direct_OpenCL(8000, B)
    for et in (False, True):
        dec.early_termination = et
        for _ in range(3): dec.decode_OpenCL(rec, buffer_in=True, return_buffer=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 20
        for _ in range(n): dec.decode_OpenCL(rec, buffer_in=True, return_buffer=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
        print(f"B={B:5d} ET={int(et)} {dt*1e3:8.3f} ms/decode  {4000*B/dt/1e9:7.3f} Gbit/s  i_num={dec.last_i_num}", flush=True)
