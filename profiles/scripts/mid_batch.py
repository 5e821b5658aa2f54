"""Decode time vs batch size in the middle range (512..16384 frames): fused per-phase kernels vs one launch per degree class
(run with IBLDPC_COOP_MAX_B=0 and with / without IBLDPC_PHASE=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import informationbottleneckdecodingldpc_b200 as pkg
from informationbottleneckdecodingldpc_b200 import codes, luts, graph
for name, H, irregular, Bs in (("c1", codes.regular_random(8000, 3, 6, seed=20181001), False, (512, 1024, 2048, 4096, 8192)),
                               ("dvbs2", codes.dvbs2_like_half_rate(), True, (256, 512, 1024, 2048, 4096)),
                               ("wlan1296", codes.wlan_80211n(54), True, (512, 1024, 2048, 4096, 8192))):
    t = graph.edge_tables(H)
    tb = luts.random_tables(16, t.d_c_max, t.d_v_max, 50, seed=1, matching=irregular)
    K = t.n_var - t.n_chk
    for B in Bs:
        if irregular:
            dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, 50, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                            tb.matching_vector_checknode, tb.matching_vector_varnode, B)
        else:
            dec = pkg.Discrete_LDPC_Decoder_class(H, 50, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a, B)
        dec.early_termination = False
        ch = torch.randint(0, 16, (t.n_var, B), dtype=torch.uint8, device="cuda")
        for _ in range(3): dec.decode_OpenCL(ch, buffer_in=True, return_buffer=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 10
        for _ in range(n): dec.decode_OpenCL(ch, buffer_in=True, return_buffer=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
        print(f"{name} B={B:6d} {dt*1e3:8.3f} ms/decode {K*B/dt/1e9:7.3f} Gbit/s launches={dec.info()[1]}", flush=True)
