#!/usr/bin/env python3
"""Flooding vs layered schedule of the float64 min-sum / BP decoders with early termination on (the BER drivers' mode):
ms per batch, passes until the batch stops, bit errors.  (3,6) n=8000 and 802.11n n=1296."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import informationbottleneckdecodingldpc_b200 as pkg
from informationbottleneckdecodingldpc_b200 import codes


def run(name, H, B, ebn0, imax=50):
    N = H.shape[1]
    K = N - H.shape[0]
    q = pkg.AWGN_Channel_Quantizer(10 ** (-ebn0 / 10) / (2 * K / N), 3, 16, 2000)
    q.llr_dtype = np.float64
    q.init_OpenCL_quanti(N, B, return_buffer_only=True)
    llr = q.quantize_direct_OpenCL_LLR(N, B)
    for cls, algo in ((pkg.Min_Sum_Decoder_class_irregular, "min-sum"), (pkg.BeliefPropagationDecoderClassIrregular, "BP")):
        for sched in ("flooding", "layered"):
            for early in (False, True):
                dec = cls(H, imax, 16, B)
                dec.init_OpenCL_decoding(B)
                dec.schedule = sched
                dec.early_termination = early
                for _ in range(2):
                    out = dec.decode(llr, buffer_in=True, return_buffer=True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    out = dec.decode(llr, buffer_in=True, return_buffer=True)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 3
                extra = f" layers {dec.layer_count()}" if sched == "layered" else ""
                print(f"{name} {ebn0} dB {algo} f64 {sched:8s} ET {'on ' if early else 'off'}: {ms:8.2f} ms/batch B={B} -> {K * B / ms / 1e6:.3f} Gbit/s, "
                      f"i_num {dec.last_i_num}, launches {dec.info()[1]}, bit errors {dec.return_errors_all_zero(out)}{extra}", flush=True)
                del dec, out
                torch.cuda.empty_cache()


run("c1", codes.regular_random(8000, 3, 6), 16384, 2.4)
run("wlan1296", codes.wlan_80211n(54), 32768, 2.5)
