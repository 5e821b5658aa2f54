#!/usr/bin/env python3
"""BER simulation with encoded (non-zero) codewords -- the loop of the reference's
Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL_enc.py:98-160 with every stage on the GPU:
random bits -> systematic LDPC encoder -> BPSK -> AWGN channel -> |T|=16 quantizer -> IB decoder with message
alignment -> comparison with the transmitted information bits.  Frames are sharded over the ranks of a torchrun
job; the error counters are all-reduced with NCCL after every batch.

  python examples/BER_simulation_IB_irregular_enc.py
  torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 8 examples/BER_simulation_IB_irregular_enc.py --code dvbs2
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import informationbottleneckdecodingldpc_b200 as pkg
from informationbottleneckdecodingldpc_b200 import codes
from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_irregular_config
from informationbottleneckdecodingldpc_b200.engine import count_errors_async
from informationbottleneckdecodingldpc_b200.parallel import init_distributed

ap = argparse.ArgumentParser()
ap.add_argument("--code", default="wlan", choices=["wlan", "wlan1944", "dvbs2"])
ap.add_argument("--ebn0", type=float, nargs="+", default=None)
ap.add_argument("--msg-at-time", type=int, default=0)
ap.add_argument("--min-errors", type=int, default=2000)
ap.add_argument("--max-batches", type=int, default=40)
args = ap.parse_args()

rank, world, local = init_distributed()
torch.cuda.set_device(local)
H = {"wlan": lambda: codes.wlan_80211n(54), "wlan1944": lambda: codes.wlan_80211n(81),
     "dvbs2": codes.dvbs2_like_half_rate}[args.code]()
N_var = H.shape[1]
msg_at_time = args.msg_at_time or (512 if args.code == "dvbs2" else 16384)
AD_max_abs, cardinality_Y_channel, cardinality_T_channel, cardinality_T_decoder_ops, imax = 3, 2000, 16, 16, 50

cfg, _ = generate_irregular_config(1.0, H, cardinality_T_decoder_ops, imax)
transi = pkg.LDPC_BPSK_Transmitter(H, msg_at_time)
transi.return_buffer_only = True
transi.set_stream(rank)                               # independent Philox sub-stream per rank
decodi = pkg.Discrete_LDPC_Decoder_class_irregular(H, imax, cardinality_T_channel, cardinality_T_decoder_ops,
                                                   cfg.Trellis_checknodevector_a, cfg.Trellis_varnodevector_a,
                                                   cfg.matching_vector_checknode, cfg.matching_vector_varnode, msg_at_time)
if rank == 0:
    print(f"# {args.code}: N={N_var} K={transi.data_len} encoder '{transi.encoder.EncodingAlgorithm}', "
          f"{msg_at_time} frames per GPU and batch, {world} GPU(s)")
for EbN0_dB in (args.ebn0 or ([0.8, 1.0, 1.2] if args.code == "dvbs2" else [1.0, 1.5, 2.0, 2.5])):
    sigma_n2 = 10 ** (-EbN0_dB / 10) / (2 * transi.R_c)
    chani = pkg.AWGN_channel(sigma_n2)
    chani.set_stream(rank)
    quanti = pkg.AWGN_Channel_Quantizer(sigma_n2, AD_max_abs, cardinality_T_channel, cardinality_Y_channel)
    quanti.init_OpenCL_quanti(N_var, msg_at_time, return_buffer_only=True)
    decodi.init_OpenCL_decoding(msg_at_time, quanti.context)
    totals = torch.zeros(4, dtype=torch.int64, device="cuda")
    start = time.time()
    batches = 0
    while batches < args.max_batches:
        coded = transi.transmit_bits()                                        # random bits + encoder
        rec_data_quantized = quanti.quantize_OpenCL(chani.transmission_bits(coded))   # BPSK + AWGN, quantizer
        decoded_mat = decodi.decode_OpenCL(rec_data_quantized, buffer_in=True, return_buffer=True)
        c = torch.zeros(4, dtype=torch.int64, device="cuda")
        count_errors_async(decoded_mat, transi.data_len, cardinality_T_decoder_ops // 2, c,
                           ref_bits=transi.last_transmitted_bits)
        c[2] += msg_at_time
        if world > 1:
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
        totals.add_(c)
        batches += 1
        if batches % 4 == 0 and int(totals[0]) >= args.min_errors:           # one host read every 4 batches
            break
    tot = totals.tolist()
    spent = time.time() - start
    if rank == 0:
        print(f"EbN0_dB={EbN0_dB:.2f} frames={tot[2]} BER={tot[0] / (tot[2] * transi.data_len):.3e} "
              f"FER={tot[1] / tot[2]:.3e} datarate_Bps={tot[2] * transi.data_len / spent:.3e}", flush=True)
if world > 1:
    dist.destroy_process_group()
