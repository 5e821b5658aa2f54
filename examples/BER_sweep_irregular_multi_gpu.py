#!/usr/bin/env python3
"""Eb/N0 sweep on an irregular code: IB decoder with message alignment against the min-sum and the
belief-propagation benchmark decoders -- the loops of the reference's
Irregular_LDPC_Decoding/{WLAN,DVB-S2}/BER_simulation_OpenCL.py, ..._min_sum.py and ..._quant_BP.py
in one run on the B200 engine, frames sharded over the ranks of a torchrun job and the bit/frame
error counters all-reduced with NCCL after every batch (simulation.ber_point).

  python examples/BER_sweep_irregular_multi_gpu.py --code wlan
  torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 8 \\
      examples/BER_sweep_irregular_multi_gpu.py --code dvbs2 --ebn0 0.8 1.0 1.2 --msg-at-time 1024

All three decoders see channel outputs of the same |T|=16 quantizer (the benchmark decoders get
its cluster LLRs, quantize_direct_OpenCL_LLR, as in the reference), all-zero codeword, BPSK/AWGN.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import informationbottleneckdecodingldpc_b200 as pkg
from informationbottleneckdecodingldpc_b200 import codes
from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_irregular_config
from informationbottleneckdecodingldpc_b200.parallel import init_distributed
from informationbottleneckdecodingldpc_b200.simulation import ber_point

ap = argparse.ArgumentParser()
ap.add_argument("--code", default="wlan", choices=["wlan", "wlan1944", "dvbs2"])
ap.add_argument("--ebn0", type=float, nargs="+", default=None, help="Eb/N0 points in dB")
ap.add_argument("--design-ebn0", type=float, default=1.0, help="Eb/N0 the IB tables are designed for")
ap.add_argument("--msg-at-time", type=int, default=0, help="frames per GPU and batch")
ap.add_argument("--min-errors", type=int, default=2000)
ap.add_argument("--max-batches", type=int, default=50)
ap.add_argument("--imax", type=int, default=50)
ap.add_argument("--llr-precision", default="f64", choices=["f32", "f64"],
                help="f64 = reference arithmetic (meets the >= 99.99 % identical-frames bar); f32 is faster but outside that tolerance")
args = ap.parse_args()

rank, world, local = init_distributed()
torch.cuda.set_device(local)

H = {"wlan": lambda: codes.wlan_80211n(54), "wlan1944": lambda: codes.wlan_80211n(81),
     "dvbs2": codes.dvbs2_like_half_rate}[args.code]()
N_var = H.shape[1]
msg_at_time = args.msg_at_time or (512 if args.code == "dvbs2" else 16384)
ebn0_points = args.ebn0 or ([0.8, 1.0, 1.2] if args.code == "dvbs2" else [1.0, 1.5, 2.0, 2.5])
AD_max_abs, cardinality_Y_channel, cardinality_T_channel, cardinality_T_decoder_ops = 3, 2000, 16, 16

# "Load stored data": the reference unpickles decoder_config_EbN0_gen_*.pkl; here the config is designed on the fly
cfg, _ = generate_irregular_config(args.design_ebn0, H, cardinality_T_decoder_ops, args.imax)
decoders = {
    "IB+align": pkg.Discrete_LDPC_Decoder_class_irregular(H, args.imax, cardinality_T_channel, cardinality_T_decoder_ops,
                                                          cfg.Trellis_checknodevector_a, cfg.Trellis_varnodevector_a,
                                                          cfg.matching_vector_checknode, cfg.matching_vector_varnode,
                                                          msg_at_time),
    "min-sum": pkg.Min_Sum_Decoder_class_irregular(H, args.imax, cardinality_T_channel, msg_at_time),
    "BP": pkg.BeliefPropagationDecoderClassIrregular(H, args.imax, cardinality_T_channel, msg_at_time),
}
for name in ("min-sum", "BP"):
    decoders[name].precision = args.llr_precision
R_c = decoders["IB+align"].R_c

if rank == 0:
    print(f"# {args.code}: N={N_var}, R_c={R_c:.4f}, i_max={args.imax}, {msg_at_time} frames per GPU and batch, {world} GPU(s)")
    print(f"# {'EbN0_dB':>7s} {'decoder':>9s} {'frames':>10s} {'BER':>10s} {'FER':>10s} {'info Mbit/s':>12s}")
for EbN0_dB in ebn0_points:
    sigma_n2 = 10 ** (-EbN0_dB / 10) / (2 * R_c)
    for name, decodi in decoders.items():
        quanti = pkg.AWGN_Channel_Quantizer(sigma_n2, AD_max_abs, cardinality_T_channel, cardinality_Y_channel)
        quanti.set_stream(rank)                    # independent Philox sub-stream per rank
        quanti.init_OpenCL_quanti(N_var, msg_at_time, return_buffer_only=True)
        decodi.init_OpenCL_decoding(msg_at_time, quanti.context)
        res = ber_point(decodi, quanti, msg_at_time, min_errors=args.min_errors,
                        max_frames=args.max_batches * msg_at_time * world, llr=(name != "IB+align"))
        if rank == 0:
            print(f"  {EbN0_dB:7.2f} {name:>9s} {res['frames']:10d} {res['ber']:10.3e} {res['fer']:10.3e} "
                  f"{res['info_bit_rate'] / 1e6:12.1f}", flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
