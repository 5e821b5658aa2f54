#!/usr/bin/env python3
"""BER simulation of the IB decoder on the (3,6) n=8000 code -- the loop of the reference's
Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py on the B200 engine.

  python examples/BER_simulation_IB_regular.py                       # one GPU
  torchrun --standalone --nproc-per-node 8 examples/BER_simulation_IB_regular.py   # frames sharded over 8 GPUs
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import informationbottleneckdecodingldpc_b200 as pkg
from informationbottleneckdecodingldpc_b200 import codes
from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_regular_config
from informationbottleneckdecodingldpc_b200.parallel import init_distributed
from informationbottleneckdecodingldpc_b200.simulation import ber_point

rank, world, local = init_distributed()
torch.cuda.set_device(local)

# "Load stored data": the reference unpickles decoder_config_EbN0_gen_*.pkl; here the config is designed on the fly
generated_decoder, _ = generate_regular_config(1.2, 3, 6, 16, 50)
Trellis_checknodevector_a = generated_decoder.Trellis_checknodevector_a
Trellis_varnodevector_a = generated_decoder.Trellis_varnodevector_a
AD_max_abs, cardinality_Y_channel, cardinality_T_channel, cardinality_T_decoder_ops = 3, 2000, 16, 16
msg_at_time, min_errors, imax, N_var = 4096, 2000, 50, 8000
H = codes.regular_random(N_var, 3, 6)

decodi = pkg.Discrete_LDPC_Decoder_class(H, imax, cardinality_T_channel, cardinality_T_decoder_ops,
                                         Trellis_checknodevector_a, Trellis_varnodevector_a, msg_at_time)
R_c = 0.5
for EbN0_dB in np.arange(1.0, 1.81, 0.2):
    sigma_n2 = 10 ** (-EbN0_dB / 10) / (2 * R_c)
    quanti = pkg.AWGN_Channel_Quantizer(sigma_n2, AD_max_abs, cardinality_T_channel, cardinality_Y_channel)
    quanti.set_stream(rank)
    quanti.init_OpenCL_quanti(N_var, msg_at_time, return_buffer_only=True)
    decodi.init_OpenCL_decoding(msg_at_time, quanti.context)
    res = ber_point(decodi, quanti, msg_at_time, min_errors=min_errors, max_frames=400 * msg_at_time * world)
    if rank == 0:
        print(f"EbN0_dB={EbN0_dB:.1f} frames={res['frames']} BER={res['ber']:.3e} FER={res['fer']:.3e} "
              f"datarate_Bps={res['info_bit_rate']:.3e} ({world} GPU)")
    if res["bit_errors"] == 0:
        break
