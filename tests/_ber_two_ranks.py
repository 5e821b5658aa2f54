"""Helper of tests/test_gpu_round2.py::test_ber_point_two_ranks (launched with torchrun, 2 ranks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import informationbottleneckdecodingldpc_b200 as pkg
from informationbottleneckdecodingldpc_b200 import codes
from informationbottleneckdecodingldpc_b200.decoder_config_generation import generate_irregular_config
from informationbottleneckdecodingldpc_b200.parallel import init_distributed
from informationbottleneckdecodingldpc_b200.simulation import ber_point

rank, world, local = init_distributed()
torch.cuda.set_device(local)
H = codes.wlan_80211n(54)
tb, _ = generate_irregular_config(1.0, H, 16, 20)
dec = pkg.Discrete_LDPC_Decoder_class_irregular(H, 20, 16, 16, tb.Trellis_checknodevector_a, tb.Trellis_varnodevector_a,
                                                tb.matching_vector_checknode, tb.matching_vector_varnode, 512)
q = pkg.AWGN_Channel_Quantizer(10 ** (-1.5 / 10) / (2 * 0.5), 3, 16, 2000)
q.init_OpenCL_quanti(1296, 512, return_buffer_only=True)      # sub-stream = rank (never chosen explicitly)
dec.init_OpenCL_decoding(512, q.context)
assert q.stream == rank
first = q.quantize_direct_OpenCL(1296, 512).tensor.clone()
gathered = [torch.empty_like(first) for _ in range(world)]
dist.all_gather(gathered, first)
assert not torch.equal(gathered[0], gathered[1]), "both ranks drew the same channel realisations"
q.set_stream(rank)
res = ber_point(dec, q, 512, min_errors=800)
tot = torch.tensor([res["bit_errors"], res["frame_errors"], res["frames"]], dtype=torch.int64, device="cuda")
both = [torch.empty_like(tot) for _ in range(world)]
dist.all_gather(both, tot)
assert torch.equal(both[0], both[1]), "ranks disagree on the all-reduced totals"
assert res["frames"] % (512 * world) == 0 and res["bit_errors"] >= 800
# the totals are the sum of two DIFFERENT per-rank contributions: replay this rank's own stream
q.set_stream(rank)
mine = 0
for _ in range(res["frames"] // (512 * world)):
    out = dec.decode_OpenCL(q.quantize_direct_OpenCL(1296, 512), buffer_in=True, return_buffer=True)
    mine += dec.return_errors_all_zero(out)
parts = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
dist.all_gather(parts, torch.tensor([mine], dtype=torch.int64, device="cuda"))
assert int(parts[0]) + int(parts[1]) == res["bit_errors"] and int(parts[0]) != int(parts[1])
if rank == 0:
    print("TWO_RANK_OK", os.environ.get("IBLDPC_ABI_ALLREDUCE"), res["bit_errors"], res["frames"], getattr(dec, "_nccl_ready", False))
dist.destroy_process_group()
